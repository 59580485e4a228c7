timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ring or peer" 2>&1 | tail -3
export B200_SPMV_VERBOSE=1
timeout 300 python scripts/sweep.py D/8 "pr!STAGE_X=0,pr" 30 2>&1 | grep -v "^libb200" | tee gpurun_out/sweep58.txt
timeout 300 python scripts/sweep.py D/4 "pr!STAGE_X=0,pr" 30 2>&1 | grep -v "^libb200" | tee -a gpurun_out/sweep58.txt
timeout 600 python scripts/sweep.py D "pr!STAGE_X=0,pr" 10 2>&1 | grep -v "^libb200" | tee -a gpurun_out/sweep58.txt
