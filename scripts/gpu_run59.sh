timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ring" 2>&1 | tail -3
timeout 300 python scripts/sweep.py D/8 "pr,pr:B=2" 30 2>&1 | grep -v "^libb200" | tee gpurun_out/sweep59.txt
timeout 300 python scripts/sweep.py D/4 "pr,pr:B=2" 30 2>&1 | grep -v "^libb200" | tee -a gpurun_out/sweep59.txt
