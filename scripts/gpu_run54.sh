timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ring" 2>&1 | tail -3
export B200_SPMV_VERBOSE=1
timeout 400 python scripts/sweep.py D/8 "sell,pr,pr:B=2" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep54.txt
timeout 300 python scripts/sweep.py D/4 "pr" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep54.txt
timeout 300 python scripts/sweep.py C "panel,pr:G=2" 200 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep54.txt
