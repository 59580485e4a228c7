timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "panel" 2>&1 | tail -3
timeout 300 python scripts/sweep.py C "panel,panel" 200 2>&1 | tee gpurun_out/sweep40.txt
timeout 300 python scripts/sweep.py B "panel" 200 2>&1 | tee -a gpurun_out/sweep40.txt
