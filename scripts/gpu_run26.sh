python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B200_SPMV_VERBOSE=1 python scripts/sweep.py pl22 "auto,sell:C=32,sell:C=64" 30 2>&1 | grep -v "^libb200" | tee gpurun_out/sweep26.txt
