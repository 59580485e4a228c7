timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ring or peer or npb" 2>&1 | tail -3
timeout 300 python scripts/sweep.py D/8 "auto" 30 2>&1 | grep -v "^libb200"
