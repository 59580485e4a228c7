timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "panel_kernel_bit_exact or golden or kat" 2>&1 | tail -3
timeout 300 python scripts/sweep.py C "panel,panel" 300 2>&1 | grep -v "^libb200"
timeout 300 python scripts/sweep.py B "panel" 300 2>&1 | grep -v "^libb200"
