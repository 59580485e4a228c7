set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "flagged" 2>&1 | tail -5
B200_SPMV_VERBOSE=1 python scripts/sweep.py D/8 "sell,pg,pg:B=1,pg:G=8,pg:W=8192,pg:G=2;R=1024" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep28.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py C "panel,pg:G=2,pg:G=4" 100 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep28.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py D/4 "sell,pg,pg:B=1,pg:G=4" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep28.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py D/2 "sell,pg" 20 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep28.txt
