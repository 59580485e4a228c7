timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "panel" 2>&1 | tail -4
export B200_SPMV_VERBOSE=1
timeout 400 python scripts/sweep.py D/8 "sell,pr:S=2,pr:S=3,pr:S=2;B=1,pr:S=3;B=1,pr:G=8;S=2,pr:G=2;K=2;S=3!TMAX=640,pr:G=2;K=4;S=2!TMAX=640,pg" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep36.txt
timeout 300 python scripts/sweep.py C "panel,pr:G=2;S=2,pr:G=4;S=2,pg:G=2" 100 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep36.txt
timeout 300 python scripts/sweep.py D/4 "sell,pr:S=2,pr:S=3,pr:S=3;B=1,pr:G=4;K=4;S=2!TMAX=640" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep36.txt
