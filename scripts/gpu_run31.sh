timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "panel" 2>&1 | tail -4
export B200_SPMV_VERBOSE=1
timeout 300 python scripts/sweep.py D/8 "sell,pr,pr:Q=64,pr:Q=128,pr:K=2,pr:B=1,pr:G=2;R=1024,pr:G=8,pr:W=8192" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep31.txt
timeout 300 python scripts/sweep.py C "panel,pr:G=2,pr:G=2;Q=64,pr:G=4,pr:G=2;B=1" 100 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep31.txt
timeout 300 python scripts/sweep.py D/4 "sell,pr,pr:B=1,pr:G=4;R=2048" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep31.txt
