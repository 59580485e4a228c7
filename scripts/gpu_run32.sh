export B200_SPMV_VERBOSE=1
timeout 300 python scripts/sweep.py C "panel!PF=0,panel!PF=1,panel!PF=2,panel!PF=3,pr:G=2,pr:G=2!PF=16,pr:G=2!PF=0" 100 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep32.txt
timeout 300 python scripts/sweep.py D/8 "sell,pr!PF=0,pr,pr!PF=16,pr!PF=32,pr:B=1,pr:B=1!PF=16,pr:K=2!PF=16,pr:W=8192!PF=16,pr:G=8!PF=16" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep32.txt
timeout 300 python scripts/sweep.py D/4 "pr,pr!PF=16,pr:B=1!PF=16" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep32.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "panel" 2>&1 | tail -4
