export B200_SPMV_VERBOSE=1
timeout 300 python scripts/sweep.py D/8 "auto" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep48.txt
timeout 300 python scripts/sweep.py D/2 "auto,pr:B=2" 20 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep48.txt
timeout 600 python scripts/sweep.py D "auto,pr:B=2" 10 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep48.txt
