export B200_SPMV_VERBOSE=1
timeout 300 python scripts/sweep.py C "panel,11552x1024xb1,8192x1024xb1,16384x1024xb1,13984x1024xb1,6144x1024xb1" 200 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep43.txt
timeout 300 python scripts/sweep.py B "panel,10720x512xb1,16384x512xb1" 200 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep43.txt
