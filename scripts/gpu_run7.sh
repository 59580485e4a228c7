set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B200_SPMV_VERBOSE=1 python scripts/sweep.py D/8 "panel,27000x1024xb1,20000x1024xb1,16384x1024xb1,vector" 30 2>&1 | grep -v "kernel=\(vector\|ordered\) panel" | tee gpurun_out/sweep7.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py B "panel,12288x512xg2,6144x512xg2,6144x256xg2,6144x256xg1,vector" 100 2>&1 | grep -v "kernel=\(vector\|ordered\) panel" | tee -a gpurun_out/sweep7.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py C "panel" 100 2>&1 | tee -a gpurun_out/sweep7.txt
