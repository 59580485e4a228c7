set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B200_SPMV_VERBOSE=1 python scripts/sweep.py C "panel,sell,vector" 100 2>&1 | grep -v "^libb200" | tee gpurun_out/sweep8.txt
python scripts/sweep.py A "panel,sell,vector" 200 2>&1 | tee -a gpurun_out/sweep8.txt
python scripts/sweep.py S "panel,sell,vector" 200 2>&1 | tee -a gpurun_out/sweep8.txt
python scripts/sweep.py B "panel,sell,vector" 100 2>&1 | tee -a gpurun_out/sweep8.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py D/8 "sell,vector" 30 2>&1 | grep -v "kernel=vector panel" | tee -a gpurun_out/sweep8.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py crsmat170 "auto,sell,vector,ordered" 50 2>&1 | grep -v "kernel=\(vector\|ordered\) panel" | tee -a gpurun_out/sweep8.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py pl22 "auto,vector,ordered" 30 2>&1 | grep -v "kernel=\(vector\|ordered\) panel" | tee -a gpurun_out/sweep8.txt
