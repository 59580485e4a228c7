set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B200_SPMV_VERBOSE=1 python scripts/sweep.py D/8 "sell,vector" 30 2>&1 | grep -v "kernel=vector panel" | tee gpurun_out/sweep9.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py crsmat170 "sell,vector" 50 2>&1 | grep -v "kernel=\(vector\|ordered\) panel" | tee -a gpurun_out/sweep9.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py pl22 "sell,vector" 30 2>&1 | grep -v "kernel=\(vector\|ordered\) panel" | tee -a gpurun_out/sweep9.txt
python scripts/sweep.py A "panel,sell,vector" 200 2>&1 | tee -a gpurun_out/sweep9.txt
