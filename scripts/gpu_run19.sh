set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B200_SPMV_VERBOSE=1 python scripts/sweep.py pl22 "auto,sell:C=32,sell:C=128,merge,vector" 30 2>&1 | grep -v "kernel=vector panel" | tee gpurun_out/sweep19.txt
python scripts/sweep.py crsmat170 "auto,merge" 50 2>&1 | tee -a gpurun_out/sweep19.txt
python scripts/sweep.py C "merge,vector" 50 2>&1 | tee -a gpurun_out/sweep19.txt
