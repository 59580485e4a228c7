set -x
free -g | head -2
B200_SPMV_VERBOSE=1 python scripts/sweep.py D "auto,vector" 20 2>&1 | grep -v "kernel=vector panel" | tee gpurun_out/sweep20_classD_full.txt
