export B200_SPMV_VERBOSE=1
timeout 300 python scripts/sweep.py D/8 "pr:S=2" 10 > gpurun_out/plain35.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_panelr_kernel -s 5 -c 1 -o gpurun_out/prof_ring35 python scripts/sweep.py D/8 "pr:S=2" 10 > gpurun_out/ncu35.log 2>&1
tail -2 gpurun_out/ncu35.log
timeout 300 python scripts/sweep.py D/2 "sell,pr:S=2" 20 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep35.txt
timeout 600 python scripts/sweep.py D "sell,pr:S=2" 10 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep35.txt
