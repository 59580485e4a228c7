set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
B200_SPMV_VERBOSE=1 python scripts/sweep.py C "13984x1024,13984x1024xnotma,12288x1024,8192x1024,6144x512,4096x512,ordered,vector" 100 2>&1 | tee gpurun_out/sweep3.txt
python scripts/sweep.py A "16384x1024,16384x96,16384x256,ordered,vector" 200 2>&1 | tee -a gpurun_out/sweep3.txt
python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmv_panel_kernel -s 5 -c 1 -o gpurun_out/prof_panel2 python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
