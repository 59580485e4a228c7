set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
B200_SPMV_VERBOSE=1 python scripts/sweep.py C "12288x1024xg2,12288x1024xg2xu5,12288x1024xg2xu3,13984x1024xg2,10240x1024xg2,8192x1024xg2" 100 2>&1 | tee gpurun_out/sweep5.txt
python scripts/sweep.py A "16384x128xg2,16384x128xg1,16384x256xg2,vector" 200 2>&1 | tee -a gpurun_out/sweep5.txt
python scripts/sweep.py B "12288x1024xg2,12288x512xg2,12288x512xg1,vector" 100 2>&1 | tee -a gpurun_out/sweep5.txt
python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmv_panel_kernel -s 5 -c 1 -o gpurun_out/prof_panel4 python bench.py --steps 20 --warmup 3 --no-npb --no-cpu > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
