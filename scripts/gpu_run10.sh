set -x
python scripts/sweep.py D/8 "sell" 5 > gpurun_out/plain10.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmv_sell_kernel -s 3 -c 1 -o gpurun_out/prof_sell_d8 python scripts/sweep.py D/8 "sell" 5 > gpurun_out/ncu10.log 2>&1
tail -2 gpurun_out/ncu10.log
