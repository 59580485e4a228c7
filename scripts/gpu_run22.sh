python scripts/sweep.py B "panel,12288x512xg1,12288x512xg2xu3,12288x512xg1xu3,12288x1024xg2,8192x512xg1,sell" 100 2>&1 | tee gpurun_out/sweep22.txt
python scripts/sweep.py A "panel,16384x128xg1,16384x64xg1,16384x192xg2,sell" 200 2>&1 | tee -a gpurun_out/sweep22.txt
python scripts/sweep.py B "panel" 5 > gpurun_out/plain22.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmv_panel_kernel -s 3 -c 1 -o gpurun_out/prof_panel_B python scripts/sweep.py B "panel" 5 > gpurun_out/ncu22.log 2>&1
tail -1 gpurun_out/ncu22.log
