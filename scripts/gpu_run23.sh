python scripts/sweep.py B "panel,12288x512xg2xu8,12288x512xg1,12288x256xg1,8192x256xg1" 100 2>&1 | tee gpurun_out/sweep23.txt
python scripts/sweep.py A "panel,16384x128xg1,16384x128xg2xu8,16384x256xg1" 200 2>&1 | tee -a gpurun_out/sweep23.txt
python scripts/sweep.py W "panel,sell,vector" 200 2>&1 | tee -a gpurun_out/sweep23.txt
python scripts/sweep.py C "panel" 100 2>&1 | tee -a gpurun_out/sweep23.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
