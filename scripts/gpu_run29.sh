set -x
./scripts/bin/l2probe 12 96 20 2>&1 | tee gpurun_out/l2probe29.txt
./scripts/bin/l2probe 12 64 20 2>&1 | tee -a gpurun_out/l2probe29.txt
./scripts/bin/l2probe 1 96 50 2>&1 | tee -a gpurun_out/l2probe29.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py C "panel,pg:G=2,pg:G=2;S=227,pg:G=2;S=180,pg:G=4" 100 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep29.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py D/8 "pg,pg:S=180,pg:S=160,pg:B=1;S=180,pg:G=2;R=1280,pg:G=8" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep29.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py D/4 "pg,pg:S=180,pg:G=4;R=2560" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep29.txt
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "flagged and panel_env4-shape6-float64" 2>&1 | grep -v "^$" | head -60 > gpurun_out/sanitizer29.txt
tail -5 gpurun_out/sanitizer29.txt
