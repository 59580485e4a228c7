export CUDA_LAUNCH_BLOCKING=1 B200_SPMV_VERBOSE=1
for extra in "B200_SPMV_PANEL_FMT=1" "B200_SPMV_PANEL_FMT=1 B200_SPMV_PANEL_TMA=0" "B200_SPMV_PANEL_FMT=1 B200_SPMV_PANEL_NBUF=1" "B200_SPMV_PANEL_FMT=2" "B200_SPMV_PANEL_FMT=1 B200_SPMV_PANEL_COLS=8192"; do
  echo "== $extra"; env $extra B200_SPMV_PANEL_G=8 B200_SPMV_PANEL_COLS=4096 B200_SPMV_PANEL_ROWS=4096 python scripts/debug_case.py 4500 1500000 450 2>&1 | grep -v "^$" | tail -3
done > gpurun_out/debug30.txt 2>&1
cat gpurun_out/debug30.txt
unset CUDA_LAUNCH_BLOCKING
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "flagged" 2>&1 | tail -8
B200_SPMV_VERBOSE=1 python scripts/sweep.py D/8 "sell,pr,pr:Q=64,pr:Q=128,pr:K=2,pr:B=1,pr:G=2;R=1024,pr:G=8,pg" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep30.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py C "panel,pr:G=2,pr:G=2;Q=64,pr:G=4" 100 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep30.txt
B200_SPMV_VERBOSE=1 python scripts/sweep.py D/4 "sell,pr,pr:B=1,pr:G=4;R=2048" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee -a gpurun_out/sweep30.txt
