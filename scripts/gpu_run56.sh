timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ring" 2>&1 | tail -3
for so in "" scripts/bin/b200_xc64.so; do
  echo "== $so"
  if [ -n "$so" ]; then export B200_SPMV_SO=$PWD/$so; fi
  timeout 300 python scripts/sweep.py D/8 "pr,pr:B=2" 30 2>&1 | grep -v "^libb200"
  timeout 300 python scripts/sweep.py D/4 "pr" 30 2>&1 | grep -v "^libb200"
done | tee gpurun_out/sweep56.txt
