for so in "" scripts/bin/b200_ld1.so scripts/bin/b200_ld2.so; do
  echo "== $so"
  if [ -n "$so" ]; then export B200_SPMV_SO=$PWD/$so; fi
  timeout 300 python scripts/sweep.py C "panel,13984x1024xg2,10240x1024xg2" 200 2>&1 | grep -v "^libb200"
  timeout 300 python scripts/sweep.py B "panel" 200 2>&1 | grep -v "^libb200"
done | tee gpurun_out/sweep42.txt
