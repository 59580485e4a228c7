P="python scripts/e2e_probe.py"
run() { echo "== $*"; timeout ${TMO:-45} env "$@" 2>&1 | grep -v "^libb200\|call [1-7] done" | tail -4; echo "rc=${PIPESTATUS[0]}"; }
TMO=150 run B200_X=1 $P C pinned 100
run B200_X=1 $P C pageable 100
run B200_SPMV_X_CHUNKS=1 $P C pageable 100
run B200_SPMV_FLAG_WRITE=0 $P C pageable 100
run B200_SPMV_X_PRELAUNCH=1 $P C pageable 100
run B200_SPMV_X_PRELAUNCH=1 $P C pinned 100
run B200_SPMV_X_OVERLAP=0 $P C pageable 100
run B200_SPMV_X_OVERLAP=0 $P C pinned 100
run B200_SPMV_PANEL_ROWS=512 $P C pageable 100
run B200_X=1 $P B pageable 100
