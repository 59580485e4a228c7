# one-off A/B of the drop-in path's x upload variants (scripts/e2e_probe.py), every step under a timeout
P="python scripts/e2e_probe.py"
run() { echo "== $*"; timeout ${TMO:-60} env "$@" 2>&1 | grep -v "^libb200-spmv: up\|call [0-7] done\|matrix ready" | tail -3; echo "rc=${PIPESTATUS[0]}"; }
TMO=150 run B200_X=1 $P C pageable 300
run B200_SPMV_NT_COPY=0 $P C pageable 300
run B200_X=1 $P C pageable 300
run B200_SPMV_NT_COPY=0 $P C pageable 300
run B200_X=1 $P C pinned 300
