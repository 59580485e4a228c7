#!/usr/bin/env bash
# One parametrised runner for everything that is sent to the GPU box
# (`gpurun -- 'bash scripts/gpu.sh <task> [args] ; bash scripts/gpu.sh <task> ...'`).
# Every task writes what it produces under gpurun_out/<tag>*; summaries worth
# keeping are copied into profiles/ by hand, named rNN_<what>.
#
#   test [pytest -k expr]                     pytest -m gpu
#   bench TAG [bench.py args]                 one-GPU bench line
#   benchn N TAG [bench.py args]              N ranks under torch.distributed.run
#   sweep TAG SHAPE "cfg,cfg,..." [iters] [graph|cloop]   scripts/sweep.py (kernel-only timing of layouts; "graph": one CUDA graph replay, "cloop": launches from a C loop)
#   launches TAG [bench.py args]              ncu launch list (gpu__time_duration) of a short bench run
#   ncufull TAG KERNEL_REGEX CMD...           one ncu --set full capture (+ raw/source CSV, stall summary)
#   e2e TAG CLASS MODE "ENV=V ENV=V|..." [calls]   scripts/e2e_probe.py (drop-in call, C caller loop, steady state) once per
#                                             "|"-separated set of environment assignments, each run under a timeout
#   smoke                                     __graft_entry__.smoke()
set -u
mkdir -p gpurun_out
task=${1:?task}; shift
PORT=${PORT:-29555}
case "$task" in
test)
    if [ $# -gt 0 ]; then timeout 1500 python -m pytest tests -m gpu -x -q -k "$1" 2>&1 | tail -15
    else timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi ;;
bench)
    tag=$1; shift
    timeout 900 python bench.py "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
    echo "rc=$?"; cut -c1-400 gpurun_out/$tag.json; grep -v '^\*\|OMP_NUM\|^$\|^libb200' gpurun_out/$tag.err | tail -5 ;;
benchn)
    n=$1; tag=$2; shift 2
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port $PORT bench.py --gpus $n "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
    echo "rc=$?"; cut -c1-400 gpurun_out/$tag.json; grep -v '^\*\|OMP_NUM\|^$\|^libb200' gpurun_out/$tag.err | tail -8 ;;
sweep)
    tag=$1; shape=$2; cfgs=$3; iters=${4:-50}; mode=${5:-}
    timeout 900 python scripts/sweep.py "$shape" "$cfgs" $iters $mode 2>&1 | grep -v '^libb200' | tee -a gpurun_out/$tag.txt ;;
launches)
    tag=$1; shift
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-npb --no-cpu "$@" \
        > gpurun_out/${tag}_ncu.log 2>&1
    echo "rc=$?"; tail -3 gpurun_out/${tag}_launches.csv ;;
ncufull)
    tag=$1; regex=$2; shift 2
    timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name "regex:$regex" \
        --launch-skip 3 --launch-count 1 -f -o gpurun_out/$tag "$@" > gpurun_out/${tag}_ncu.log 2>&1
    echo "rc=$?"
    ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
    ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/${tag}_src.csv 2>/dev/null
    python scripts/ncu_stalls.py gpurun_out/${tag}_src.csv 25 > gpurun_out/${tag}_stalls.txt 2>&1
    head -12 gpurun_out/${tag}_stalls.txt ;;
e2e)
    tag=$1; cls=$2; mode=$3; sets=${4:-}; calls=${5:-300}
    IFS='|' read -ra variants <<< "${sets:-B200_X=1}"
    for v in "${variants[@]}"; do
        echo "== $v" | tee -a gpurun_out/$tag.txt
        # shellcheck disable=SC2086
        timeout 150 env $v python scripts/e2e_probe.py "$cls" "$mode" "$calls" 2>&1 \
            | grep -v "^libb200-spmv: up\|matrix ready" | tail -3 | tee -a gpurun_out/$tag.txt
    done ;;
smoke)
    timeout 600 python -c "import __graft_entry__ as e; e.smoke()" ;;
*)
    echo "unknown task $task"; exit 2 ;;
esac
