timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
export B200_SPMV_VERBOSE=1
timeout 300 python scripts/sweep.py D/8 "auto,pr:B=1" 30 2>&1 | grep -v "kernel=\(vector\|ordered\|sell\) panel" | tee gpurun_out/sweep39.txt
unset B200_SPMV_VERBOSE
timeout 900 python bench.py --steps 1000 --warmup 20 > gpurun_out/bench39.json 2> gpurun_out/bench39.err; cat gpurun_out/bench39.json; tail -3 gpurun_out/bench39.err
