timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 1000 --warmup 20 > gpurun_out/bench67.json 2> gpurun_out/bench67.err; cut -c1-160 gpurun_out/bench67.json
