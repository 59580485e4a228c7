set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "
import __graft_entry__ as e; e.smoke()"
timeout 900 python bench.py --steps 1000 --warmup 20 > gpurun_out/bench64.json 2> gpurun_out/bench64.err; cut -c1-200 gpurun_out/bench64.json; tail -2 gpurun_out/bench64.err
