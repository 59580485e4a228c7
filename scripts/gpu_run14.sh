set -x
python -m pytest tests -m gpu -x -q -k "device_resident" 2>&1 | tail -8
python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench14.json 2> gpurun_out/bench14.err; python -c "
import json; d=json.load(open('gpurun_out/bench14.json')); print(json.dumps({k:d[k] for k in ('value','npb_cg','npb_cg_device_resident')}, indent=1))"; tail -3 gpurun_out/bench14.err
