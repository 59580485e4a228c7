timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cg or device or sharded" 2>&1 | tail -3
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/bench61.json 2> gpurun_out/bench61.err; python -c "
import json; d=json.loads(open('gpurun_out/bench61.json').read().strip().splitlines()[-1]); print(d['value'], d['npb_cg_device_resident'])"; tail -2 gpurun_out/bench61.err
