set -x
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err
echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2c.json').read().strip().splitlines()[-1]); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','npb_cg_device_resident')}, indent=1))"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n2c.err | tail -12
