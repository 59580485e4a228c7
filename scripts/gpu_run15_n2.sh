set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2b.json')); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','npb_cg_device_resident')}, indent=1))"; tail -5 gpurun_out/bench_n2b.err
