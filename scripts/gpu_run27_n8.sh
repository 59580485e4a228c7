timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err
echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8b.json').read().strip().splitlines()[-1]); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','roofline','e2e','npb_cg_device_resident')}, indent=1)); print(d['config']['nccl_allgather_variant_ms_per_step'])"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n8b.err | tail -8
