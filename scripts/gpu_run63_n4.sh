timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 4 --steps 200 --warmup 10 > gpurun_out/bench63_n4.json 2> gpurun_out/bench63_n4.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench63_n4.json').read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}))
print(d['roofline']['frac'], d['e2e']['ms_per_step'], d['npb_cg_device_resident']['time_s'], d['npb_cg_device_resident']['verified'], d['npb_cg_device_resident']['nccl_variant']['time_s'])
print(d['config']['kernel'], d['config']['nccl_allgather_variant_ms_per_step'])
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench63_n4.err | tail -5
