set -x
(while true; do free -g | sed -n 2p; sleep 10; done) > gpurun_out/mem_n8E.log 2>&1 &
MEMPID=$!
timeout 780 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 8 --workload E --steps 30 --warmup 3 > gpurun_out/bench_n8_E.json 2> gpurun_out/bench_n8_E.err
echo rc=$?
kill $MEMPID
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8_E.json').read().strip().splitlines()[-1]); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','config','npb_cg_device_resident')}, indent=1))"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n8_E.err | tail -12; sort -k3 -n -r gpurun_out/mem_n8E.log | head -2
