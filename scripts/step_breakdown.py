"""Where does a sharded step's time go on N GPUs?  Under torch.distributed.run (one rank per GPU):
per rank, CUDA events around the exchange kernel and around the product of every step, for the
blocking form (exchange kernel, then product) and the overlapped form (post, then sliced
product); prints per-rank means and the per-step maximum over ranks.
usage: python -m torch.distributed.run --nproc-per-node N scripts/step_breakdown.py [D] [steps]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
entry.load_package()
from lilac_benchmarks_b200 import libspmv, npb, sharded  # noqa: E402

libspmv.lib().b200_spmv_init(local)
letter = sys.argv[1] if len(sys.argv) > 1 else "D"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cls = npb.cg_class(letter)
layout = sharded.ShardLayout.build(cls.na, world)
lo, hi = layout.local_range(rank)
dm = npb.NpbDeviceMatrix(letter, lo, hi)
rm = dm.resident()
dm.free()
x_local = torch.from_numpy(np.random.default_rng(1).random(hi - lo)).to(dev)
L = libspmv.lib()
s = torch.cuda.current_stream().cuda_stream
for name, overlap in (("blocking", False), ("overlapped", True)):
    sh = sharded.PeerShardedSpmv(libspmv, rm, layout, rank, dist=dist, device=dev, overlap=overlap, fused=False)
    n = hi - lo
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for _ in range(10):
        sh.step(x_local)
    torch.cuda.synchronize()
    dist.barrier()
    for k in range(steps):
        sh.epoch += 1
        e = sh.epoch
        ev[k][0].record()
        if overlap:
            L.b200_peer_post(sh.g, x_local.data_ptr(), n, lo, e, s)
        else:
            L.b200_peer_exchange(sh.g, x_local.data_ptr(), n, lo, e, s)
        ev[k][1].record()
        if overlap:
            rm.exec_sliced_ptr(L.b200_peer_xbuf(sh.g, e), sh.y_local.data_ptr(), s, sh.vflags, e, layout.slot, world)
        else:
            rm.exec_ptr(sh.xfull, sh.y_local.data_ptr(), s)
        ev[k][2].record()
    torch.cuda.synchronize()
    ex = np.array([ev[k][0].elapsed_time(ev[k][1]) for k in range(steps)]) * 1e3
    pr = np.array([ev[k][1].elapsed_time(ev[k][2]) for k in range(steps)]) * 1e3
    tot = ev[0][0].elapsed_time(ev[-1][2]) * 1e3 / steps
    t = torch.tensor(np.stack([ex, pr]), device=dev)
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        a = torch.stack(allt).cpu().numpy()          # [rank][0:exchange,1:product][step]
        print(f"{name}: step {tot:.1f} us (rank 0, incl. event overhead)")
        print("  exchange kernel, mean per rank:", np.round(a[:, 0, 5:].mean(axis=1), 1))
        print("  product kernel,  mean per rank:", np.round(a[:, 1, 5:].mean(axis=1), 1))
        print("  product kernel,  std  per rank:", np.round(a[:, 1, 5:].std(axis=1), 1))
        print("  per-step max over ranks of the product: mean", round(float(a[:, 1, 5:].max(axis=0).mean()), 1),
              " min over ranks:", round(float(a[:, 1, 5:].min(axis=0).mean()), 1))
    sh.close()
dist.barrier()
dist.destroy_process_group()
