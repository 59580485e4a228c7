"""Where does the exchange time of a sharded step go?  One GPU, a group of one rank: the
1/8 row block of NPB class D (the N = 8 shape), x pushed into the rank's own buffer.
Times the product alone, the post kernel alone, post + sliced product, and the older
one-kernel exchange + product.  usage: python scripts/peer_overhead_probe.py [D] [8]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
from lilac_benchmarks_b200 import libspmv, npb, sharded  # noqa: E402

letter = sys.argv[1] if len(sys.argv) > 1 else "D"
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cls = npb.cg_class(letter)
dm = npb.NpbDeviceMatrix(letter, 0, cls.na // parts)
rm = dm.resident()
dm.free()
print("kernel", rm.kernel_name, "rows", rm.rows, "ncols", rm.ncols, "waits_in_kernel", rm.waits_in_kernel)
L = libspmv.lib()
layout = sharded.ShardLayout.build(cls.na, 1)
rng = np.random.default_rng(0)
x_local = torch.from_numpy(rng.random(cls.na)).cuda()
y = torch.zeros(cls.na, dtype=torch.float64, device="cuda")


def timed(fn, iters=50, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


s = torch.cuda.current_stream().cuda_stream
print(f"product alone           {timed(lambda i: rm.exec(x_local, y)):8.1f} us")
for overlap, fused in ((True, True), (True, False), (False, False)):
    sh = sharded.PeerShardedSpmv(libspmv, rm, layout, 0, overlap=overlap, fused=fused)
    print("fused", sh.fused)
    sh.y_local = y
    state = {"e": 0}

    def post_only(i):
        state["e"] += 1
        if overlap:
            L.b200_peer_post(sh.g, x_local.data_ptr(), cls.na, 0, state["e"], s)
        else:
            L.b200_peer_exchange(sh.g, x_local.data_ptr(), cls.na, 0, state["e"], s)
    print(f"overlap={overlap}: exchange kernel alone {timed(post_only):8.1f} us")
    sh.epoch = state["e"]
    print(f"overlap={overlap}: step                  {timed(lambda i: sh.step(x_local)):8.1f} us")
    xb = L.b200_peer_xbuf(sh.g, 0)
    print(f"overlap={overlap}: product on the peer buffer alone {timed(lambda i: rm.exec_ptr(xb, y.data_ptr(), s)):8.1f} us")
    sh.close()
