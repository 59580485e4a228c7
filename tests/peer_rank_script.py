"""One rank of the 2-GPU peer-memory tests (launched by tests/test_multi_gpu.py with
torch.distributed.run, one process per GPU).  Every rank builds its row block of an NPB
matrix, runs a few sharded products with changing x through (a) the overlapped exchange
-- b200_peer_post + the product waiting per x slice --, (b) the one-kernel exchange
followed by the product, (c) the NCCL allgather, and checks its y block ELEMENT-WISE
against the oracle's product of the whole matrix.  Then the peer-memory NPB CG."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402


def main():
    cls_letter = sys.argv[1] if len(sys.argv) > 1 else "A"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    entry.load_package()
    oracle = entry.load_oracle()
    from lilac_benchmarks_b200 import libspmv, npb, sharded
    libspmv.lib().b200_spmv_init(local)
    whole = npb.NpbMatrix(cls_letter)
    layout = sharded.ShardLayout.build(whole.n, world)
    lo, hi = layout.local_range(rank)
    # the ring layout is the kernel that can wait in-kernel for its x slices
    os.environ["B200_SPMV_PANEL_FMT"] = "2"
    rm = libspmv.ResidentMatrix(whole.a, whole.rowstr, whole.colidx, rows=hi - lo, row_lo=lo, kernel="panel")
    os.environ.pop("B200_SPMV_PANEL_FMT")
    assert rm.waits_in_kernel, rm.kernel_name
    rng = np.random.default_rng(77)          # same stream on every rank
    xs = [rng.standard_normal(whole.n) for _ in range(5)]
    refs = [oracle.spmv(whole.a, x, whole.rowstr, whole.colidx, omp=True)[lo:hi] for x in xs]
    for overlap, fused in ((True, True), (True, False), (False, False)):
        sh = sharded.PeerShardedSpmv(libspmv, rm, layout, rank, dist=dist, device=dev, overlap=overlap,
                                     fused=fused)
        assert sh.overlap == overlap and sh.fused == (fused and rm.can_push), (sh.overlap, sh.fused)
        try:
            for it in range(3):                   # several sweeps: buffers and epochs get reused
                for x, ref in zip(xs, refs):
                    y = sh.step(torch.from_numpy(x[lo:hi]).to(dev)).cpu().numpy()
                    assert np.array_equal(y, ref), f"rank {rank} overlap={overlap}: y differs"
            # back-to-back without a host sync in between: the flag protocol alone orders things
            xd = [torch.from_numpy(x[lo:hi]).to(dev) for x in xs]
            outs = []
            for it in range(20):
                outs.append(sh.step(xd[it % 5]).clone())
            torch.cuda.synchronize()
            for it, y in enumerate(outs):
                assert np.array_equal(y.cpu().numpy(), refs[it % 5]), f"rank {rank} step {it}"
        finally:
            sh.close()
    nccl = sharded.ShardedSpmv(layout, rank, lambda xf, yl: rm.exec(xf, yl), dist=dist, device=dev)
    for x, ref in zip(xs, refs):
        y = nccl.step(torch.from_numpy(x[lo:hi]).to(dev)).cpu().numpy()
        assert np.array_equal(y, ref)
    cg = sharded.PeerNpbCg(libspmv, rm, layout, rank, whole.cls.shift, dist=dist, device=dev)
    try:
        zeta, _, _ = cg.run(whole.cls.niter)
    finally:
        cg.close()
    zv = float(whole.cls.zeta_verify)
    assert abs(zeta[-1] - zv) / zv <= 1e-10, zeta[-1]
    dist.barrier()
    if rank == 0:
        print(f"peer ok world={world} class={cls_letter} zeta={zeta[-1]:.13f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
