"""Host-side logic of the row-block sharded path on CPU:
world_size-2 (and 3) gloo process groups, the oracle standing in for the
rank-local kernel.  Checks the partition, the allgather layout (equal and
ragged blocks) and that the assembled y equals the single-process product."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, cls, use_nnz_bounds, out_dir):
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as entry
    entry.load_package()
    oracle = entry.load_oracle()
    from lilac_benchmarks_b200 import libspmv, npb, sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = npb.cg_class(cls)
        full_rowstr = None
        if use_nnz_bounds:
            full = npb.NpbMatrix(cls)
            bounds = libspmv.partition_rows(full.rowstr, world)
            full_rowstr = full.rowstr
        else:
            bounds = None
        layout = sharded.ShardLayout.build(c.na, world, bounds)
        lo, hi = layout.local_range(rank)
        blk = npb.NpbMatrix(cls, lo, hi)                 # this rank's rows only

        def local_spmv(x_full, y_local):
            y = oracle.spmv(blk.a, x_full.numpy(), blk.rowstr, blk.colidx)
            y_local.copy_(torch.from_numpy(y))

        sh = sharded.ShardedSpmv(layout, rank, local_spmv, dist=dist)
        x = np.random.default_rng(5).standard_normal(c.na)      # same on every rank
        y_local = sh.step(torch.from_numpy(x[lo:hi].copy())).clone()   # step() reuses its buffer
        # two chained products: y of step 1 is the x slice of step 2 (power iteration)
        y2_local = sh.step(y_local)
        np.save(os.path.join(out_dir, f"y1_{rank}.npy"), y_local.numpy())
        np.save(os.path.join(out_dir, f"y2_{rank}.npy"), y2_local.numpy())
        if rank == 0:
            np.save(os.path.join(out_dir, "bounds.npy"), np.asarray(layout.bounds))
            if full_rowstr is not None:
                np.save(os.path.join(out_dir, "imb.npy"),
                        np.array([sharded.block_imbalance(full_rowstr, layout.bounds)]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,use_nnz_bounds", [(2, False), (2, True), (3, True)])
def test_sharded_product_matches_single_process(tmp_path, world, use_nnz_bounds):
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as entry
    entry.build()
    entry.load_package()
    oracle = entry.load_oracle()
    from lilac_benchmarks_b200 import npb
    cls = "S"
    mp.spawn(_worker, args=(world, _free_port(), cls, use_nnz_bounds, str(tmp_path)),
             nprocs=world, join=True)
    full = npb.NpbMatrix(cls)
    x = np.random.default_rng(5).standard_normal(full.n)
    y1 = oracle.spmv(full.a, x, full.rowstr, full.colidx)
    y2 = oracle.spmv(full.a, y1, full.rowstr, full.colidx)
    got1 = np.concatenate([np.load(tmp_path / f"y1_{r}.npy") for r in range(world)])
    got2 = np.concatenate([np.load(tmp_path / f"y2_{r}.npy") for r in range(world)])
    assert np.array_equal(got1, y1)
    assert np.array_equal(got2, y2)
    bounds = np.load(tmp_path / "bounds.npy")
    assert bounds[0] == 0 and bounds[-1] == full.n
    if use_nnz_bounds:
        assert float(np.load(tmp_path / "imb.npy")[0]) < 1.02


def test_layout_helpers():
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as entry
    entry.load_package()
    from lilac_benchmarks_b200 import sharded
    lay = sharded.ShardLayout.build(10, 4)
    assert list(lay.bounds) == [0, 3, 6, 9, 10] and lay.slot == 3 and lay.contiguous
    lay = sharded.ShardLayout.build(10, 3, [0, 2, 7, 10])
    assert lay.slot == 5 and not lay.contiguous and lay.local_range(1) == (2, 7)
    lay = sharded.ShardLayout.build(1500000, 8)
    assert lay.contiguous and lay.slot == 187500


def _cg_worker(rank, world, port, cls, out_dir):
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as entry
    entry.load_package()
    oracle = entry.load_oracle()
    from lilac_benchmarks_b200 import npb, sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = npb.cg_class(cls)
        layout = sharded.ShardLayout.build(c.na, world)
        lo, hi = layout.local_range(rank)
        blk = npb.NpbMatrix(cls, lo, hi)

        def local_spmv(x_full, y_local):
            y_local.copy_(torch.from_numpy(oracle.spmv(blk.a, x_full.numpy(), blk.rowstr, blk.colidx)))

        sh = sharded.ShardedSpmv(layout, rank, local_spmv, dist=dist)
        cg = sharded.ShardedNpbCg(sh, sharded.TorchVectorOps(), c.shift)
        zeta, rnorm, _ = cg.run(c.niter)
        if rank == 0:
            np.save(os.path.join(out_dir, "zeta.npy"), np.array(zeta))
            np.save(os.path.join(out_dir, "counts.npy"), np.array([cg.spmv_count, cg.collectives]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_device_resident_cg_verifies_on_two_ranks(tmp_path):
    """The multi-rank CG driver (allgather of p + allreduce of the dot products)
    reaches NPB's zeta for class S on a world_size-2 gloo group."""
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as entry
    entry.build()
    entry.load_package()
    from lilac_benchmarks_b200 import npb
    mp.spawn(_cg_worker, args=(2, _free_port(), "S", str(tmp_path)), nprocs=2, join=True)
    c = npb.cg_class("S")
    zeta = np.load(tmp_path / "zeta.npy")
    assert len(zeta) == c.niter
    assert abs(zeta[-1] - c.zeta_verify) / c.zeta_verify <= 1e-10
    spmv, coll = np.load(tmp_path / "counts.npy")
    assert spmv == 26 * c.niter
    # per conj_grad: 26 allgathers + 1 + 25*2 + 1 allreduces, plus 1 for the norms
    assert coll == c.niter * (26 + 1 + 50 + 1 + 1)
