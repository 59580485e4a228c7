"""Row-block sharded SpMV across the GPUs of one box.

The reference is single-process / single-device (libspmv/gpu.c); this is new
ground required by the north star: contiguous row blocks, one per rank, each
rank keeps its block resident and owns the matching slice of every vector.
The one exchange step of the path is re-assembling the full x from the ranks'
slices before a product -- an allgather -- after which the product itself is
rank-local.  One process per GPU; `torch.distributed` (NCCL over NVLink /
NVSwitch on the GPU box, gloo in the CPU tests) is the plumbing.

The local product is injected (`local_spmv`) so the same host logic runs on a
GPU (lilac_benchmarks_b200.libspmv.ResidentMatrix.exec) and in the world_size-2
gloo tests on a CPU box.
"""
from dataclasses import dataclass

import numpy as np


def equal_row_bounds(rows, parts):
    """Equal-count contiguous row blocks (the last ones may be one row shorter).
    NPB CG rows are near-uniform (row-length CV 0.21-0.24), so equal rows are
    within ~1 % of nnz balance and allow the plain equal-count allgather."""
    per = -(-rows // parts)
    b = np.minimum(np.arange(parts + 1, dtype=np.int64) * per, rows)
    return b.astype(np.int32)


def block_imbalance(rowstr, bounds):
    """max block nnz / mean block nnz."""
    per = np.diff(rowstr[np.asarray(bounds)].astype(np.int64))
    return float(per.max() / max(per.mean(), 1.0))


@dataclass
class ShardLayout:
    rows: int            # global rows (= global length of x and y; square operator)
    parts: int
    bounds: np.ndarray   # parts+1 row boundaries
    slot: int            # padded slice length used by the equal-count allgather

    @classmethod
    def build(cls, rows, parts, bounds=None):
        bounds = equal_row_bounds(rows, parts) if bounds is None else np.asarray(bounds, dtype=np.int32)
        assert bounds[0] == 0 and bounds[-1] == rows and len(bounds) == parts + 1
        slot = int(np.diff(bounds).max()) if parts else 0
        return cls(rows=rows, parts=parts, bounds=bounds, slot=slot)

    def local_range(self, rank):
        return int(self.bounds[rank]), int(self.bounds[rank + 1])

    @property
    def contiguous(self):
        """True when the padded slots tile x without gaps (every block but the
        last is exactly `slot` rows), so the gathered buffer IS x."""
        d = np.diff(self.bounds)
        return bool(np.all(d[:-1] == self.slot)) if self.parts > 1 else True


class ShardedSpmv:
    """y_local = A[rows of this rank, :] @ allgather(x_local)."""

    def __init__(self, layout, rank, local_spmv, dist=None, device=None, dtype=None):
        import torch
        self.torch = torch
        self.layout = layout
        self.rank = rank
        self.local_spmv = local_spmv          # (x_full_tensor, y_local_tensor) -> None
        self.dist = dist
        self.lo, self.hi = layout.local_range(rank)
        dtype = dtype or torch.float64
        device = device or "cpu"
        # gathered buffer: parts slots of `slot` elements; with a contiguous
        # layout its first `rows` elements are x itself
        self.x_gather = torch.zeros(layout.parts * layout.slot, dtype=dtype, device=device)
        self.x_full = (self.x_gather if layout.contiguous
                       else torch.zeros(layout.rows, dtype=dtype, device=device))
        self.x_slot = torch.zeros(layout.slot, dtype=dtype, device=device)
        self.y_local = torch.zeros(self.hi - self.lo, dtype=dtype, device=device)

    def assemble_x(self, x_local):
        """The exchange step: allgather of the ranks' x slices."""
        n_local = self.hi - self.lo
        self.x_slot[:n_local].copy_(x_local[:n_local])
        if self.dist is None or self.layout.parts == 1:
            self.x_gather[: self.layout.slot].copy_(self.x_slot)
        else:
            self.dist.all_gather_into_tensor(self.x_gather, self.x_slot)
        if not self.layout.contiguous:
            for p in range(self.layout.parts):
                lo, hi = self.layout.local_range(p)
                self.x_full[lo:hi].copy_(self.x_gather[p * self.layout.slot: p * self.layout.slot + hi - lo])
        return self.x_full

    def step(self, x_local):
        """One sharded product: exchange, then the rank-local kernel."""
        x_full = self.assemble_x(x_local)
        self.local_spmv(x_full, self.y_local)
        return self.y_local
