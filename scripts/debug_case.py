"""Run one flagged-panel parity case with explicit env and report the CUDA error text."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as entry
entry.load_package()
from lilac_benchmarks_b200 import libspmv
from conftest import make_csr
import torch
oracle = entry.load_oracle()
n, ncols, mean = [int(v) for v in sys.argv[1:4]]
rng = np.random.default_rng(n * 11 + mean)
lens = rng.poisson(mean, n)
lens[rng.random(n) < 0.1] = 0
a, c, rowstr, x = make_csr(rng, n, ncols, lens, dtype=np.float64, sort=True)
y0 = oracle.spmv(a, x, rowstr, c)
m = libspmv.ResidentMatrix(a, rowstr, c, kernel="panel")
dx = torch.from_numpy(np.ascontiguousarray(x[:max(m.ncols, 1)])).cuda()
dy = torch.full((m.rows,), float("nan"), dtype=torch.float64, device="cuda")
try:
    m.exec(dx, dy)
    torch.cuda.synchronize()
    y = dy.cpu().numpy()
    bad = np.flatnonzero(y != y0)
    print("kernel", m.kernel_name, "mismatches", len(bad), bad[:10])
except Exception as e:
    print("CUDA error:", str(e).splitlines()[0])
