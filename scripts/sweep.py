"""Kernel-only timing sweep over panel geometry / kernel family on one NPB class.
usage: python scripts/sweep.py C "16384x1024,8192x512,ordered,vector" [iters] [graph]
"graph": the launches are captured in one CUDA graph and the replay is timed -- for kernels of a
few microseconds, which a Python loop of ctypes calls (~10 us per launch) cannot issue fast enough.
"cloop": the launches are issued by the C loop of callers/npb (npb_issue_exec_calls): plain stream
launches as a compiled device-resident caller makes them, dependent launches included."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
from lilac_benchmarks_b200 import libspmv, npb  # noqa: E402

cls = sys.argv[1] if len(sys.argv) > 1 else "C"
configs = (sys.argv[2] if len(sys.argv) > 2 else "16384x1024,ordered,vector").split(",")
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100
use_graph = len(sys.argv) > 4 and sys.argv[4] == "graph"
use_cloop = len(sys.argv) > 4 and sys.argv[4] == "cloop"      # launches issued by callers/npb's C loop
if cls.startswith("crsmat"):
    from lilac_benchmarks_b200 import gen

    class _M:
        pass
    m = _M()
    m.a, m.colidx, m.rowstr, m.n = gen.crsmat(int(cls[6:]))
    m.nnz = len(m.a)
elif cls.startswith("pl"):
    from lilac_benchmarks_b200 import gen

    class _M:
        pass
    m = _M()
    m.a, m.colidx, m.rowstr, _x0 = gen.powerlaw_graph(1 << int(cls[2:]))
    m.n = len(m.rowstr) - 1
    m.nnz = len(m.a)
elif "/" in cls:                       # "D/8": first 1/8 of the rows of class D (one rank's block)
    letter, parts = cls.split("/")
    na = npb.cg_class(letter).na
    m = npb.NpbDeviceMatrix(letter, 0, na // int(parts), release_vectors=False)   # assembled on the GPU
else:
    m = npb.NpbDeviceMatrix(cls, release_vectors=False)
on_device = isinstance(m, npb.NpbDeviceMatrix)
ncols = npb.cg_class(cls.split("/")[0]).na if on_device else int(m.colidx.max())
rng = np.random.default_rng(0)
xs = [torch.from_numpy(rng.random(ncols + 2)).cuda() for _ in range(4)]
y = torch.zeros(m.n, dtype=torch.float64, device="cuda")
B = 12 * m.nnz + 4 * (m.n + 1) + 8 * ncols + 8 * m.n
y_ref = None
for cfg_full in configs:
    # "cfg!NAME=V!NAME2=V2" adds B200_SPMV_PANEL_<NAME>=V (B200_SPMV_<NAME>=V for SMALL_* names) to
    # the environment of that upload
    cfg, *extras = cfg_full.split("!")
    env = {}
    for kv in extras:
        k, v = kv.split("=")
        env[("B200_SPMV_" if k.startswith("SMALL_") else "B200_SPMV_PANEL_") + k] = v
    kernel = cfg
    if cfg.startswith("sell:"):
        kernel = "sell"
        for kv in cfg[5:].split(";"):
            k, v = kv.split("=")
            env["B200_SPMV_SELL_" + {"R": "ROWS", "G": "G", "U": "U", "C": "CAP", "F": "FMT"}[k]] = v
    elif cfg.startswith("pr"):
        # ring panel for wide matrices: "pr:R=1280;G=4;W=12288;B=2;K=4;S=2"
        kernel = "panel"
        env["B200_SPMV_PANEL_FMT"] = "2"
        for kv in cfg[3:].split(";"):
            if not kv:
                continue
            k, v = kv.split("=")
            env["B200_SPMV_PANEL_" + {"R": "ROWS", "G": "G", "W": "COLS", "B": "NBUF", "T": "TMA", "K": "RING_K", "Q": "RING_KB", "S": "RING_S", "M": "SMEM_KB"}[k]] = v
    elif "x" in cfg and cfg[0].isdigit():
        parts = cfg.split("x")
        env.update({"B200_SPMV_PANEL_COLS": parts[0], "B200_SPMV_PANEL_ROWS": parts[1]})
        for opt in parts[2:]:
            if opt == "notma":
                env["B200_SPMV_PANEL_TMA"] = "0"
            elif opt.startswith("g"):
                env["B200_SPMV_PANEL_G"] = opt[1:]
            elif opt.startswith("u"):
                env["B200_SPMV_PANEL_U"] = opt[1:]
            elif opt.startswith("b"):
                env["B200_SPMV_PANEL_NBUF"] = opt[1:]
        kernel = "panel"
    os.environ.update(env)
    rm = m.resident(kernel=kernel) if on_device else libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx, kernel=kernel)
    for i in range(10):
        rm.exec(xs[i & 3], y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if use_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                rm.exec(xs[i & 3], y)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    elif use_cloop:
        stream = torch.cuda.current_stream().cuda_stream
        npb.issue_exec_calls(libspmv.exec_address(), rm.handle, [v.data_ptr() for v in xs], y.data_ptr(), stream, iters)
        torch.cuda.synchronize()
        e0.record()
        npb.issue_exec_calls(libspmv.exec_address(), rm.handle, [v.data_ptr() for v in xs], y.data_ptr(), stream, iters)
        e1.record()
    else:
        e0.record()
        for i in range(iters):
            rm.exec(xs[i & 3], y)
        e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    rm.exec(xs[0], y)
    torch.cuda.synchronize()
    yy = y.cpu().numpy()
    if y_ref is None:
        y_ref = yy.copy()
    same = bool(np.array_equal(yy, y_ref))
    print(f"{cls} {cfg_full:>26s} kernel={rm.kernel_name:8s} {us:9.1f} us  {B / us / 1e3:8.1f} GB/s  "
          f"frac={B / us / 1e3 / 6533.5:5.3f}  same_as_first={same}", flush=True)
    rm.release()
    for k in env:
        os.environ.pop(k, None)
