"""Time the drop-in ABI call (host vectors) on one NPB class; library knobs come from the env.
usage: [B200_SPMV_ZEROCOPY=0|1] [B200_SPMV_PIN_HOST=1] python scripts/e2e_probe.py C pinned|pageable|registered [iters]
(registered: pageable numpy vectors handed to b200_spmv_pin_host, i.e. cudaHostRegister)"""
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

cls = sys.argv[1]
mode = sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 300
m = npb.NpbMatrix(cls)
rng = np.random.default_rng(0)
if mode == "pinned":
    hx = [torch.from_numpy(rng.random(m.n + 2)).pin_memory().numpy() for _ in range(4)]
    hy = torch.zeros(m.n, dtype=torch.float64).pin_memory().numpy()
elif mode == "registered":
    # one pageable buffer, registered whole (cudaHostRegister), the five vectors carved out of it
    per = (m.n + 2 + 511) // 512 * 512
    buf = np.zeros(5 * per + 1024)
    lo = (buf.ctypes.data + 4095) // 4096 * 4096
    first = (lo - buf.ctypes.data) // 8
    nbytes = (5 * per * 8) // 4096 * 4096
    rc = libspmv.lib().b200_spmv_pin_host(lo, nbytes)
    print(f"registered {nbytes} bytes: rc={rc}", file=sys.stderr, flush=True)
    hx = [buf[first + k * per:first + k * per + m.n + 2] for k in range(4)]
    for v in hx:
        v[:] = rng.random(m.n + 2)
    hy = buf[first + 4 * per:first + 4 * per + m.n]
else:
    hx = [rng.random(m.n + 2) for _ in range(4)]
    hy = np.zeros(m.n)
print(f"{cls} {mode}: matrix ready", file=sys.stderr, flush=True)
addr = libspmv.harness_address()
# steady state first: a caller in the middle of a solve (clocks and the PCIe link ramp up over
# tens of milliseconds of traffic), then the timed calls, issued by the C caller loop
t_warm = time.perf_counter()
while time.perf_counter() - t_warm < 0.3:
    npb.time_spmv_calls(addr, hy, m.a, hx, m.rowstr, m.colidx, m.n, 100)
if os.environ.get('PROBE_TIME_KERNELS'):
    libspmv.lib().b200_spmv_set_time_kernels(1)
res = []
for rep in range(3):
    libspmv.reset_stats()
    res.append(npb.time_spmv_calls(addr, hy, m.a, hx, m.rowstr, m.colidx, m.n, iters) * 1e6)
st = libspmv.stats()
knobs = {k: v for k, v in os.environ.items() if k.startswith("B200_SPMV")}
print(f"{cls} {mode:9s} {knobs}  e2e (C loop, 3 x {iters} calls) " + " ".join(f"{r:7.1f}" for r in res) +
      f" us/call  kernel {st['kernel_ms'] / iters * 1e3:7.1f} us  overlapped {st['x_overlapped_calls']} "
      f"timeouts {st['x_overlap_timeouts']}", flush=True)
