#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (BASELINE.json metric).

A step is ONE CSR SpMV y = A x over the named matrix.

  N = 1   default workload = NPB3.3.1 CG class C (BASELINE config 2): na=150000,
          nnz=36 121 058, the exact matrix of cg.f's makea.  436 MB of algorithmic
          traffic per step, larger than the 126 MB L2, so no L2 flush is needed.
          --workload crsmat170u  SparseBench big_gen.py matrix at 170^3 (config 3)
          --workload pl22        pagerank power-law graph, 2^22 vertices (config 4)
          --workload S|W|A|B|D   other NPB classes
  N > 1   default workload = NPB CG class D (config 5; E with --workload E at 8 GPUs),
          equal row blocks, one per rank (one process per GPU), every block assembled ON
          its GPU (include/b200_npb.h).  A step pushes the rank's x slice into every
          rank's buffer over NVLink peer memory and runs the rank-local kernel, which
          waits per slice.  Strong scaling; the one-GPU time of the SAME matrix is
          measured in the same run (rank 0), not quoted.

`value`  device-timed (CUDA events on the launching stream), operands resident in HBM,
         whole-job algorithmic GB/s (12 nnz + 4 (n+1) + 8 ncols + 8 n bytes per
         product).
`e2e`    the same metric through the drop-in C-ABI symbol `spmv_harness_` with HOST
         vectors (x host->device and y device->host inside the timed region, wall
         clock, calls issued by a C caller loop).  At N > 1 it is the SAME symbol in
         one process driving all N GPUs (ABI mode, B200_SPMV_DEVICES), run by rank 0.
--impl reference  times the reference's own CPU implementation of the path
         (oracle/_ref/native.so = libspmv/native.c built from the reference tree; else
         the oracle port) on the host cores, same matrix, same `config`.

`config` holds only what names the workload and is identical in both arms; everything
about HOW this arm ran it is under `details`.  Only the cpu_baseline / parity /
--impl reference legs touch oracle/.
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

METRIC = "spmv_algorithmic_bandwidth"
UNIT = "GB/s"
L2_BYTES = 126e6
PRECONDITION = 50       # untimed steps before every timed region, at least (reported in details)


def algorithmic_bytes(nnz, rows, ncols, es=8):
    return (es + 4) * nnz + 4 * (rows + 1) + es * ncols + es * rows


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(workload, kernel, world=1):
    """DRAM bytes per launch from the committed ncu --set full capture of this command, if any."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            key = f"{workload}:{kernel}" + (f"@{world}" if world > 1 else "")
            return json.loads(p.read_text()).get(key)
        except Exception:
            return None
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    collect = True

    def _sample(self):
        if not self.collect:
            return
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
                 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}
        for bit, name in names.items():
            if r & bit:
                self.reasons.add(name)

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self._sample()
            except Exception:
                break
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if self.ok and not self.samples:
            try:
                self.collect = True
                self._sample()
            except Exception:
                pass
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def stage(msg):
    """Progress on stderr (the one JSON line owns stdout): says where a run that was cut off stood."""
    if os.environ.get("B200_BENCH_QUIET"):
        return
    print(f"bench.py [{time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------
# workloads: the same host matrix for both arms
# --------------------------------------------------------------------------
class HostMatrix:
    def __init__(self, a, rowstr, colidx, label, kind, npb_class=None, x0=None):
        self.a, self.rowstr, self.colidx = a, rowstr, colidx
        self.n = len(rowstr) - 1
        self.nnz = int(len(a))
        self.label, self.kind, self.npb_class, self.x0 = label, kind, npb_class, x0


def workload_label(name):
    if name in ("S", "W", "A", "B", "C", "D", "E"):
        return f"npb-cg-class-{name}"
    if name.startswith("crsmat"):
        return f"sparsebench-{name}"
    if name.startswith("pl"):
        return f"pagerank-powerlaw-2^{name[2:]}"
    raise SystemExit(f"bench.py: unknown workload {name!r}")


def load_host_matrix(name):
    """The named matrix as host CSR arrays (1-based, as every caller of the ABI holds it)."""
    from lilac_benchmarks_b200 import gen, npb
    label = workload_label(name)
    if name in ("S", "W", "A", "B", "C", "D"):
        m = npb.NpbMatrix(name)
        return HostMatrix(m.a, m.rowstr, m.colidx, label, "npb", npb_class=m.cls), m
    if name.startswith("crsmat"):
        size = int(name[6:].rstrip("u") or 170)
        a, colidx, rowstr, _ = gen.crsmat(size)
        return HostMatrix(a, rowstr, colidx, label, "crsmat"), None
    if name.startswith("pl"):
        a, colidx, rowstr, x0 = gen.powerlaw_graph(1 << int(name[2:]))
        return HostMatrix(a, rowstr, colidx, label, "powerlaw", x0=x0), None
    raise SystemExit(f"bench.py: workload {name!r} cannot be held as one host matrix")


def workload_config(label, rows, nnz, ncols, world):
    """What names the workload -- identical in the b200 and the reference arm."""
    B = algorithmic_bytes(nnz, rows, ncols)
    return {"workload": label, "rows": int(rows), "nnz": int(nnz), "ncols": int(ncols),
            "algorithmic_bytes_per_step": int(B),
            "l2_policy": "inputs larger than L2 (no flush)" if B > L2_BYTES * 1.5 * world
            else "matrix comparable to the 126 MB L2: HBM fraction may read > 1"}


# --------------------------------------------------------------------------
# CPU legs (the only users of oracle/)
# --------------------------------------------------------------------------
def cpu_reference_spmv(a, rowstr, colidx, x, steps, warmup, omp=False):
    """Time the reference's CPU path.  Returns (seconds per product, kind, cores, last y)."""
    oracle = entry.load_oracle()
    use_ref = oracle.ref_available() and not omp
    kind = "reference" if use_ref else "port"
    y = None
    for _ in range(warmup):
        y = oracle.spmv(a, x, rowstr, colidx, omp=omp, use_ref=use_ref)
    t0 = time.perf_counter()
    for _ in range(steps):
        y = oracle.spmv(a, x, rowstr, colidx, omp=omp, use_ref=use_ref)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, kind, (host_cores() if omp else 1), y


def run_reference_arm(args, world, rank):
    """--impl reference: the reference's own CPU implementation on the same matrix, rank 0 only."""
    if rank != 0:
        return
    if world > 1:
        os.environ["OMP_NUM_THREADS"] = str(host_cores())      # torchrun pins it to 1; rank 0 runs alone
    entry.load_package()
    name = args.workload or ("C" if args.gpus == 1 else "D")
    if name == "E":
        print(json.dumps({"impl": "reference", "unavailable":
                          "NPB class E (6.3e9 nonzeros) exceeds the int32 ABI of the reference's single-process path"}))
        return
    t0 = time.perf_counter()
    hm, _ = load_host_matrix(name)
    t_gen = time.perf_counter() - t0
    ncols = int(hm.colidx.max())
    x = np.random.default_rng(1234).random(ncols + 2)
    B = algorithmic_bytes(hm.nnz, hm.n, ncols)
    # bound the run to about a minute of products: time one first (class D gathers from a 12 MB x
    # run at 1-3 GB/s on one core, class C at 7-10 GB/s), then size the sample
    oracle = entry.load_oracle()
    use_ref = oracle.ref_available()
    t1 = time.perf_counter()
    oracle.spmv(hm.a, x, hm.rowstr, hm.colidx, use_ref=use_ref)
    est = max(time.perf_counter() - t1, 1e-4)
    steps = max(1, min(max(args.steps, 1), int(40.0 / est) or 1, 200))
    warmup = max(1, min(max(args.warmup, 1), int(10.0 / est) or 1))     # the sizing product is the first
    dt, kind, cores, _ = cpu_reference_spmv(hm.a, hm.rowstr, hm.colidx, x, steps, warmup - 1)
    dt_omp, _, cores_omp, _ = cpu_reference_spmv(hm.a, hm.rowstr, hm.colidx, x, max(1, min(steps, 50)), 1, omp=True)
    val = B / dt / 1e9
    sample = f"whole {hm.label} matrix, {steps} products of {dt * 1e3:.1f} ms"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gflops": 2.0 * hm.nnz / dt / 1e9,
        "config": workload_config(hm.label, hm.n, hm.nnz, ncols, args.gpus),
        "details": {"implementation": "libspmv/native.c (sequential by construction)" if kind == "reference"
                    else "oracle port of libspmv/native-impl.c", "gen_s": round(t_gen, 2)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "cpu_baseline_omp": {"value": B / dt_omp / 1e9, "unit": UNIT, "cores": cores_omp, "kind": "port",
                             "note": "row-parallel OpenMP loop, stand-in for libspmv/mkl.c (MKL not in image)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# B200 arm, one GPU
# --------------------------------------------------------------------------
def time_steps(torch, step, K, W, barrier, local_rank, bulk=None):
    """W untimed steps, then exactly K timed ones between two CUDA events on the launching
    stream.  The clock sampler is created and started BEFORE the warm-up (initialising NVML
    takes tens of milliseconds of host time: with the GPU idle that long right before the timed
    region the first timed steps ran ~10 % slow) and only collects during the timed region."""
    sampler = ClockSampler(local_rank)
    sampler.collect = False
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # a short requested warm-up (the driver passes 5) is topped up to PRECONDITION untimed steps:
    # a few hundred microseconds of work do not bring a GPU that has just idled back to its
    # steady state (class C: 75.7 us per step over 20 steps after 5 warm-ups, 74.1 us after 50)
    if bulk is not None:
        bulk(max(W, PRECONDITION))
    else:
        for i in range(max(W, PRECONDITION)):
            step(i)
    barrier()
    sampler.collect = True
    e0.record()
    if bulk is not None:
        bulk(K)                     # exactly K launches, issued by the C caller loop
    else:
        for i in range(K):
            step(i)
    e1.record()
    barrier()
    sampler.collect = False
    return e0.elapsed_time(e1), sampler.stop()


def run_one_gpu(args):
    import torch
    entry.load_package()
    from lilac_benchmarks_b200 import callers, libspmv, npb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 platform has no CPU fallback")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    libspmv.lib().b200_spmv_init(0)
    if args.abi_devices > 1:
        # the end-to-end legs (and the callers) go through spmv_harness_ spread over this many
        # GPUs by one process (ABI mode); `value` stays the one-GPU kernel
        ndev = min(args.abi_devices, torch.cuda.device_count())
        os.environ["B200_SPMV_DEVICES"] = ",".join(str(d) for d in range(ndev))
    K, W = max(args.steps, 1), max(args.warmup, 3)
    rng = np.random.default_rng(1234)
    name = args.workload or "C"
    t0 = time.perf_counter()
    hm, npb_m = load_host_matrix(name)
    t_gen = time.perf_counter() - t0
    stage(f"{hm.label}: host matrix ready ({t_gen:.1f} s)")
    t0 = time.perf_counter()
    rm = libspmv.ResidentMatrix(hm.a, hm.rowstr, hm.colidx, kernel=args.kernel)
    t_upload = time.perf_counter() - t0
    stage(f"resident on the GPU ({rm.kernel_name}, {t_upload:.2f} s)")
    ncols = rm.ncols
    B = algorithmic_bytes(hm.nnz, hm.n, ncols)
    xs = [torch.from_numpy(rng.random(ncols + 2)).to(dev) for _ in range(4)]
    y = torch.zeros(hm.n, dtype=torch.float64, device=dev)

    def step(i):
        rm.exec(xs[i & 3], y)

    # the launches of the timed region come from the C loop of callers/npb (what a compiled
    # device-resident caller does): a Python loop of ctypes calls cannot issue them faster than
    # one per ~10 us, which would be the number measured for the small classes
    x_ptrs = [v.data_ptr() for v in xs]
    stream0 = torch.cuda.current_stream().cuda_stream

    def bulk(count):
        npb.issue_exec_calls(libspmv.exec_address(), rm.handle, x_ptrs, y.data_ptr(), stream0, count)

    def barrier():
        torch.cuda.synchronize()

    # ---- end to end through the ABI with host vectors ---------------------
    Ke = min(K, 2000)
    hx = [torch.from_numpy(rng.random(ncols + 2)).pin_memory() for _ in range(4)]
    hy = torch.zeros(hm.n, dtype=torch.float64).pin_memory()
    hx_np, hy_np = [t.numpy() for t in hx], hy.numpy()
    # the calls are issued by the C caller loop of callers/npb (what a compiled caller of the
    # ABI such as cg.f pays per call; a ctypes call from Python adds ~15-30 us of argument
    # marshalling that is not the library's)
    addr = libspmv.harness_address()

    def precondition_e2e(y_np, x_list, seconds=0.25):
        """Untimed calls until the GPU, its copy engines and the PCIe link are in the state a caller
        in the middle of a solve sees (they ramp up over tens of milliseconds of traffic: the
        first few hundred calls after an idle period ran 20-30 us slower)."""
        npb.time_spmv_calls(addr, y_np, hm.a, x_list, hm.rowstr, hm.colidx, hm.n, 3)   # the first call uploads
        t_start, n = time.perf_counter(), 3
        while time.perf_counter() - t_start < seconds:
            npb.time_spmv_calls(addr, y_np, hm.a, x_list, hm.rowstr, hm.colidx, hm.n, 50)
            n += 50
        return n

    e2e_untimed = precondition_e2e(hy_np, hx_np)
    libspmv.reset_stats()
    e2e_sec = npb.time_spmv_calls(addr, hy_np, hm.a, hx_np, hm.rowstr, hm.colidx, hm.n, Ke)
    st = libspmv.stats()
    h2d, d2h = st["h2d_bytes"] // Ke, st["d2h_bytes"] // Ke
    e2e_overlapped = st["x_overlapped_calls"]
    # where the time of such a call goes: the same calls once more with the library's CUDA-event
    # timing of the product switched on (off by default: it costs ~3 us per call)
    libspmv.lib().b200_spmv_set_time_kernels(1)
    libspmv.reset_stats()
    Kt = min(Ke, 100)
    npb.time_spmv_calls(addr, hy_np, hm.a, hx_np, hm.rowstr, hm.colidx, hm.n, Kt)
    e2e_kernel_ms = libspmv.stats()["kernel_ms"] / Kt
    libspmv.lib().b200_spmv_set_time_kernels(0)
    e2e_devices = libspmv.devices_in_use()
    e2e_y = hy_np.copy()                       # result of the last call: x = hx_np[(Ke - 1) & 3]
    stage(f"e2e, pinned caller vectors: {e2e_sec * 1e6:.1f} us per call")
    # pageable caller vectors (what NPB's COMMON arrays / pagerank's std::vectors are): pinned
    # bounce buffers inside the library, filled by its copy threads
    px = [np.array(v) for v in hx_np]
    py = np.zeros(hm.n)
    precondition_e2e(py, px, 0.1)
    e2e_pageable_sec = npb.time_spmv_calls(addr, py, hm.a, px, hm.rowstr, hm.colidx, hm.n, min(Ke, 500))
    stage(f"e2e, pageable caller vectors: {e2e_pageable_sec * 1e6:.1f} us per call")
    # ... and the same pageable vectors with B200_SPMV_PIN_HOST=3 (opt-in): the library registers a
    # vector on its third sighting and checks the mapping on every call
    libspmv.lib().b200_spmv_set_auto_pin(3)
    npb.time_spmv_calls(addr, py, hm.a, px, hm.rowstr, hm.colidx, hm.n, 16)
    e2e_autopin_sec = npb.time_spmv_calls(addr, py, hm.a, px, hm.rowstr, hm.colidx, hm.n, min(Ke, 500))
    stage(f"e2e, pageable vectors registered on third sight: {e2e_autopin_sec * 1e6:.1f} us per call")
    libspmv.lib().b200_spmv_set_auto_pin(0)      # gives the registrations back ...
    libspmv.lib().b200_spmv_set_auto_pin(int(os.environ.get("B200_SPMV_PIN_HOST", "0") or 0))   # ... and restores the setting
    t0 = time.perf_counter()
    for i in range(min(Ke, 200)):
        libspmv.spmv_harness(hy_np, hm.a, hx_np[i & 3], hm.rowstr, hm.colidx, hm.n)
    e2e_python_sec = (time.perf_counter() - t0) / min(Ke, 200)

    # ---- the headline: device-timed, after the end-to-end legs (clocks, TLB and L2 are in
    # the state a caller in the middle of a solve sees, not the state right after the upload)
    stage("e2e legs done; device-timed steps")
    ms, clocks = time_steps(torch, step, K, W, barrier, 0, bulk=bulk)
    sec_per_step = ms / 1e3 / K
    value = B / sec_per_step / 1e9
    stage(f"kernel: {sec_per_step * 1e6:.1f} us per step")

    peak, peak_src = measured_peak()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gflops": 2.0 * hm.nnz / sec_per_step / 1e9,
        "gpu_launches": K * rm.launches_per_exec,
        "config": workload_config(hm.label, hm.n, hm.nnz, ncols, 1),
        "details": {"kernel": rm.kernel_name, "launches_per_step": rm.launches_per_exec,
                    "launches_issued_by": "C caller loop (callers/npb, npb_issue_exec_calls)",
                    "untimed_steps_before_timing": max(W, PRECONDITION),
                    "x_vectors_rotated": 4, "resident_bytes": rm.resident_bytes,
                    "gen_s": round(t_gen, 2), "upload_s": round(t_upload, 3)},
        "roofline": {"bound": "hbm", "achieved": value, "peak": peak, "unit": "GB/s",
                     "frac": value / peak, "traffic": recorded_traffic(hm.label, rm.kernel_name),
                     "traffic_source": "profiles/roofline_traffic.json (ncu --set full capture of this command)",
                     "peak_source": peak_src, "kernel": f"spmv ({rm.kernel_name})"},
        "e2e": {"value": B / e2e_sec / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_sec * 1e3, "steps": Ke,
                "untimed_calls_before_timing": e2e_untimed,
                "api": "spmv_harness_ called from the C caller loop (callers/npb), pinned caller vectors"
                       + (f", one process driving {e2e_devices} GPUs (B200_SPMV_DEVICES)" if e2e_devices > 1 else ""),
                "devices": e2e_devices,
                "kernel_ms_per_step": e2e_kernel_ms,
                "kernel_ms_note": "CUDA events around the product inside the call (separate leg with the library's "
                                  "event timing on): includes its wait for the first chunk of x and the y stores "
                                  "over PCIe",
                "x_upload": (f"overlapped with the product: {e2e_overlapped} of {Ke} calls (first chunk by a "
                             "PCIe-reading copy kernel, the rest by the copy engine behind per-chunk flags)")
                if e2e_overlapped else "copy kernel before the product",
                "pageable_ms_per_step": e2e_pageable_sec * 1e3,
                "pageable_value": B / e2e_pageable_sec / 1e9,
                "pageable_pin_host_3_ms_per_step": e2e_autopin_sec * 1e3,
                "python_ctypes_ms_per_step": e2e_python_sec * 1e3},
        "clocks": clocks,
    }

    # ---- the callers of the ABI on this matrix ------------------------------
    if hm.kind == "npb" and not args.no_npb and name != "D":
        stage("NPB CG through the ABI")
        res = npb.run_cg(npb_m, addr)
        line["npb_cg"] = {"class": name, "mops": res["mops"], "time_s": res["t_bench"],
                          "zeta": res["zeta"], "verified": res["verified"],
                          "spmv_calls": res["spmv_calls"], "vectors": "pageable host (as cg.f COMMON)"}
    if hm.kind == "npb" and not args.no_npb:
        stage("NPB CG, device-resident")
        cls_ = hm.npb_class
        dres = rm.npb_cg_device(cls_.nonzer, cls_.niter, cls_.shift, use_graph=True)
        line["npb_cg_device_resident"] = {
            "class": name, "mops": dres["mops"], "time_s": dres["seconds"], "zeta": dres["zeta"],
            "verified": bool(abs(dres["zeta"] - cls_.zeta_verify) / cls_.zeta_verify <= 1e-10),
            "spmv_launches": dres["spmv_launches"], "vector_launches": dres["vector_launches"],
            "note": "vectors resident in HBM, CUDA graph per conj_grad (include/b200_cg.h)"}
    if hm.kind == "crsmat" and not args.no_npb:
        r = callers.bicg(hm.a, hm.rowstr, hm.colidx, addr, maxit=20, rtol=0.0)
        line["sparsebench_bicg"] = {"iterations": abs(int(r["its"])), "matprod_calls": int(r["matprod_calls"]),
                                    "ms_per_matprod": 1e3 * r["t_matprod"] / max(r["matprod_calls"], 1),
                                    "t_iter_s": r["t_iter"], "rnorm0": r["rnorm0"], "rnorm": r["rnorm"],
                                    "note": "callers/sparsebench (iter.f:18-104), pageable host vectors"}
    if hm.kind == "powerlaw" and not args.no_npb:
        lens = np.diff(hm.rowstr)
        _, err, sec = callers.pagerank(hm.a, hm.rowstr, hm.colidx, hm.x0, addr, iters=20)
        line["pagerank"] = {"iterations": 20, "ms_per_iteration": 1e3 * sec / 20, "last_delta": err,
                            "row_len_max": int(lens.max()), "row_len_mean": float(lens.mean()),
                            "row_len_cv": float(lens.std() / lens.mean()),
                            "note": "callers/pagerank (main.cpp:125-149), pageable host vectors"}

    # ---- parity + CPU baseline beside it -------------------------------------
    if not args.no_cpu:
        stage("CPU baseline and parity")
        x_host = xs[0].cpu().numpy()
        B_cpu = B
        est = B_cpu / 7.0e9
        nsamp = max(2, min(args.cpu_steps, int(15.0 / est) or 1))
        dt, kind, cores, y_cpu = cpu_reference_spmv(hm.a, hm.rowstr, hm.colidx, x_host, nsamp, 1)
        rm.exec(xs[0], y)
        torch.cuda.synchronize()
        y_gpu = y.cpu().numpy()
        exact = bool(np.array_equal(y_gpu, y_cpu))
        nz = y_cpu != 0
        rel = float(np.max(np.abs(y_gpu - y_cpu)[nz] / np.abs(y_cpu)[nz])) if nz.any() else 0.0
        line["cpu_baseline"] = {
            "value": B / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{nsamp} products of the same {hm.label} matrix ({dt * 1e3:.1f} ms each, libspmv/native.c loop)",
            "gpu_result_bit_identical": exact, "gpu_max_rel_diff": rel}
        dt_omp, _, cores_omp, _ = cpu_reference_spmv(hm.a, hm.rowstr, hm.colidx, x_host, max(nsamp // 2, 1), 1, omp=True)
        line["cpu_baseline_omp"] = {
            "value": B / dt_omp / 1e9, "unit": UNIT, "cores": cores_omp, "kind": "port",
            "note": "row-parallel OpenMP loop, stand-in for libspmv/mkl.c (MKL not in image)"}
        # the end-to-end leg's own result against the oracle (its last call)
        oracle = entry.load_oracle()
        y_e2e_ref = oracle.spmv(hm.a, np.ascontiguousarray(hx_np[(Ke - 1) % len(hx_np)]), hm.rowstr, hm.colidx, omp=True)
        nz2 = y_e2e_ref != 0
        line["e2e"]["y_bit_identical"] = bool(np.array_equal(e2e_y, y_e2e_ref))
        line["e2e"]["y_max_rel_diff"] = (float(np.max(np.abs(e2e_y - y_e2e_ref)[nz2] / np.abs(y_e2e_ref)[nz2]))
                                         if nz2.any() else 0.0)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# B200 arm, N > 1 (one process per GPU under torch.distributed.run)
# --------------------------------------------------------------------------
def run_multi_gpu(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    os.environ["OMP_NUM_THREADS"] = str(max(1, host_cores() // world))   # torchrun pins it to 1
    entry.load_package()
    from lilac_benchmarks_b200 import libspmv, npb, sharded
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 platform has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    host_group = dist.new_group(backend="gloo")       # host-side barriers that keep the GPUs free
    libspmv.lib().b200_spmv_init(local_rank)
    K, W = max(args.steps, 1), max(args.warmup, 3)
    name = args.workload or "D"
    if name not in ("A", "B", "C", "D", "E"):
        raise SystemExit("bench.py: N > 1 runs an NPB class (row-block sharded)")
    label = workload_label(name)
    cls = npb.cg_class(name)
    layout = sharded.ShardLayout.build(cls.na, world)
    lo, hi = layout.local_range(rank)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        dist.barrier(group=host_group)

    # ---- every rank assembles its row block on its own GPU --------------------
    t0 = time.perf_counter()
    dm = npb.NpbDeviceMatrix(name, lo, hi, release_vectors=False)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    rm = dm.resident(kernel=args.kernel)
    t_upload = time.perf_counter() - t0
    n_global, ncols = cls.na, cls.na
    nnz_t = torch.tensor([int(dm.nnz)], dtype=torch.int64, device=dev)
    dist.all_reduce(nnz_t)
    nnz_global = int(nnz_t.item())
    B = algorithmic_bytes(nnz_global, n_global, ncols)

    rng = np.random.default_rng(1234)             # the same global x on every rank
    x_global = rng.random(n_global)
    x_local = torch.from_numpy(x_global[lo:hi]).to(dev)
    sh = sharded.ShardedSpmv(layout, rank, lambda xf, yl: rm.exec(xf, yl), dist=dist, device=dev)
    try:
        psh = sharded.PeerShardedSpmv(libspmv, rm, layout, rank, dist=dist, device=dev)
        ok = 1
    except Exception as exc:                          # no peer access between the GPUs
        print(f"bench.py: peer-memory exchange unavailable ({exc}); using the NCCL allgather", file=sys.stderr)
        psh, ok = None, 0
    okt = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if int(okt.item()) == 0 and psh is not None:
        psh.close()
        psh = None
    # the peer-memory step has three forms (sharded.PeerShardedSpmv); the fastest on these GPUs
    # for this block shape is kept, all three are reported
    peer_forms = psh.calibrate(x_local) if psh is not None else None
    exchange_kind = "nccl" if psh is None else {"fused": "peer-fused", "overlapped": "peer-overlapped",
                                                 "blocking": "peer"}[psh.form]
    stepper = psh or sh

    # ---- parity: one step, every y element of every rank against the OpenMP oracle ------
    y_ok = None
    if not args.no_cpu:
        oracle = entry.load_oracle()
        a_h, rs_h, ci_h = dm.to_host()
        y_ref = oracle.spmv(a_h, x_global, rs_h, ci_h, omp=True)
        y_dev = stepper.step(x_local).cpu().numpy()
        same = 1 if np.array_equal(y_dev, y_ref) else 0
        y2 = sh.step(x_local).cpu().numpy()           # and the NCCL variant
        same = same if np.array_equal(y2, y_ref) else 0
        st = torch.tensor([same], dtype=torch.int32, device=dev)
        dist.all_reduce(st, op=dist.ReduceOp.MIN)
        y_ok = bool(int(st.item()))
        del a_h, rs_h, ci_h, y_ref
    dm.free()

    # same step with the NCCL allgather as the exchange
    ms_n, _ = time_steps(torch, lambda i: sh.step(x_local), K, W, barrier, local_rank)
    tn = torch.tensor([ms_n], dtype=torch.float64, device=dev)
    dist.all_reduce(tn, op=dist.ReduceOp.MAX)
    nccl_ms_per_step = float(tn.item()) / K
    # the rank-local kernel alone (x complete): what the exchange adds on top
    xf = torch.from_numpy(x_global).to(dev)
    yl = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
    ms_k, _ = time_steps(torch, lambda i: rm.exec(xf, yl), K, W, barrier, local_rank)
    tk = torch.tensor([ms_k], dtype=torch.float64, device=dev)
    dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    kernel_only_ms = float(tk.item()) / K
    del xf
    # ---- the headline: the sharded step, device-timed (after the comparison legs) ----------
    ms, clocks = time_steps(torch, lambda i: stepper.step(x_local), K, W, barrier, local_rank)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec_per_step = float(t.item()) / 1e3 / K
    value = B / sec_per_step / 1e9

    # ---- NPB CG, device-resident and row-block sharded ---------------------------
    cg_line = None
    if not args.no_npb:
        def run_cg(kind):
            if kind == "peer":
                drv = sharded.PeerNpbCg(libspmv, rm, layout, rank, cls.shift, dist=dist, device=dev)
            else:
                drv = sharded.ShardedNpbCg(sh, sharded.B200VectorOps(libspmv, dev), cls.shift)
            zh, rh, sec = drv.run(cls.niter, sync=barrier)
            tt = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            count = drv.spmv_count
            if kind == "peer":
                drv.close()
            return zh, float(tt.item()), count

        zeta_nccl, sec_nccl, spmv_count = run_cg("nccl")
        if psh is not None:
            zeta_h, cg_sec, spmv_count = run_cg("peer")
        else:
            zeta_h, cg_sec = zeta_nccl, sec_nccl
        nz1 = cls.nonzer * (cls.nonzer + 1)
        mops = 2.0 * cls.niter * cls.na * (3.0 + nz1 + 25.0 * (5.0 + nz1) + 3.0) / cg_sec / 1e6
        cg_line = {
            "class": name, "mops": mops, "time_s": cg_sec, "zeta": zeta_h[-1],
            "verified": bool(abs(zeta_h[-1] - cls.zeta_verify) / cls.zeta_verify <= 1e-10),
            "spmv_launches_per_rank": spmv_count * rm.launches_per_exec,
            "exchange": "fused into the update / dot kernels over NVLink peer memory "
                        "(include/b200_peer.h), no NCCL call inside conj_grad"
                        if psh is not None else "NCCL allgather + allreduce",
            "nccl_variant": {"time_s": sec_nccl, "zeta": zeta_nccl[-1],
                             "exchange": "allgather of p + 2 one-scalar allreduces per CG iteration"}}

    # ---- rank 0 alone: the same matrix on ONE GPU, and the ABI symbol driving all N ---------
    launches_per_step = rm.launches_per_exec + (0 if (psh is not None and psh.fused) else 1)
    kernel_name = rm.kernel_name
    if psh is not None:
        psh.close()
    rm.release()
    del sh
    torch.cuda.empty_cache()
    host_barrier()
    one_gpu = None
    e2e = None
    if rank == 0 and name != "E":
        os.environ["OMP_NUM_THREADS"] = str(host_cores())
        t0 = time.perf_counter()
        whole = npb.NpbDeviceMatrix(name, 0, cls.na)
        rm1 = whole.resident(kernel=args.kernel)
        torch.cuda.synchronize()
        t_whole = time.perf_counter() - t0
        xs = [torch.from_numpy(rng.random(ncols + 2)).to(dev) for _ in range(4)]
        y1 = torch.zeros(cls.na, dtype=torch.float64, device=dev)
        K1 = max(10, min(K, 100))
        ms1, _ = time_steps(torch, lambda i: rm1.exec(xs[i & 3], y1), K1, 3, torch.cuda.synchronize, local_rank)
        one_gpu = {"ms_per_step": ms1 / K1, "value": B / (ms1 / 1e3 / K1) / 1e9, "unit": UNIT,
                   "kernel": rm1.kernel_name, "measured_in_this_run": True, "steps": K1,
                   "gen_plus_upload_s": round(t_whole, 2)}
        if not args.no_e2e:
            a_h, rs_h, ci_h = whole.to_host()
        rm1.release()
        whole.free()
        del xs, y1
        torch.cuda.empty_cache()
        if not args.no_e2e:
            # ABI mode: spmv_harness_ in THIS process spreads the matrix over all N devices
            os.environ["B200_SPMV_DEVICES"] = ",".join(str(d) for d in range(world))
            hx = [torch.from_numpy(rng.random(ncols + 2)).pin_memory() for _ in range(2)]
            hy = torch.zeros(cls.na, dtype=torch.float64).pin_memory()
            hx_np, hy_np = [v.numpy() for v in hx], hy.numpy()
            addr = libspmv.harness_address()
            t0 = time.perf_counter()
            npb.time_spmv_calls(addr, hy_np, a_h, hx_np, rs_h, ci_h, cls.na, 3)
            t_up = time.perf_counter() - t0
            libspmv.reset_stats()
            Ke = max(10, min(K, 200))
            e2e_sec = npb.time_spmv_calls(addr, hy_np, a_h, hx_np, rs_h, ci_h, cls.na, Ke)
            st = libspmv.stats()
            libspmv.lib().b200_spmv_set_time_kernels(1)
            libspmv.reset_stats()
            npb.time_spmv_calls(addr, hy_np, a_h, hx_np, rs_h, ci_h, cls.na, min(Ke, 50))
            st["kernel_ms"] = libspmv.stats()["kernel_ms"] * Ke / min(Ke, 50)
            libspmv.lib().b200_spmv_set_time_kernels(0)
            px = [np.array(v) for v in hx_np]
            py = np.zeros(cls.na)
            npb.time_spmv_calls(addr, py, a_h, px, rs_h, ci_h, cls.na, 2)
            e2e_pageable_sec = npb.time_spmv_calls(addr, py, a_h, px, rs_h, ci_h, cls.na, max(5, Ke // 4))
            e2e_ok = None
            if not args.no_cpu:
                oracle = entry.load_oracle()
                e2e_ok = bool(np.array_equal(py, oracle.spmv(a_h, px[(max(5, Ke // 4) - 1) % 2], rs_h, ci_h, omp=True)))
            e2e = {"value": B / e2e_sec / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": int(st["h2d_bytes"] // Ke), "d2h_bytes_per_step": int(st["d2h_bytes"] // Ke),
                   "ms_per_step": e2e_sec * 1e3, "steps": Ke,
                   "api": f"spmv_harness_ in one process driving {libspmv.devices_in_use()} GPUs "
                          "(ABI mode, B200_SPMV_DEVICES), C caller loop, pinned caller vectors",
                   "kernel_ms_per_step": st["kernel_ms"] / Ke,
                   "pageable_ms_per_step": e2e_pageable_sec * 1e3,
                   "pageable_value": B / e2e_pageable_sec / 1e9,
                   "y_bit_identical": e2e_ok, "first_calls_incl_upload_s": round(t_up, 2)}
            libspmv.invalidate()
    host_barrier()

    if rank == 0:
        peak, peak_src = measured_peak()
        ach = (B / world) / sec_per_step / 1e9        # per rank: each launch moves B / world
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "gflops": 2.0 * nnz_global / sec_per_step / 1e9,
            "gpu_launches": K * launches_per_step,
            "y_bit_identical": y_ok,
            "config": workload_config(label, n_global, nnz_global, ncols, world),
            "details": {"parallelism": f"{world} equal row blocks, one process per GPU",
                        "untimed_steps_before_timing": max(W, PRECONDITION),
                        "kernel": kernel_name, "exchange": exchange_kind,
                        "exchange_note": {
                            "peer-fused": "one launch per step: the product kernel stores the rank's x slice into every "
                                          "rank's buffer over NVLink peer memory in its prologue, then waits per slice",
                            "peer-overlapped": "b200_peer_post pushes the x slice into every rank's buffer over "
                                               "NVLink peer memory; the product waits per slice in-kernel",
                            "peer": "one exchange kernel over NVLink peer memory, then the product",
                            "nccl": "allgather of x per step (NCCL)"}[exchange_kind],
                        "peer_forms_ms_per_step": peer_forms,
                        "nccl_allgather_variant_ms_per_step": nccl_ms_per_step,
                        "kernel_only_ms_per_step": kernel_only_ms,
                        "same_workload_on_one_gpu": one_gpu,
                        "matrix": "every row block assembled on its GPU (include/b200_npb.h)",
                        "gen_s": round(t_gen, 2), "upload_s": round(t_upload, 3)},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": recorded_traffic(label, kernel_name, world),
                         "traffic_source": "profiles/roofline_traffic.json (ncu --set full capture, one rank's launch)",
                         "peak_source": peak_src,
                         "kernel": f"spmv ({kernel_name}) + exchange, per rank"},
            "e2e": e2e,
            "clocks": clocks,
        }
        if cg_line is not None:
            line["npb_cg_device_resident"] = cg_line
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    # Everything a library prints on stdout (NCCL's version banner, for one) goes to
    # stderr; the one JSON line is written to the real stdout at the end.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None,
                    help="NPB class letter, crsmat170u, pl22 (default: C at N=1, D at N>1)")
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--no-npb", action="store_true", help="skip the caller runs (NPB CG, BiCG, pagerank)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline and the parity check")
    ap.add_argument("--no-e2e", action="store_true", help="N>1: skip the ABI-mode end-to-end leg")
    ap.add_argument("--cpu-steps", type=int, default=120)
    ap.add_argument("--abi-devices", type=int, default=1,
                    help="N=1 run: spread the end-to-end (spmv_harness_) legs over this many GPUs (ABI mode)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, world, rank)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if world == 1:
        run_one_gpu(args)
    else:
        run_multi_gpu(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
