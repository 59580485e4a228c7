#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (BASELINE.json metric).

A step is ONE CSR SpMV y = A x over the named NPB CG matrix.

  N = 1   workload = NPB3.3.1 CG class C (BASELINE config 2): na=150000,
          nnz=36 121 058, the exact matrix of cg.f's makea.  436 MB of
          algorithmic traffic per step, larger than the 126 MB L2, so no L2
          flush is needed between steps.
  N > 1   workload = NPB CG class D (config 5), equal row blocks, one per
          rank; every step re-assembles x with an allgather (NCCL over
          NVLink) and then runs the rank-local kernel.  Strong scaling.

`value`  device-timed (CUDA events on the launching stream), operands resident
         in HBM, whole-job algorithmic GB/s (12 nnz + 4 (n+1) + 8 ncols + 8 n
         bytes per product).
`e2e`    the same metric through the drop-in C-ABI symbol `spmv_harness_`
         with HOST vectors: x host->device and y device->host inside the timed
         region, wall clock.
--impl reference  times the reference's own CPU implementation of the path
         (oracle/_ref/native.so = libspmv/native.c built from the reference
         tree; else the oracle port) on the host cores, same metric/config.

Only the cpu_baseline / --impl reference legs touch oracle/.
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
    """DRAM bytes per launch from the committed ncu --set full capture, if any."""
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

    def _sample(self):
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
                self._sample()
            except Exception:
                pass
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------
# CPU legs (the only users of oracle/)
# --------------------------------------------------------------------------
def cpu_reference_spmv(matrix, x, steps, warmup, omp=False):
    """Time the reference's CPU path on (matrix, x).  Returns (seconds per
    product, kind, cores, last y)."""
    oracle = entry.load_oracle()
    use_ref = oracle.ref_available() and not omp
    kind = "reference" if use_ref else "port"
    y = None
    for _ in range(warmup):
        y = oracle.spmv(matrix.a, x, matrix.rowstr, matrix.colidx, omp=omp, use_ref=use_ref)
    t0 = time.perf_counter()
    for _ in range(steps):
        y = oracle.spmv(matrix.a, x, matrix.rowstr, matrix.colidx, omp=omp, use_ref=use_ref)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, kind, (host_cores() if omp else 1), y


def run_reference_arm(args, world, rank):
    """--impl reference: the reference's own CPU implementation, rank 0 only."""
    if rank != 0:
        return
    entry.load_package()
    from lilac_benchmarks_b200 import npb
    if args.gpus == 1:
        workload = args.workload or "C"
        m = npb.NpbMatrix(workload)
        sample = f"whole NPB class {workload} matrix, one product per step"
        label = f"npb-cg-class-{workload}"
    else:
        workload = args.workload or "D"
        cls = npb.cg_class(workload)
        hi = max(cls.na // 16, 1)
        m = npb.NpbMatrix(workload, 0, hi)
        sample = (f"row block [0,{hi}) of NPB class {workload} (1/16 of the rows, "
                  f"{m.nnz} nnz), one product of the block per step")
        label = f"npb-cg-class-{workload}-rowblock-sharded"
    ncols = int(m.colidx.max())
    x = np.random.default_rng(1234).random(ncols + 2)
    steps, warmup = max(args.steps, 1), max(args.warmup, 1)
    # bound the run: native class C is ~70 ms per product on one core
    steps = min(steps, 200)
    dt, kind, cores, _ = cpu_reference_spmv(m, x, steps, min(warmup, 5))
    dt_omp, _, cores_omp, _ = cpu_reference_spmv(m, x, min(steps, 50), 2, omp=True)
    B = algorithmic_bytes(m.nnz, m.n, ncols)
    val = B / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(warmup, 5), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gflops": 2.0 * m.nnz / dt / 1e9,
        "config": {"workload": label, "rows": m.n, "nnz": int(m.nnz),
                   "implementation": "libspmv/native.c (sequential by construction)"
                   if kind == "reference" else "oracle port of libspmv/native-impl.c"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "cpu_baseline_omp": {"value": B / dt_omp / 1e9, "unit": UNIT, "cores": cores_omp, "kind": "port",
                             "note": "row-parallel OpenMP loop, stand-in for libspmv/mkl.c (MKL not in image)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------
def run_b200(args, world, rank, local_rank):
    import torch
    if world > 1:
        # torchrun pins OMP_NUM_THREADS=1; the matrix generator (callers lib, loaded
        # below) may use this rank's share of the host cores
        os.environ["OMP_NUM_THREADS"] = str(max(1, host_cores() // world))
    entry.load_package()
    from lilac_benchmarks_b200 import libspmv, npb, sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 platform has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    libspmv.lib().b200_spmv_init(local_rank)

    K, W = max(args.steps, 1), max(args.warmup, 3)
    rng = np.random.default_rng(1234 + rank)

    if world == 1:
        workload = args.workload or "C"
        label = f"npb-cg-class-{workload}"
        t0 = time.perf_counter()
        m = npb.NpbMatrix(workload)
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        rm = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx, kernel=args.kernel)
        t_upload = time.perf_counter() - t0
        n_global, nnz_global, ncols = m.n, int(m.nnz), rm.ncols
        xs = [torch.from_numpy(rng.random(ncols + 2)).to(dev) for _ in range(4)]
        y = torch.zeros(m.n, dtype=torch.float64, device=dev)

        def step(i):
            rm.exec(xs[i & 3], y)
        launches_per_step = rm.launches_per_exec
        layout = None
    else:
        workload = args.workload or "D"
        label = f"npb-cg-class-{workload}-rowblock-sharded"
        cls = npb.cg_class(workload)
        layout = sharded.ShardLayout.build(cls.na, world)
        lo, hi = layout.local_range(rank)
        t0 = time.perf_counter()
        m = npb.NpbMatrix(workload, lo, hi, pieces=args.pieces)
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        rm = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx, kernel=args.kernel)
        t_upload = time.perf_counter() - t0
        n_global, ncols = cls.na, cls.na
        nnz_t = torch.tensor([int(m.nnz)], dtype=torch.int64, device=dev)
        dist.all_reduce(nnz_t)
        nnz_global = int(nnz_t.item())
        sh = sharded.ShardedSpmv(layout, rank, lambda xf, yl: rm.exec(xf, yl), dist=dist, device=dev)
        x_local = torch.from_numpy(rng.random(hi - lo)).to(dev)

        # headline exchange: this library's own push kernel over NVLink peer memory;
        # the NCCL allgather variant is timed beside it
        try:
            psh = sharded.PeerShardedSpmv(libspmv, rm, layout, rank, dist=dist, device=dev)
            ok = 1
        except Exception as exc:                          # no peer access between the GPUs
            print(f"bench.py: peer-memory exchange unavailable ({exc}); using the NCCL allgather",
                  file=sys.stderr)
            psh, ok = None, 0
        okt = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 0 and psh is not None:
            psh.close()
            psh = None
        exchange_kind = "peer" if psh is not None else "nccl"

        def step(i):
            (psh or sh).step(x_local)
        # peer: exchange kernel + product; nccl: slot copy + product (+ the NCCL kernel)
        launches_per_step = rm.launches_per_exec + 1
        y = (psh or sh).y_local

    B = algorithmic_bytes(nnz_global, n_global, ncols)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sec_per_step = ms / 1e3 / K
    value = B / sec_per_step / 1e9
    nccl_ms_per_step = None
    if world > 1:
        # same step with the NCCL allgather as the exchange
        for i in range(W):
            sh.step(x_local)
        barrier()
        e0.record()
        for i in range(K):
            sh.step(x_local)
        e1.record()
        barrier()
        tn = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        nccl_ms_per_step = float(tn.item()) / K

    # ---- end to end through the ABI with host vectors --------------------
    Ke = min(K, 2000)
    if world == 1:
        hx = [torch.from_numpy(rng.random(ncols + 2)).pin_memory() for _ in range(4)]
        hy = torch.zeros(m.n, dtype=torch.float64).pin_memory()
        hx_np, hy_np = [t.numpy() for t in hx], hy.numpy()
        # the calls are issued by the C caller loop of callers/npb (what a compiled caller of
        # the ABI such as cg.f pays per call; a ctypes call from Python adds ~15-30 us of
        # argument marshalling that is not the library's)
        addr = libspmv.harness_address()
        npb.time_spmv_calls(addr, hy_np, m.a, hx_np, m.rowstr, m.colidx, m.n, 3)
        libspmv.reset_stats()
        e2e_sec = npb.time_spmv_calls(addr, hy_np, m.a, hx_np, m.rowstr, m.colidx, m.n, Ke)
        st = libspmv.stats()
        h2d, d2h = st["h2d_bytes"] // Ke, st["d2h_bytes"] // Ke
        e2e_kernel_ms = st["kernel_ms"] / Ke
        # pageable caller vectors (what NPB's COMMON arrays are): pinned bounce inside the library
        px = [np.array(v) for v in hx_np]
        py = np.zeros(m.n)
        npb.time_spmv_calls(addr, py, m.a, px, m.rowstr, m.colidx, m.n, 3)
        e2e_pageable_sec = npb.time_spmv_calls(addr, py, m.a, px, m.rowstr, m.colidx, m.n, min(Ke, 500))
        # the same pinned-vector call issued from Python through ctypes
        t0 = time.perf_counter()
        for i in range(min(Ke, 500)):
            libspmv.spmv_harness(hy_np, m.a, hx_np[i & 3], m.rowstr, m.colidx, m.n)
        e2e_python_sec = (time.perf_counter() - t0) / min(Ke, 500)
    else:
        lo, hi = layout.local_range(rank)
        hx = torch.from_numpy(rng.random(hi - lo)).pin_memory()
        hy = torch.zeros(hi - lo, dtype=torch.float64).pin_memory()
        dx = torch.zeros(hi - lo, dtype=torch.float64, device=dev)

        def e2e_step():
            dx.copy_(hx, non_blocking=True)
            sh.step(dx)
            hy.copy_(sh.y_local, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            e2e_step()
        barrier()
        e2e_sec = (time.perf_counter() - t0) / Ke
        t = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
        h2d = d2h = n_global * 8          # summed over ranks
        e2e_kernel_ms = None
        e2e_pageable_sec = None
        e2e_python_sec = None

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel: the SpMV kernel; its average launch duration over the
        # timed region (per rank each launch moves B/world algorithmic bytes)
        ach = (B / world) / sec_per_step / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "gflops": 2.0 * nnz_global / sec_per_step / 1e9,
            "config": {"workload": label, "rows": n_global, "nnz": nnz_global, "ncols": ncols,
                       "kernel": rm.kernel_name, "algorithmic_bytes_per_step": B,
                       "l2_policy": "inputs larger than L2 (no flush)" if B / world > 126e6 * 1.5
                       else "matrix block comparable to L2: HBM fraction may read > 1",
                       "x_vectors_rotated": 4 if world == 1 else 1,
                       "exchange": None if world == 1 else
                       ("x slices pushed into every rank's buffer by this library's kernel over NVLink "
                        "peer memory (include/b200_peer.h)" if exchange_kind == "peer"
                        else "allgather of x per step (NCCL)"),
                       "nccl_allgather_variant_ms_per_step": nccl_ms_per_step,
                       "same_workload_on_one_gpu": None if (world == 1 or workload != "D") else {
                           "ms_per_step": 1.5457, "value": 5410.5, "unit": UNIT,
                           "source": "profiles/r01_run58_sweep_register_staged_x.txt (class D, ring PANEL "
                                     "kernel, 1xB200; SELL kernel: 2.738 ms)"},
                       "gen_s": round(t_gen, 2), "upload_s": round(t_upload, 3)},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": recorded_traffic(label, rm.kernel_name, world),
                         "peak_source": peak_src,
                         "kernel": f"spmv ({rm.kernel_name})" + ("" if world == 1 else " + exchange, per rank")},
            "e2e": {"value": B / e2e_sec / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_sec * 1e3, "steps": Ke,
                    "api": "spmv_harness_ called from the C caller loop (callers/npb), pinned caller vectors"
                    if world == 1
                    else "ShardedSpmv.step with pinned host slices",
                    "kernel_ms_per_step": e2e_kernel_ms,
                    "pageable_ms_per_step": None if e2e_pageable_sec is None else e2e_pageable_sec * 1e3,
                    "python_ctypes_ms_per_step": None if e2e_python_sec is None else e2e_python_sec * 1e3},
            "gpu_launches": K * launches_per_step,
            "clocks": clocks,
        }

    # ---- NPB CG, device-resident and row-block sharded (N > 1) --------------
    if world > 1 and not args.no_npb:
        cls_ = npb.cg_class(workload)
        def run_cg(kind):
            if kind == "peer":
                drv = sharded.PeerNpbCg(libspmv, rm, layout, rank, cls_.shift, dist=dist, device=dev)
            else:
                drv = sharded.ShardedNpbCg(sh, sharded.B200VectorOps(libspmv, dev), cls_.shift)
            zh, rh, sec = drv.run(cls_.niter, sync=barrier)
            tt = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            count = drv.spmv_count
            if kind == "peer":
                drv.close()
            return zh, float(tt.item()), count

        zeta_nccl, sec_nccl, spmv_count = run_cg("nccl")
        if exchange_kind == "peer":
            zeta_h, cg_sec, spmv_count = run_cg("peer")
        else:
            zeta_h, cg_sec = zeta_nccl, sec_nccl

        class _Cnt:
            pass
        cg = _Cnt()
        cg.spmv_count, cg.collectives = spmv_count, 0
        if rank == 0:
            nz1 = cls_.nonzer * (cls_.nonzer + 1)
            mops = 2.0 * cls_.niter * cls_.na * (3.0 + nz1 + 25.0 * (5.0 + nz1) + 3.0) / cg_sec / 1e6
            line["npb_cg_device_resident"] = {
                "class": workload, "mops": mops, "time_s": cg_sec, "zeta": zeta_h[-1],
                "verified": bool(abs(zeta_h[-1] - cls_.zeta_verify) / cls_.zeta_verify <= 1e-10),
                "spmv_launches_per_rank": cg.spmv_count * rm.launches_per_exec,
                "collectives": cg.collectives,
                "exchange": "fused into the update / dot kernels over NVLink peer memory "
                            "(include/b200_peer.h), no NCCL call inside conj_grad"
                            if exchange_kind == "peer" else "NCCL allgather + allreduce",
                "nccl_variant": {"time_s": sec_nccl, "zeta": zeta_nccl[-1],
                                 "exchange": "allgather of p + 2 one-scalar allreduces per CG iteration"}}

    # ---- NPB CG whole benchmark through the ABI + CPU baseline (N = 1) ----
    if world == 1 and rank == 0:
        if not args.no_npb:
            res = npb.run_cg(m, libspmv.harness_address())
            line["npb_cg"] = {"class": workload, "mops": res["mops"], "time_s": res["t_bench"],
                              "zeta": res["zeta"], "verified": res["verified"],
                              "spmv_calls": res["spmv_calls"], "vectors": "pageable host (as cg.f COMMON)"}
        if not args.no_npb:
            cls_ = npb.cg_class(workload)
            dev = rm.npb_cg_device(cls_.nonzer, cls_.niter, cls_.shift, use_graph=True)
            line["npb_cg_device_resident"] = {
                "class": workload, "mops": dev["mops"], "time_s": dev["seconds"], "zeta": dev["zeta"],
                "verified": bool(abs(dev["zeta"] - cls_.zeta_verify) / cls_.zeta_verify <= 1e-10),
                "spmv_launches": dev["spmv_launches"], "vector_launches": dev["vector_launches"],
                "note": "vectors resident in HBM, CUDA graph per conj_grad (include/b200_cg.h)"}
        if not args.no_cpu:
            x_host = xs[0].cpu().numpy()
            nsamp = args.cpu_steps
            dt, kind, cores, y_cpu = cpu_reference_spmv(m, x_host, nsamp, 1)
            rm.exec(xs[0], y)
            torch.cuda.synchronize()
            line["cpu_baseline"] = {
                "value": B / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                "sample": f"{nsamp} products of the same class {workload} matrix "
                          f"({dt * 1e3:.1f} ms each, libspmv/native.c loop)",
                "gpu_result_bit_identical": bool(np.array_equal(y.cpu().numpy(), y_cpu))}
            dt_omp, _, cores_omp, _ = cpu_reference_spmv(m, x_host, max(nsamp // 2, 1), 1, omp=True)
            line["cpu_baseline_omp"] = {
                "value": B / dt_omp / 1e9, "unit": UNIT, "cores": cores_omp, "kind": "port",
                "note": "row-parallel OpenMP loop, stand-in for libspmv/mkl.c (MKL not in image)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        if psh is not None:
            psh.close()
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
    ap.add_argument("--workload", default=None, help="NPB class letter (default C at N=1, D at N>1)")
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--no-npb", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=120)
    ap.add_argument("--pieces", type=int, default=8,
                    help="N>1: build each rank's row block in this many pieces (bounds host memory)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, world, rank)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run "
                             "(one rank per GPU)")
    run_b200(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
