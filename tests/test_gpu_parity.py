"""Parity tests proper: the CUDA path, called through the C ABI, against the
oracle on the same seeded inputs.  Bit-exact for the order-preserving kernel
(integer/byte-style bar, because it performs the reference's operations in the
reference's order); the re-ordering kernels are held to the north star's
1e-12 relative tolerance on non-cancelling inputs and to 1e-13 of sum|terms|
on any input."""
import subprocess

import numpy as np
import pytest

from conftest import make_csr

pytestmark = pytest.mark.gpu

@pytest.fixture(autouse=True)
def fresh_cache(libspmv):
    """numpy may hand a freed matrix's address to the next test's arrays; the
    resident cache is keyed by host pointers, so start every test clean."""
    libspmv.invalidate()
    yield


REL_TOL_F64 = 1e-12      # north star: per-element relative, fp64
REL_TOL_F32 = 2e-3       # parboil tools/compare-output:18-25 (0.2 % relative)


def _harness(libspmv, a, x, rowstr, colidx, rows=None):
    rows = len(rowstr) - 1 if rows is None else rows
    y = np.full(rows, np.nan, dtype=a.dtype)
    fn = libspmv.f_spmv_harness if a.dtype == np.float32 else libspmv.spmv_harness
    fn(y, a, x, rowstr, colidx, rows)
    return y


def test_reference_kat_through_the_abi(libspmv, kat):
    """libspmv/test.cpp:44-54, both precisions, exact equality."""
    for dt in (np.float64, np.float32):
        a = np.array(kat["a"], dtype=dt)
        x = np.array(kat["x"], dtype=dt)
        rowstr = np.array(kat["rowstr"], dtype=np.int32)
        colidx = np.array(kat["colidx"], dtype=np.int32)
        y = _harness(libspmv, a, x, rowstr, colidx, kat["rows"])
        assert np.array_equal(y, np.array(kat["y"], dtype=dt))


def test_reference_test_binary_passes_on_b200_so(libspmv, oracle):
    """The reference's own unit test (libspmv/test.cpp) dlopens ./b200.so."""
    if not oracle.REF_TEST_BIN.exists():
        pytest.skip("oracle/_ref/test not built")
    proc = subprocess.run([str(oracle.REF_TEST_BIN), "b200"], cwd=str(libspmv.B200_SO.parent),
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert "success!" in proc.stderr, proc.stderr


def test_golden_reference_outputs_bit_exact(libspmv, native_vectors):
    for name, v in native_vectors.items():
        if "long_row" in name:
            continue                      # rows longer than a tile are tree-reduced
        y = _harness(libspmv, v["a"], v["x"], v["rowstr"], v["colidx"])
        assert np.array_equal(y, v["y"]), name


def test_golden_long_rows_within_tolerance(libspmv, oracle, native_vectors):
    for name, v in native_vectors.items():
        if "long_row" not in name:
            continue
        y = _harness(libspmv, v["a"], v["x"], v["rowstr"], v["colidx"])
        f32 = v["a"].dtype == np.float32
        a64, x64 = v["a"].astype(np.float64), v["x"].astype(np.float64)
        _, mag = oracle.spmv_extended(a64, x64, v["rowstr"], v["colidx"])
        eps = 6e-8 if f32 else 1.2e-16
        assert np.all(np.abs(y.astype(np.float64) - v["y"]) <= 64 * eps * mag + 1e-300), name


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [
    dict(n=1, ncols=1, mean=1), dict(n=7, ncols=3, mean=2), dict(n=1000, ncols=1000, mean=5),
    dict(n=3000, ncols=500, mean=130), dict(n=64, ncols=20000, mean=1500),
    dict(n=20000, ncols=20000, mean=1),
])
def test_random_matrices_bit_exact(libspmv, oracle, dtype, shape):
    rng = np.random.default_rng(shape["n"] * 31 + shape["mean"])
    lens = rng.poisson(shape["mean"], shape["n"])
    lens[rng.random(shape["n"]) < 0.15] = 0
    a, c, rowstr, x = make_csr(rng, shape["n"], shape["ncols"], lens, dtype=dtype, sort=False)
    y = _harness(libspmv, a, x, rowstr, c)
    assert np.array_equal(y, oracle.spmv(a, x, rowstr, c))


def _exec_resident(libspmv, a, x, rowstr, c, kernel, env=None):
    import os
    import torch
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update({k: str(v) for k, v in (env or {}).items()})
    try:
        m = libspmv.ResidentMatrix(a, rowstr, c, kernel=kernel)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    tdt = torch.float32 if a.dtype == np.float32 else torch.float64
    dx = torch.from_numpy(np.ascontiguousarray(x[:max(m.ncols, 1)])).cuda()
    dy = torch.full((m.rows,), float("nan"), dtype=tdt, device="cuda")
    m.exec(dx, dy)
    return m, dy.cpu().numpy()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [
    dict(n=1, ncols=1, mean=1), dict(n=33, ncols=70, mean=9), dict(n=1000, ncols=1000, mean=25),
    dict(n=5000, ncols=40000, mean=130), dict(n=300, ncols=70000, mean=900),
    dict(n=2500, ncols=131072, mean=64),
])
@pytest.mark.parametrize("panel_env", [
    {}, {"B200_SPMV_PANEL_COLS": 256, "B200_SPMV_PANEL_ROWS": 64},
    {"B200_SPMV_PANEL_COLS": 4096, "B200_SPMV_PANEL_ROWS": 1024},
])
def test_panel_kernel_bit_exact_on_sorted_rows(libspmv, oracle, dtype, shape, panel_env):
    """The column-panel layout keeps every row's left-to-right order when the
    columns are sorted: bit-identical to the reference loop for any panel
    width / row-block height, including duplicate columns and empty rows."""
    rng = np.random.default_rng(shape["n"] * 7 + shape["mean"])
    lens = rng.poisson(shape["mean"], shape["n"])
    lens[rng.random(shape["n"]) < 0.1] = 0
    a, c, rowstr, x = make_csr(rng, shape["n"], shape["ncols"], lens, dtype=dtype, sort=True)
    y0 = oracle.spmv(a, x, rowstr, c)
    # SMALL family off: this test is about the panel layout and its place in the automatic choice
    m, y = _exec_resident(libspmv, a, x, rowstr, c, "auto", dict(panel_env, B200_SPMV_SMALL=0))
    if not panel_env and len(c) and shape["ncols"] <= 20000 and shape["mean"] >= 9:
        assert m.kernel_name == "panel", (m.kernel_name, m.ncols, m.nnz)
    assert np.array_equal(y, y0)
    # the ordered kernel on the same input agrees too
    m2, y2 = _exec_resident(libspmv, a, x, rowstr, c, "ordered")
    assert m2.kernel_name == "ordered" and np.array_equal(y2, y0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [
    dict(n=1, ncols=1, mean=1), dict(n=33, ncols=70, mean=9), dict(n=1000, ncols=1000, mean=25),
    dict(n=5000, ncols=40000, mean=130), dict(n=300, ncols=70000, mean=900),
    dict(n=9000, ncols=131072, mean=64), dict(n=4500, ncols=1500000, mean=450),
])
@pytest.mark.parametrize("panel_env", [
    {"B200_SPMV_PANEL_G": 2}, {"B200_SPMV_PANEL_G": 4}, {"B200_SPMV_PANEL_G": 8},
    {"B200_SPMV_PANEL_G": 4, "B200_SPMV_PANEL_COLS": 256, "B200_SPMV_PANEL_ROWS": 256},
    {"B200_SPMV_PANEL_G": 8, "B200_SPMV_PANEL_COLS": 4096, "B200_SPMV_PANEL_ROWS": 4096},
    {"B200_SPMV_PANEL_G": 4, "B200_SPMV_PANEL_ROWS": 1280, "B200_SPMV_PANEL_NBUF": 1},
    {"B200_SPMV_PANEL_G": 2, "B200_SPMV_PANEL_ROWS": 700, "B200_SPMV_PANEL_TMA": 0},
    {"B200_SPMV_PANEL_G": 4, "B200_SPMV_PANEL_RING_K": 2, "B200_SPMV_PANEL_RING_S": 3},
    {"B200_SPMV_PANEL_G": 2, "B200_SPMV_PANEL_ROWS": 1280, "B200_SPMV_PANEL_TMAX": 640, "B200_SPMV_PANEL_RING_K": 2, "B200_SPMV_PANEL_RING_S": 3},
])
def test_ring_panel_kernel_bit_exact_on_sorted_rows(libspmv, oracle, dtype, shape, panel_env):
    """Tall row blocks with G rows per lane stream and the matrix stream
    through per-warp shared-memory rings (spmv_panelg.cu + spmv_panelr.cu, the
    layout for wide matrices such as NPB class D row blocks): bit-identical to
    the reference loop for every G, row-block height, panel width and ring
    geometry, including duplicate columns, empty rows and ragged last
    blocks."""
    rng = np.random.default_rng(shape["n"] * 11 + shape["mean"])
    lens = rng.poisson(shape["mean"], shape["n"])
    lens[rng.random(shape["n"]) < 0.1] = 0
    a, c, rowstr, x = make_csr(rng, shape["n"], shape["ncols"], lens, dtype=dtype, sort=True)
    y0 = oracle.spmv(a, x, rowstr, c)
    env = dict(panel_env, B200_SPMV_PANEL_FMT=2)
    if shape["ncols"] > 250 * env.get("B200_SPMV_PANEL_COLS", shape["ncols"]):
        del env["B200_SPMV_PANEL_COLS"]          # more than 256 panels: the plan is refused
    m, y = _exec_resident(libspmv, a, x, rowstr, c, "panel", env)
    if len(c):
        assert m.kernel_name == "panel", (m.kernel_name, m.ncols, m.nnz)
    assert np.array_equal(y, y0)


def test_panel_falls_back_when_rows_are_unsorted(libspmv, oracle):
    rng = np.random.default_rng(77)
    a, c, rowstr, x = make_csr(rng, 2000, 3000, rng.poisson(40, 2000), sort=False)
    m, y = _exec_resident(libspmv, a, x, rowstr, c, "panel")
    assert m.kernel_name == "sell"
    assert np.array_equal(y, oracle.spmv(a, x, rowstr, c))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("sort", [True, False])
@pytest.mark.parametrize("shape", [
    dict(n=1, ncols=1, mean=1), dict(n=65, ncols=9, mean=3), dict(n=5000, ncols=300000, mean=5),
    dict(n=3000, ncols=1500000, mean=120), dict(n=777, ncols=50000, mean=400),
])
@pytest.mark.parametrize("env", [{}, {"B200_SPMV_SELL_FMT": 1}, {"B200_SPMV_SELL_FMT": 1, "B200_SPMV_SELL_G": 2},
                                 {"B200_SPMV_SELL_FMT": 1, "B200_SPMV_SELL_G": 16, "B200_SPMV_SELL_U": 4},
                                 {"B200_SPMV_SELL_FMT": 1, "B200_SPMV_SELL_U": 6},
                                 {"B200_SPMV_SELL_FMT": 1, "B200_SPMV_SELL_G": 8, "B200_SPMV_SELL_U": 8},
                                 {"B200_SPMV_SELL_FMT": 0}, {"B200_SPMV_SELL_FMT": 0, "B200_SPMV_SELL_U": 6},
                                 {"B200_SPMV_SELL_FMT": 0, "B200_SPMV_SELL_G": 1, "B200_SPMV_SELL_ROWS": 64}])
def test_sell_kernel_bit_exact_any_column_order(libspmv, oracle, dtype, sort, shape, env):
    """Lane streams over L2 gathers: left-to-right rows whatever the column
    order, for every row up to the cap."""
    rng = np.random.default_rng(shape["n"] + shape["mean"] + int(sort))
    lens = rng.poisson(shape["mean"], shape["n"])
    lens[rng.random(shape["n"]) < 0.1] = 0
    a, c, rowstr, x = make_csr(rng, shape["n"], shape["ncols"], lens, dtype=dtype, sort=sort)
    m, y = _exec_resident(libspmv, a, x, rowstr, c, "sell", env)
    assert m.kernel_name == "sell" or m.nnz == 0
    assert np.array_equal(y, oracle.spmv(a, x, rowstr, c))


def test_sell_long_rows_go_to_the_cta_reduction(libspmv, oracle):
    """Rows above the cap are tree-reduced (re-ordered): exact on the short
    rows, within 1e-12 relative on the long ones for non-cancelling input."""
    rng = np.random.default_rng(123)
    n = 4000
    lens = rng.poisson(6, n)
    long_ids = rng.choice(n, 25, replace=False)
    lens[long_ids] = rng.integers(300, 70000, 25)
    a, c, rowstr, x = make_csr(rng, n, 200000, lens, positive=True, sort=False)
    m, y = _exec_resident(libspmv, a, x, rowstr, c, "auto")
    assert m.kernel_name == "sell" and m.launches_per_exec == 3     # tiles + chunks + carry fix-up
    y0 = oracle.spmv(a, x, rowstr, c)
    is_long = np.zeros(n, dtype=bool)
    is_long[long_ids] = True
    assert np.array_equal(y[~is_long], y0[~is_long])
    assert np.all(np.abs(y - y0)[is_long] <= REL_TOL_F64 * np.abs(y0)[is_long])


def test_ragged_edges(libspmv, oracle):
    rng = np.random.default_rng(11)
    # all rows empty, rowstr base offset, last column == ncols, x longer than ncols
    a = np.zeros(0)
    c = np.zeros(0, dtype=np.int32)
    rowstr = np.ones(6, dtype=np.int32)
    assert np.array_equal(_harness(libspmv, a, np.ones(4), rowstr, c), np.zeros(5))
    lens = rng.integers(0, 9, 200)
    a, c, rowstr, x = make_csr(rng, 200, 77, lens, base=1234)
    x_long = np.concatenate([x, rng.standard_normal(50)])
    assert np.array_equal(_harness(libspmv, a, x_long, rowstr, c), oracle.spmv(a, x, rowstr, c))


def test_rows_just_around_the_tile(libspmv, oracle):
    """Rows of length tile-3 .. tile+1 exercise the block packing limits."""
    rng = np.random.default_rng(5)
    for dtype, tile in ((np.float64, 4096), (np.float32, 8192)):
        lens = [tile - 3, 1, tile - 2, 0, tile - 1, 3, tile, tile + 1, 2]
        a, c, rowstr, x = make_csr(rng, len(lens), 5000, lens, dtype=dtype, positive=True)
        y = _harness(libspmv, a, x, rowstr, c)
        y0 = oracle.spmv(a, x, rowstr, c)
        short = np.array(lens) <= tile - 2
        assert np.array_equal(y[short], y0[short])
        tol = REL_TOL_F64 if dtype == np.float64 else 1e-5
        assert np.all(np.abs(y - y0) <= tol * np.abs(y0))


def test_cache_keys_on_pointers_and_shape(libspmv, oracle):
    rng = np.random.default_rng(2)
    libspmv.invalidate()
    libspmv.reset_stats()
    a, c, rowstr, x = make_csr(rng, 500, 400, rng.poisson(20, 500))
    y0 = oracle.spmv(a, x, rowstr, c)
    for _ in range(5):
        assert np.array_equal(_harness(libspmv, a, x, rowstr, c), y0)
    st = libspmv.stats()
    assert st["uploads"] == 1 and st["calls"] == 5 and st["kernel_launches"] == 5
    assert st["h2d_bytes"] == 5 * 8 * c.max() and st["d2h_bytes"] == 5 * 8 * 500
    # a second matrix lives beside the first one
    a2, c2, rowstr2, x2 = make_csr(rng, 300, 350, rng.poisson(10, 300))
    assert np.array_equal(_harness(libspmv, a2, x2, rowstr2, c2), oracle.spmv(a2, x2, rowstr2, c2))
    assert np.array_equal(_harness(libspmv, a, x, rowstr, c), y0)
    assert libspmv.stats()["uploads"] == 2
    # explicit invalidation after an in-place change of the values
    a *= 2.0
    libspmv.invalidate()
    assert np.array_equal(_harness(libspmv, a, x, rowstr, c), oracle.spmv(a, x, rowstr, c))
    assert libspmv.stats()["uploads"] == 3


def test_vector_kernel_tolerance(libspmv, oracle):
    import torch
    rng = np.random.default_rng(9)
    a, c, rowstr, x = make_csr(rng, 4000, 3000, rng.poisson(70, 4000), positive=True)
    y0 = oracle.spmv(a, x, rowstr, c)
    m = libspmv.ResidentMatrix(a, rowstr, c, kernel="vector")
    assert m.kernel_name == "vector" and m.ncols == c.max()
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty(4000, dtype=torch.float64, device="cuda")
    m.exec(dx, dy)
    y = dy.cpu().numpy()
    nz = y0 != 0
    assert np.all(np.abs(y - y0)[nz] <= REL_TOL_F64 * np.abs(y0)[nz])
    # cancelling input: bound against sum|terms| instead
    a2, c2, rowstr2, x2 = make_csr(rng, 4000, 3000, rng.poisson(70, 4000))
    m2 = libspmv.ResidentMatrix(a2, rowstr2, c2, kernel="vector")
    m2.exec(torch.from_numpy(x2).cuda(), dy)
    _, mag = oracle.spmv_extended(a2, x2, rowstr2, c2)
    assert np.all(np.abs(dy.cpu().numpy() - oracle.spmv(a2, x2, rowstr2, c2)) <= 1e-13 * mag + 1e-300)


def test_resident_matrix_row_block(libspmv, oracle, npb):
    """A row block is (a, colidx, rowstr + lo, hi - lo): the ABI's own addressing."""
    import torch
    m = npb.NpbMatrix("S")
    x = np.random.default_rng(1).random(m.n)
    y0 = oracle.spmv(m.a, x, m.rowstr, m.colidx)
    bounds = libspmv.partition_rows(m.rowstr, 3)
    dx = torch.from_numpy(x).cuda()
    parts = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        blk = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx, rows=int(hi - lo), row_lo=int(lo))
        dy = torch.empty(int(hi - lo), dtype=torch.float64, device="cuda")
        blk.exec(dx, dy)
        parts.append(dy.cpu().numpy())
        hist, mn, mx = blk.row_histogram()
        assert sum(hist) == hi - lo and mn >= 1 and mx <= 127
    assert np.array_equal(np.concatenate(parts), y0)


@pytest.mark.parametrize("cls", ["S", "A"])
def test_npb_cg_history_bit_identical_to_cpu_run(libspmv, oracle, npb, npb_history, cls):
    """Whole NPB CG through libb200-spmv: zeta verifies (cg.f:363-368) and --
    because every product is bit-identical -- the complete ||r|| / zeta
    history equals the CPU run and the reference C twin's printout."""
    m = npb.NpbMatrix(cls)
    gpu = npb.run_cg(m, libspmv.harness_address())
    assert gpu["verified"] and gpu["spmv_calls"] == (m.cls.niter + 1) * 26
    cpu = npb.run_cg(m, oracle.harness_address())
    assert np.array_equal(gpu["zeta_hist"], cpu["zeta_hist"])
    assert np.array_equal(gpu["rnorm_hist"], cpu["rnorm_hist"])
    gold = npb_history["classes"][cls]
    assert [f"{z:.13f}" for z in gpu["zeta_hist"]] == gold["zeta"]


def test_class_c_full_size_properties(libspmv, oracle, npb):
    """BASELINE config 2 at full size: element-wise equality with the (OpenMP)
    oracle on a seeded x and on the all-ones start vector, linearity in x, and
    the nnz self-check of the generator."""
    import torch
    m = npb.NpbMatrix("C")
    assert m.n == 150000 and m.nnz == 36121058
    rm = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx)
    assert rm.ncols == m.n and rm.algorithmic_bytes == 12 * m.nnz + 4 * (m.n + 1) + 16 * m.n
    rng = np.random.default_rng(42)
    dy = torch.empty(m.n, dtype=torch.float64, device="cuda")
    for x in (np.ones(m.n), rng.standard_normal(m.n)):
        rm.exec(torch.from_numpy(x).cuda(), dy)
        assert np.array_equal(dy.cpu().numpy(), oracle.spmv(m.a, x, m.rowstr, m.colidx, omp=True))
    # exact scaling by a power of two commutes with every rounding
    x = rng.standard_normal(m.n)
    rm.exec(torch.from_numpy(x).cuda(), dy)
    y1 = dy.cpu().numpy().copy()
    rm.exec(torch.from_numpy(4.0 * x).cuda(), dy)
    assert np.array_equal(dy.cpu().numpy(), 4.0 * y1)


def test_parboil_fp32_golden(libspmv, oracle, parboil):
    """f_spmv_harness_ on the exact CSR parboil's CPU caller builds
    (parboil/benchmarks/spmv/src/cpu/main.c:80-95): bit-identical to the fp32
    oracle, and inside the reference checker's tolerance of the golden files."""
    from conftest import parboil_compare
    for name in ("small", "medium"):
        d = parboil[name]
        y = _harness(libspmv, d["a"], d["x"], d["rowstr"], d["colidx"])
        assert np.array_equal(y, oracle.spmv(d["a"], d["x"], d["rowstr"], d["colidx"])), name
        assert parboil_compare(d["y_golden"], y), name
    assert np.array_equal(_harness(libspmv, parboil["small"]["a"], parboil["small"]["x"],
                                   parboil["small"]["rowstr"], parboil["small"]["colidx"]),
                          parboil["small"]["y_golden"])


def test_reference_callers_relinked_unchanged(libspmv, oracle, tmp_path):
    """The reference's own parboil spmv CPU program and bfs, compiled from the
    reference sources and linked against b200.so instead of libnative-spmv.so
    (oracle/Makefile `relink`), run on the reference's inputs."""
    import os
    from pathlib import Path
    golden = Path(__file__).resolve().parent / "golden"
    ref_dir = oracle.REF_TEST_BIN.parent
    pb, bfs = ref_dir / "parboil_spmv.b200", ref_dir / "bfs.b200"
    if not pb.exists() or not bfs.exists():
        pytest.skip("relinked reference binaries not built (reference tree absent at build time)")
    out = tmp_path / "y.out"
    proc = subprocess.run([str(pb), "-i", f"{golden / 'parboil_small_1138_bus.mtx'},"
                           f"{golden / 'parboil_small_vector.bin'}", "-o", str(out)],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout
    assert out.read_bytes() == (golden / "parboil_small_1138_bus.mtx.out").read_bytes()
    with open(golden / "bfs_input.mtx") as fin:
        proc = subprocess.run([str(bfs)], stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout
    assert float(proc.stdout.strip().splitlines()[-1]) > 0.0


@pytest.mark.parametrize("cls,use_graph", [("S", False), ("S", True), ("A", True), ("B", True)])
def test_device_resident_npb_cg_verifies(libspmv, npb, npb_history, cls, use_graph):
    """include/b200_cg.h: NPB CG with x, z, p, q, r resident in HBM.  zeta
    verifies against cg.f:122-166 (1e-10); the per-iteration zeta agrees with
    the host-algebra history to the accuracy the re-ordered dot products allow."""
    m = npb.NpbMatrix(cls)
    rm = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx)
    res = rm.npb_cg_device(m.cls.nonzer, m.cls.niter, m.cls.shift, use_graph=use_graph)
    ref = npb_history["zeta_verify"][cls]
    assert abs(res["zeta"] - ref) / ref <= 1e-10
    assert res["spmv_launches"] == 26 * m.cls.niter * rm.launches_per_exec
    if cls in npb_history["classes"]:
        gold = np.array([float(z) for z in npb_history["classes"][cls]["zeta"]])
        assert np.allclose(res["zeta_hist"], gold, rtol=1e-9, atol=0)
    # deterministic run to run (fixed reduction order, no atomics)
    res2 = rm.npb_cg_device(m.cls.nonzer, m.cls.niter, m.cls.shift, use_graph=use_graph)
    assert np.array_equal(res["zeta_hist"], res2["zeta_hist"])


@pytest.mark.parametrize("cls", ["S", "A"])
def test_peer_memory_cg_single_rank_group(libspmv, npb, npb_history, cls):
    """include/b200_peer.h on a group of one rank: the fused update + exchange
    kernels, the epoch flags and the slot reductions, with this GPU as its own
    only peer.  zeta must verify (cg.f:122-166)."""
    from lilac_benchmarks_b200 import sharded
    m = npb.NpbMatrix(cls)
    rm = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx)
    layout = sharded.ShardLayout.build(m.n, 1)
    cg = sharded.PeerNpbCg(libspmv, rm, layout, 0, m.cls.shift)
    try:
        zeta, rnorm, _ = cg.run(m.cls.niter)
    finally:
        cg.close()
    ref = npb_history["zeta_verify"][cls]
    assert abs(zeta[-1] - ref) / ref <= 1e-10
    assert cg.spmv_count == 26 * m.cls.niter
    gold = np.array([float(z) for z in npb_history["classes"][cls]["zeta"]])
    assert np.allclose(zeta, gold, rtol=1e-9, atol=0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_merge_kernel_nnz_split_with_carry_fixup(libspmv, oracle, dtype):
    """B200_KERNEL_MERGE: every row goes through the nnz-split chunks (one warp
    per 512 entries) and multi-chunk rows are finished by the ordered carry
    fix-up.  Re-ordered sum: 1e-12 relative on non-cancelling fp64 input,
    deterministic run to run."""
    import torch
    rng = np.random.default_rng(99)
    n = 3000
    lens = rng.poisson(8, n)
    lens[rng.choice(n, 40, replace=False)] = rng.integers(500, 40000, 40)
    lens[rng.random(n) < 0.05] = 0
    a, c, rowstr, x = make_csr(rng, n, 100000, lens, dtype=dtype, positive=True, sort=False)
    m, y = _exec_resident(libspmv, a, x, rowstr, c, "merge")
    assert m.kernel_name == "merge" and m.launches_per_exec == 4   # tiles, 8-lane chunks, warp chunks, fix-up
    y0 = oracle.spmv(a, x, rowstr, c)
    tol = REL_TOL_F64 if dtype == np.float64 else 2e-5
    nz = y0 != 0
    assert np.all(np.abs(y - y0)[nz] <= tol * np.abs(y0)[nz]) and np.all(y[~nz] == 0)
    _, y2 = _exec_resident(libspmv, a, x, rowstr, c, "merge")
    assert np.array_equal(y, y2)


GUARD_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
entry.load_package()
oracle = entry.load_oracle()
from lilac_benchmarks_b200 import libspmv
rng = np.random.default_rng(4)
n, ncols = 6000, 5000
lens = rng.poisson(40, n)
rowstr = np.empty(n + 1, dtype=np.int32); rowstr[0] = 1; rowstr[1:] = 1 + np.cumsum(lens)
nnz = int(lens.sum())
colidx = rng.integers(1, ncols + 1, nnz).astype(np.int32)
a = rng.standard_normal(nnz)
x = rng.standard_normal(ncols)
y = np.zeros(n)
libspmv.spmv_harness(y, a, x, rowstr, colidx, n)
assert np.array_equal(y, oracle.spmv(a, x, rowstr, colidx))
libspmv.spmv_harness(y, a, x, rowstr, colidx, n)
assert libspmv.stats()["uploads"] == 1
a[nnz // 2] = 123.0            # lands on a protected page: SIGSEGV -> handler -> cache entry stale
colidx[nnz // 3] = 1
libspmv.spmv_harness(y, a, x, rowstr, colidx, n)
assert libspmv.stats()["uploads"] == 2, libspmv.stats()
assert np.array_equal(y, oracle.spmv(a, x, rowstr, colidx))
a[nnz // 4] = -1.0             # protection is re-armed after the re-upload (an interior page)
libspmv.spmv_harness(y, a, x, rowstr, colidx, n)
assert libspmv.stats()["uploads"] == 3
assert np.array_equal(y, oracle.spmv(a, x, rowstr, colidx))
print("guard ok")
"""


def test_write_guard_invalidates_the_resident_matrix(tmp_path):
    """B200_SPMV_GUARD=1: the opt-in mprotect/SIGSEGV coherence device of
    libspmv/gpu.c:140-262 -- a host write to a resident matrix marks the cache
    entry stale (fingerprint check switched off to isolate the guard)."""
    import os
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    script = tmp_path / "guard.py"
    script.write_text(GUARD_SCRIPT.format(root=root))
    env = dict(os.environ, B200_SPMV_GUARD="1", B200_SPMV_VALIDATE="0")
    proc = subprocess.run([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0 and "guard ok" in proc.stdout, proc.stdout


def test_peer_exchange_spmv_single_rank_group(libspmv, oracle, npb):
    """PeerShardedSpmv on a group of one: push kernel, publish / consumed
    flags and the product, three back-to-back steps with changing x."""
    import torch
    from lilac_benchmarks_b200 import sharded
    m = npb.NpbMatrix("W")
    rm = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx)
    sh = sharded.PeerShardedSpmv(libspmv, rm, sharded.ShardLayout.build(m.n, 1), 0)
    try:
        rng = np.random.default_rng(8)
        for _ in range(3):
            x = rng.standard_normal(m.n)
            y = sh.step(torch.from_numpy(x).cuda()).cpu().numpy()
            assert np.array_equal(y, oracle.spmv(m.a, x, m.rowstr, m.colidx))
    finally:
        sh.close()


@pytest.mark.parametrize("cls", ["W", "A", "B"])
def test_caller_pinned_vectors_through_the_abi(libspmv, oracle, npb, cls):
    """The e2e headline path: x read in place over PCIe from the caller's pinned vector,
    y stored by the kernel straight into the caller's pinned vector (b200_dropin.cu);
    mixed cases too (only x pinned, only y pinned, x at an odd offset inside a pinned
    allocation)."""
    import torch
    m = npb.NpbMatrix(cls)
    rng = np.random.default_rng(5)
    for x_pinned, y_pinned, off in ((True, True, 0), (True, False, 0), (False, True, 0), (True, True, 3)):
        x = rng.standard_normal(m.n + 2 + off)
        y = np.full(m.n, np.nan)
        keep = []
        if x_pinned:
            keep.append(torch.from_numpy(x).pin_memory())
            x = keep[-1].numpy()
        if y_pinned:
            keep.append(torch.from_numpy(y).pin_memory())
            y = keep[-1].numpy()
        xv = x[off:]
        for rep in range(3):                      # repeated calls: the chunk flags are epochs
            xv[:] = rng.standard_normal(len(xv))
            libspmv.spmv_harness(y, m.a, xv, m.rowstr, m.colidx, m.n)
            assert np.array_equal(y, oracle.spmv(m.a, xv, m.rowstr, m.colidx, omp=True)), (x_pinned, y_pinned, off, rep)
    assert libspmv.stats()["uploads"] >= 1


def test_vector_that_is_only_partly_pinned_takes_the_bounce_buffer(libspmv, oracle, npb):
    """a range that is pinned only at its start must not be treated as
    device-accessible (the copy kernel would fault on the pageable tail)."""
    m = npb.NpbMatrix("W")
    rng = np.random.default_rng(6)
    buf = rng.standard_normal(m.n + 2 + 1024)
    base = buf.ctypes.data
    lo = (base + 4095) // 4096 * 4096                   # first whole page inside the buffer
    first = (lo - base) // 8
    xv = buf[first:first + m.n + 2]                     # starts exactly on the pinned page ...
    half = (m.n // 2) * 8 // 4096 * 4096                # ... but only its first half is pinned
    assert xv.ctypes.data == lo and half > 0
    assert libspmv.lib().b200_spmv_pin_host(lo, half) == 0
    try:
        for target_pinned_too in (False, True):
            ybuf = np.zeros(m.n + 1024)
            ybase = ybuf.ctypes.data
            ylo = (ybase + 4095) // 4096 * 4096
            y = ybuf[(ylo - ybase) // 8:(ylo - ybase) // 8 + m.n]
            if target_pinned_too:
                assert libspmv.lib().b200_spmv_pin_host(ylo, half) == 0
            try:
                libspmv.spmv_harness(y, m.a, xv, m.rowstr, m.colidx, m.n)
                assert np.array_equal(y, oracle.spmv(m.a, xv, m.rowstr, m.colidx))
            finally:
                if target_pinned_too:
                    assert libspmv.lib().b200_spmv_unpin_host(ylo) == 0
    finally:
        assert libspmv.lib().b200_spmv_unpin_host(lo) == 0


PIN_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
entry.load_package()
oracle = entry.load_oracle()
from lilac_benchmarks_b200 import libspmv, npb
m = npb.NpbMatrix("A")
# whole NPB CG with B200_SPMV_PIN_HOST=1: the caller's pageable vectors are registered on
# first sight and then read / written in place
gpu = npb.run_cg(m, libspmv.harness_address())
cpu = npb.run_cg(m, oracle.harness_address())
assert gpu["verified"] and np.array_equal(gpu["zeta_hist"], cpu["zeta_hist"])
assert np.array_equal(gpu["rnorm_hist"], cpu["rnorm_hist"])
rng = np.random.default_rng(1)
x = rng.standard_normal(m.n + 2); y = np.zeros(m.n)
for _ in range(3):
    x[:] = rng.standard_normal(m.n + 2)
    libspmv.spmv_harness(y, m.a, x, m.rowstr, m.colidx, m.n)
    assert np.array_equal(y, oracle.spmv(m.a, x, m.rowstr, m.colidx))
print("pin ok")
"""


def test_auto_pin_host_env(tmp_path):
    """B200_SPMV_PIN_HOST=1 through the whole NPB CG class A run."""
    import os
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    script = tmp_path / "pin.py"
    script.write_text(PIN_SCRIPT.format(root=root))
    env = dict(os.environ, B200_SPMV_PIN_HOST="1")
    proc = subprocess.run([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0 and "pin ok" in proc.stdout, proc.stdout[-3000:]


def test_guard_is_on_by_default_and_catches_in_place_edits(libspmv, oracle):
    """libspmv/gpu.c:236-262 protects the host arrays whenever a matrix is resident; so does
    this library by default.  An edit of ONE value the sampled fingerprint does not probe
    must still be seen."""
    rng = np.random.default_rng(12)
    n, ncols = 20000, 9000
    a, c, rowstr, x = make_csr(rng, n, ncols, rng.poisson(30, n))
    y = np.zeros(n)
    libspmv.spmv_harness(y, a, x, rowstr, c, n)
    up0 = libspmv.stats()["uploads"]
    a[len(a) // 2 + 12345] += 1.0                   # not on a probe position
    libspmv.spmv_harness(y, a, x, rowstr, c, n)
    assert libspmv.stats()["uploads"] == up0 + 1
    assert np.array_equal(y, oracle.spmv(a, x, rowstr, c))
    c[7777] = 1
    libspmv.spmv_harness(y, a, x, rowstr, c, n)
    assert libspmv.stats()["uploads"] == up0 + 2
    assert np.array_equal(y, oracle.spmv(a, x, rowstr, c))


def test_zero_based_offsets_are_refused():
    """rowstr[0] < 1 would read in front of the arrays; the library aborts with a
    message instead (no error channel in the ABI: libspmv/gpu.c:42-80 asserts)."""
    import os
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    code = (f"import sys; sys.path.insert(0, {root!r}); import numpy as np; import __graft_entry__ as e; "
            "e.load_package(); from lilac_benchmarks_b200 import libspmv; "
            "rs = np.array([0, 2, 4], dtype=np.int32); c = np.array([1, 2, 1, 2], dtype=np.int32); "
            "a = np.ones(4); x = np.ones(2); y = np.zeros(2); libspmv.spmv_harness(y, a, x, rs, c, 2)")
    proc = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=300)
    assert proc.returncode != 0 and "row offsets are 1-based" in proc.stderr, proc.stderr[-2000:]


def test_upload_from_device_arrays(libspmv, oracle, npb):
    """b200_spmv_upload_device: the same matrix from DEVICE arrays gives the same bits."""
    import torch
    m = npb.NpbMatrix("W")
    da, dr, dc = (torch.from_numpy(v).cuda() for v in (m.a, m.rowstr, m.colidx))
    rm = libspmv.ResidentMatrix.from_device(da.data_ptr(), dr.data_ptr(), dc.data_ptr(), m.n, keep=(da, dr, dc))
    assert rm.nnz == m.nnz and rm.ncols == m.n
    x = np.random.default_rng(2).standard_normal(m.n)
    dy = torch.empty(m.n, dtype=torch.float64, device="cuda")
    rm.exec(torch.from_numpy(x).cuda(), dy)
    assert np.array_equal(dy.cpu().numpy(), oracle.spmv(m.a, x, m.rowstr, m.colidx))


@pytest.mark.parametrize("cls,blocks", [("S", 1), ("W", 3), ("A", 1), ("B", 4)])
def test_device_makea_bit_identical_to_the_host_generator(libspmv, oracle, npb, cls, blocks):
    """include/b200_npb.h: the NPB matrix assembled on the GPU equals the host generator's
    (callers/npb/makea.c, itself pinned to the SNU CG histories and nnz self-checks) bit
    for bit -- row pointers, columns and values -- for whole matrices and row blocks."""
    import torch
    whole = npb.NpbMatrix(cls)
    na = whole.n
    cuts = [na * k // blocks for k in range(blocks + 1)]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        dm = npb.NpbDeviceMatrix(cls, lo, hi)
        a, rowstr, colidx = dm.to_host()
        ref = npb.NpbMatrix(cls, lo, hi) if blocks > 1 else whole
        assert dm.nnz == ref.nnz
        assert np.array_equal(rowstr, ref.rowstr) and np.array_equal(colidx, ref.colidx)
        assert np.array_equal(a, ref.a)
        # and it feeds the resident-matrix API without touching the host
        rm = dm.resident()
        x = np.random.default_rng(lo).standard_normal(na)
        dy = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
        rm.exec(torch.from_numpy(x).cuda(), dy)
        assert np.array_equal(dy.cpu().numpy(), oracle.spmv(ref.a, x, ref.rowstr, ref.colidx))
        rm.release()
        dm.free()


def test_class_e_row_block_generated_on_the_device(libspmv, oracle, npb):
    """BASELINE config 5, class E (na = 9 000 000, 6.3e9 nonzeros: beyond the int32 ABI as
    one matrix, cg.f cannot even be compiled for it -- CG/globals.h:80-82): a 1/64 row block
    assembled on the GPU equals the host generator's block bit for bit, and its product
    matches the oracle element-wise."""
    import torch
    cls = npb.cg_class("E")
    assert cls.na == 9000000 and cls.nonzer == 26
    lo, hi = cls.na * 17 // 64, cls.na * 18 // 64
    dm = npb.NpbDeviceMatrix("E", lo, hi, release_vectors=False)
    a, rowstr, colidx = dm.to_host()
    ref = npb.NpbMatrix("E", lo, hi)
    assert dm.nnz == ref.nnz and 0.9 * (hi - lo) * 703 < dm.nnz < 1.1 * (hi - lo) * 703   # 6 326 754 836 / 9e6 per row
    assert np.array_equal(rowstr, ref.rowstr) and np.array_equal(colidx, ref.colidx) and np.array_equal(a, ref.a)
    rm = dm.resident()
    x = np.random.default_rng(64).standard_normal(cls.na)
    dy = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
    rm.exec(torch.from_numpy(x).cuda(), dy)
    assert np.array_equal(dy.cpu().numpy(), oracle.spmv(ref.a, x, ref.rowstr, ref.colidx, omp=True))
    rm.release()
    dm.free()
    npb.lib().npb_makea_release_cache()


@pytest.mark.parametrize("mode", ["fused", "overlapped", "blocking"])
def test_peer_exchange_forms_on_the_ring_kernel_single_rank_group(libspmv, oracle, npb, mode):
    """The three forms of the sharded step on the kernel that can wait for x slices in-kernel
    (ring layout), with this GPU as its own only peer: the product pushing the slice itself
    (b200_spmv_exec_pushed), post + sliced product, exchange kernel + product.  Back-to-back
    steps with changing x and no host synchronisation in between: the epoch flags and the two
    alternating x buffers alone order the pushes against the products."""
    import os
    import torch
    from lilac_benchmarks_b200 import sharded
    m = npb.NpbMatrix("A")
    os.environ["B200_SPMV_PANEL_FMT"] = "2"
    try:
        rm = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx, kernel="panel")
    finally:
        os.environ.pop("B200_SPMV_PANEL_FMT")
    assert rm.waits_in_kernel and rm.can_push
    sh = sharded.PeerShardedSpmv(libspmv, rm, sharded.ShardLayout.build(m.n, 1), 0,
                                 overlap=mode != "blocking", fused=mode == "fused")
    assert sh.fused == (mode == "fused") and sh.overlap == (mode != "blocking")
    try:
        rng = np.random.default_rng(8)
        xs = [rng.standard_normal(m.n) for _ in range(4)]
        refs = [oracle.spmv(m.a, x, m.rowstr, m.colidx) for x in xs]
        xd = [torch.from_numpy(x).cuda() for x in xs]
        outs = [sh.step(xd[i % 4]).clone() for i in range(12)]
        torch.cuda.synchronize()
        for i, y in enumerate(outs):
            assert np.array_equal(y.cpu().numpy(), refs[i % 4]), i
    finally:
        sh.close()


AUTOPIN_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
entry.load_package()
oracle = entry.load_oracle()
from lilac_benchmarks_b200 import libspmv, npb
m = npb.NpbMatrix("A")
rng = np.random.default_rng(31)
x = rng.standard_normal(m.n + 2)
y = np.zeros(m.n)
for call in range(6):
    x[:] = rng.standard_normal(m.n + 2)
    libspmv.spmv_harness(y, m.a, x, m.rowstr, m.colidx, m.n)
    assert np.array_equal(y, oracle.spmv(m.a, x, m.rowstr, m.colidx)), call
st = libspmv.stats()
assert st["auto_pinned_calls"] == 4 and st["auto_pin_revoked"] == 0, st     # calls 3..6
keep = []
for call in range(4):                       # fresh vectors every call: bounce buffer every time
    x2 = rng.standard_normal(m.n + 2)
    y2 = np.zeros(m.n)
    keep += [x2, y2]
    libspmv.spmv_harness(y2, m.a, x2, m.rowstr, m.colidx, m.n)
    assert np.array_equal(y2, oracle.spmv(m.a, x2, m.rowstr, m.colidx))
assert libspmv.stats()["auto_pinned_calls"] == 4
libspmv.lib().b200_spmv_set_auto_pin(0)     # gives the registrations back
libspmv.spmv_harness(y, m.a, x, m.rowstr, m.colidx, m.n)
assert np.array_equal(y, oracle.spmv(m.a, x, m.rowstr, m.colidx))
assert libspmv.stats()["auto_pinned_calls"] == 4
print("autopin ok")
"""


def test_repeated_pageable_vectors_are_registered_and_stay_exact(tmp_path):
    """B200_SPMV_PIN_HOST=3: a pageable vector that comes back at the same address is registered
    on its third sighting and from then on read / written in place; results stay bit-identical,
    and vectors that are new on every call (bfs: library.cc:268-279) never qualify."""
    import os
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    script = tmp_path / "autopin.py"
    script.write_text(AUTOPIN_SCRIPT.format(root=root))
    proc = subprocess.run([sys.executable, str(script)], env=dict(os.environ, B200_SPMV_PIN_HOST="3"),
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0 and "autopin ok" in proc.stdout, proc.stdout[-3000:]


REMAP_SCRIPT = r"""
import ctypes, mmap, sys
import numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
entry.load_package()
oracle = entry.load_oracle()
from lilac_benchmarks_b200 import libspmv, npb
libc = ctypes.CDLL(None, use_errno=True)
libc.mmap.restype = ctypes.c_void_p
libc.mmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long]
libc.munmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
PROT_RW, MAP_PRIVATE, MAP_ANON, MAP_FIXED = 3, 2, 0x20, 0x10
m = npb.NpbMatrix("A")
size = ((m.n + 2) * 8 + 4095) // 4096 * 4096

def vec_at(addr):
    buf = (ctypes.c_double * (size // 8)).from_address(addr)
    return np.ctypeslib.as_array(buf)[: m.n + 2]

ax = libc.mmap(None, size, PROT_RW, MAP_PRIVATE | MAP_ANON, -1, 0)
ay = libc.mmap(None, size, PROT_RW, MAP_PRIVATE | MAP_ANON, -1, 0)
rng = np.random.default_rng(5)
x, y = vec_at(ax), vec_at(ay)[: m.n]
for call in range(5):                                   # registered from the third call on
    x[:] = rng.standard_normal(m.n + 2)
    libspmv.spmv_harness(y, m.a, x, m.rowstr, m.colidx, m.n)
    assert np.array_equal(y, oracle.spmv(m.a, x, m.rowstr, m.colidx))
assert libspmv.stats()["auto_pinned_calls"] == 3
# the owner frees both vectors and gets the SAME addresses back with fresh pages: the GPU
# mapping of the registration still points at the old ones
del x, y
for addr in (ax, ay):
    assert libc.munmap(addr, size) == 0
    got = libc.mmap(addr, size, PROT_RW, MAP_PRIVATE | MAP_ANON | MAP_FIXED, -1, 0)
    assert got == addr
x, y = vec_at(ax), vec_at(ay)[: m.n]
for call in range(3):
    x[:] = rng.standard_normal(m.n + 2)
    y[:] = -7.0
    libspmv.spmv_harness(y, m.a, x, m.rowstr, m.colidx, m.n)
    assert np.array_equal(y, oracle.spmv(m.a, x, m.rowstr, m.colidx)), call
st = libspmv.stats()
assert st["auto_pin_revoked"] >= 1, st
print("remap ok", st["auto_pin_revoked"])
"""


def test_registered_vector_remapped_by_its_owner_is_caught(tmp_path):
    """The hazard of registering memory one does not own: the owner unmaps it and the same
    addresses come back with other pages.  The per-call probes (x as the GPU sees it, y as the
    GPU wrote it, against the host's view) catch it, the registration is dropped and the call
    redone through the bounce buffer: every result stays bit-identical."""
    import os
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    script = tmp_path / "remap.py"
    script.write_text(REMAP_SCRIPT.format(root=root))
    proc = subprocess.run([sys.executable, str(script)], env=dict(os.environ, B200_SPMV_PIN_HOST="3"),
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0 and "remap ok" in proc.stdout, proc.stdout[-3000:]


OVERLAP_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
entry.load_package()
oracle = entry.load_oracle()
import torch
from lilac_benchmarks_b200 import libspmv, npb
rng = np.random.default_rng(21)
for cls, fmt in (("W", None), ("A", None), ("A", "2")):
    if fmt:
        os.environ["B200_SPMV_PANEL_FMT"] = fmt
    m = npb.NpbMatrix(cls)
    for x_pinned, y_pinned, off in ((False, False, 0), (True, True, 0), (True, False, 1), (False, True, 0)):
        x = rng.standard_normal(m.n + 2 + off)
        y = np.full(m.n, np.nan)
        keep = []
        if x_pinned:
            keep.append(torch.from_numpy(x).pin_memory()); x = keep[-1].numpy()
        if y_pinned:
            keep.append(torch.from_numpy(y).pin_memory()); y = keep[-1].numpy()
        xv = x[off:]
        for rep in range(4):                       # the chunk flags carry the call number
            xv[:] = rng.standard_normal(len(xv))
            libspmv.spmv_harness(y, m.a, xv, m.rowstr, m.colidx, m.n)
            assert np.array_equal(y, oracle.spmv(m.a, xv, m.rowstr, m.colidx, omp=True)), (cls, fmt, x_pinned, y_pinned, off, rep)
    # an in-place edit the guard cannot see (B200_SPMV_GUARD=0 here): the content checks run
    # while the product is in flight, fail, and the call is redone on a fresh upload
    up0 = libspmv.stats()["uploads"]
    m.a[0] += 1.0
    x = rng.standard_normal(m.n + 2); y = np.zeros(m.n)
    libspmv.spmv_harness(y, m.a, x, m.rowstr, m.colidx, m.n)
    assert libspmv.stats()["uploads"] == up0 + 1, (cls, fmt)
    assert np.array_equal(y, oracle.spmv(m.a, x, m.rowstr, m.colidx, omp=True)), (cls, fmt, "redo")
    libspmv.invalidate()
    os.environ.pop("B200_SPMV_PANEL_FMT", None)
# the sliced entry point on the paired layout, x NOT 16-byte aligned (cooperative slice loads)
m = npb.NpbMatrix("W")
rm = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx, kernel="panel")
assert rm.kernel_name == "panel"
xh = rng.standard_normal(m.n + 3)
xd = torch.from_numpy(xh).cuda()
yd = torch.zeros(m.n, dtype=torch.float64, device="cuda")
flags = torch.full((4,), 7, dtype=torch.int64, device="cuda")      # all four slices of epoch 7 have landed
cpr = (m.n + 3) // 4 // 2 * 2
rc = rm.exec_sliced_ptr(xd.data_ptr() + 8, yd.data_ptr(), torch.cuda.current_stream().cuda_stream,
                        flags.data_ptr(), 7, cpr, 4)
assert rc >= 1, rc
torch.cuda.synchronize()
assert np.array_equal(yd.cpu().numpy(), oracle.spmv(m.a, xh[1:], m.rowstr, m.colidx))
print("overlap ok")
"""


@pytest.mark.parametrize("chunks,flag_write", [(1, 1), (3, 1), (16, 1), (5, 0)])
def test_x_uploaded_in_chunks_while_the_product_runs(tmp_path, chunks, flag_write):
    """Drop-in path on one device: x goes up through the copy engine in `chunks` chunks on a
    second stream while the PANEL / RING kernel runs and waits per chunk (b200_dropin.cu
    run_call).  Forced on for small vectors here; pageable and pinned x / y, an x at an odd
    offset, stream memory operations and plain 8-byte copies for the flags, the redo after a
    failed content check, and the sliced entry point with a misaligned x."""
    import os
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    script = tmp_path / "overlap.py"
    script.write_text(OVERLAP_SCRIPT.format(root=root))
    env = dict(os.environ, B200_SPMV_X_OVERLAP_MIN_KB="0", B200_SPMV_X_CHUNKS=str(chunks),
               B200_SPMV_FLAG_WRITE=str(flag_write), B200_SPMV_GUARD="0",
               B200_SPMV_SMALL="0")          # the classes used here are small enough for the SMALL family
    proc = subprocess.run([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=900)
    assert proc.returncode == 0 and "overlap ok" in proc.stdout, proc.stdout[-3000:]


WATCHDOG_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
entry.load_package()
oracle = entry.load_oracle()
from lilac_benchmarks_b200 import libspmv, npb
m = npb.NpbMatrix("A")
rng = np.random.default_rng(33)
x = rng.standard_normal(m.n + 2); y = np.zeros(m.n)
for call in range(3):
    x[:] = rng.standard_normal(m.n + 2)
    libspmv.spmv_harness(y, m.a, x, m.rowstr, m.colidx, m.n)
    assert np.array_equal(y, oracle.spmv(m.a, x, m.rowstr, m.colidx)), call
st = libspmv.stats()
assert st["x_overlapped_calls"] == 1 and st["x_overlap_timeouts"] == 1, st
print("watchdog ok")
"""


def test_product_that_never_gets_its_chunk_times_out_and_is_redone(tmp_path):
    """The overlapped upload's watchdog: a chunk flag that never arrives (test hook; in the
    field: streams serialised by a profiler) ends the in-kernel wait after the timeout, the call
    is redone with x uploaded before the product, and the overlap stays off."""
    import os
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    script = tmp_path / "watchdog.py"
    script.write_text(WATCHDOG_SCRIPT.format(root=root))
    env = dict(os.environ, B200_SPMV_X_OVERLAP_MIN_KB="0", B200_SPMV_X_CHUNKS="4", B200_SPMV_X_TEST_STALL="1",
               B200_SPMV_X_TIMEOUT_MS="5", B200_SPMV_SMALL="0")
    proc = subprocess.run([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0 and "watchdog ok" in proc.stdout, proc.stdout[-3000:]


def test_vector_pinned_at_both_ends_with_a_pageable_hole_takes_the_bounce_buffer(libspmv, oracle, npb):
    """Two registrations can sit at the two ends of a vector whose middle is pageable
    (neighbouring vectors registered page by page): first and last byte are pinned and their
    device addresses are the right distance apart, yet a kernel or a copy engine reading the
    range in place would fault.  The range check walks the registrations (b200_dropin.cu,
    pinned_device_alias)."""
    m = npb.NpbMatrix("W")
    rng = np.random.default_rng(16)
    buf = rng.standard_normal(m.n + 2 + 2048)
    base = buf.ctypes.data
    lo = (base + 4095) // 4096 * 4096
    first = (lo - base) // 8
    xv = buf[first:first + m.n + 2]
    npages = (xv.nbytes + 4095) // 4096
    assert xv.ctypes.data == lo and npages >= 6
    tail = lo + (npages - 2) * 4096                      # the last two pages the vector touches
    assert libspmv.lib().b200_spmv_pin_host(lo, 2 * 4096) == 0
    assert libspmv.lib().b200_spmv_pin_host(tail, 2 * 4096) == 0
    try:
        y = np.zeros(m.n)
        for _ in range(2):
            xv[:] = rng.standard_normal(len(xv))
            libspmv.spmv_harness(y, m.a, xv, m.rowstr, m.colidx, m.n)
            assert np.array_equal(y, oracle.spmv(m.a, xv, m.rowstr, m.colidx))
    finally:
        assert libspmv.lib().b200_spmv_unpin_host(lo) == 0
        assert libspmv.lib().b200_spmv_unpin_host(tail) == 0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("sort", [True, False])
@pytest.mark.parametrize("shape", [
    dict(n=1, ncols=1, mean=1, fits=True), dict(n=33, ncols=70, mean=9, fits=True),
    dict(n=1000, ncols=1000, mean=25, fits=True), dict(n=5000, ncols=12000, mean=130, fits=True),
    dict(n=14000, ncols=14000, mean=132, fits=True),        # NPB class A's shape
    dict(n=150, ncols=9000, mean=2500, fits=True),          # few long rows
    dict(n=20000, ncols=13000, mean=1, fits=True),          # mostly empty rows, base offset below
    dict(n=3000, ncols=70000, mean=60, fits=False),         # x does not fit: falls back
    dict(n=40, ncols=5000, mean=60000, fits=False),         # rows longer than any tile: falls back
])
@pytest.mark.parametrize("env", [{}, {"B200_SPMV_SMALL_COL16": 0}, {"B200_SPMV_SMALL_CFG": 1},
                                 {"B200_SPMV_SMALL_TMA": 0, "B200_SPMV_SMALL_CFG": 1, "B200_SPMV_SMALL_COL16": 0}])
def test_small_kernel_bit_exact_any_column_order(libspmv, oracle, dtype, sort, shape, env):
    """SMALL family (whole x in shared memory, products staged per row block, one thread per
    row adds left to right): bit-identical to the reference loop for sorted, unsorted and
    repeated columns, empty rows, a row offset base > 1 -- with 16-bit or the uploaded 32-bit
    columns, either thread geometry, x by TMA or by the cooperative loop; matrices it cannot
    hold fall back to the automatic choice."""
    rng = np.random.default_rng(shape["n"] * 13 + shape["mean"])
    lens = rng.poisson(shape["mean"], shape["n"])
    lens[rng.random(shape["n"]) < 0.1] = 0
    a, c, rowstr, x = make_csr(rng, shape["n"], shape["ncols"], lens, dtype=dtype, sort=sort,
                               base=1 if shape["n"] < 1000 else 4)
    m, y = _exec_resident(libspmv, a, x, rowstr, c, "small", env)
    assert (m.kernel_name == "small") == shape["fits"], (m.kernel_name, m.ncols, m.nnz)
    assert np.array_equal(y, oracle.spmv(a, x, rowstr, c))
    m.release()


def test_small_family_is_the_automatic_choice_for_npb_class_a(libspmv, oracle, npb):
    import torch
    m = npb.NpbMatrix("A")
    rm = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx)
    assert rm.kernel_name == "small"
    rng = np.random.default_rng(3)
    x = rng.standard_normal(m.n + 2)
    dy = torch.zeros(m.n, dtype=torch.float64, device="cuda")
    for _ in range(3):                                   # back to back: dependent launches
        rm.exec(torch.from_numpy(x).cuda(), dy)
    y0 = oracle.spmv(m.a, x, m.rowstr, m.colidx)
    assert np.array_equal(dy.cpu().numpy(), y0)
    # x NOT 16-byte aligned: no TMA bulk copy, the cooperative loop brings x in
    big = torch.from_numpy(np.concatenate([[0.0], x])).cuda()
    dy.zero_()
    rm.exec(big[1:], dy)
    assert big[1:].data_ptr() % 16 == 8 and np.array_equal(dy.cpu().numpy(), y0)
    rm.release()


@pytest.mark.parametrize("cls,kernel", [("A", "auto"), ("W", "auto"), ("W", "panel"), ("B", "auto")])
def test_fused_dot_epilogue(libspmv, oracle, npb, cls, kernel):
    """b200_spmv_exec_dot (NPB conj_grad's d = p.q inside the product, cg.f:573-576): y is the
    same bits as the plain product, the row-block partials add up to dotv . y, and the fixed
    reduction order gives the same partials on every launch -- SMALL and paired PANEL kernels."""
    import torch
    m = npb.NpbMatrix(cls)
    rm = libspmv.ResidentMatrix(m.a, m.rowstr, m.colidx, kernel=kernel)
    assert rm.kernel_name == ("panel" if kernel == "panel" or cls == "B" else "small")
    assert rm.dot_partials > 0
    rng = np.random.default_rng(17)
    x, p = rng.standard_normal(m.n), rng.standard_normal(m.n)
    dx, dp = torch.from_numpy(x).cuda(), torch.from_numpy(p).cuda()
    dy = torch.zeros(m.n, dtype=torch.float64, device="cuda")
    part = rm.exec_dot(dx, dy, dp)
    y0 = oracle.spmv(m.a, x, m.rowstr, m.colidx, omp=True)
    assert np.array_equal(dy.cpu().numpy(), y0)
    ref = float(np.dot(p, y0))
    scale = float(np.dot(np.abs(p), np.abs(y0)))
    assert abs(float(part.cpu().numpy().sum()) - ref) <= 1e-13 * scale
    part2 = rm.exec_dot(dx, dy, dp)
    assert torch.equal(part, part2)
    rm.release()
