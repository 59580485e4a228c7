"""SparseBench BiCG and PageRank callers (C restatements under callers/) on the
CPU with the oracle as backend, and -- marked gpu -- the same runs through
libb200-spmv, whose histories must be bit-identical because every product is."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def callers(built):
    from lilac_benchmarks_b200 import callers as mod
    return mod


@pytest.fixture(scope="module")
def gen(built):
    from lilac_benchmarks_b200 import gen as mod
    return mod


def test_crsmat_generator_matches_big_gen_statistics(gen):
    a, colidx, rowstr, n = gen.crsmat(20)
    assert n == 8000 and rowstr[0] == 1 and rowstr[-1] == len(a) + 1
    lens = np.diff(rowstr)
    assert lens.min() >= 1 and 4.5 < lens.mean() < 5.4        # max(1, int(gauss(5, 4)))
    assert 0.18 < (lens == 1).mean() < 0.27
    assert np.all(a >= 0) and colidx.min() >= 1 and colidx.max() <= n
    a2, c2, r2, _ = gen.crsmat(20)
    assert np.array_equal(a, a2) and np.array_equal(colidx, c2)   # seeded


def test_crs_file_round_trip_and_trailing_lines(callers, gen, tmp_path):
    a, colidx, rowstr, n = gen.crsmat(6)
    path = tmp_path / "crsmat6u"
    callers.write_crs(path, a, rowstr, colidx, extra_lines=[(1, 9.0), (2, 9.0)])   # big_gen's surplus
    a2, rowstr2, colidx2 = callers.read_crs(path)
    assert np.array_equal(rowstr2, rowstr) and np.array_equal(colidx2, colidx)
    assert np.allclose(a2, a, rtol=0, atol=1e-16)


def test_bicg_on_cpu_backend(callers, gen, oracle):
    a, colidx, rowstr, n = gen.crsmat(12)
    res = callers.bicg(a, rowstr, colidx, oracle.harness_address())
    assert res["matprod_calls"] == 1 + 2 * (abs(res["its"]) - (1 if res["its"] > 0 else 0))
    assert res["hist"][0] == pytest.approx(np.sqrt(n))          # x0 = 0, rhs = 1  =>  ||r0|| = sqrt(n)
    # first residual recomputed independently: r1 = r0 - alpha A r0 with the fork's A-for-A^T product
    r0 = -np.ones(n)
    ar0 = oracle.spmv(a, r0, rowstr, colidx)
    alpha = (r0 @ r0) / (r0 @ ar0)
    assert res["hist"][1] == pytest.approx(np.linalg.norm(r0 - alpha * ar0), rel=1e-12)


def test_pagerank_on_cpu_backend(callers, gen, oracle):
    a, colidx, rowstr, x0 = gen.powerlaw_graph(n=4096, seed=3)
    x, err, _ = callers.pagerank(a, rowstr, colidx, x0, oracle.harness_address(), iters=60)
    # d*M is column-stochastic times d, so sum(x) stays 1 up to rounding; the iteration contracts
    assert x.sum() == pytest.approx(1.0, rel=1e-9) and np.all(x > 0)
    x2, err2, _ = callers.pagerank(a, rowstr, colidx, x, oracle.harness_address(), iters=1)
    assert err2 <= err and err < 1e-6
    # one step by hand
    y = oracle.spmv(a, x0, rowstr, colidx) + 0.15 * x0.sum() / len(x0)
    x1, _, _ = callers.pagerank(a, rowstr, colidx, x0, oracle.harness_address(), iters=1)
    assert np.allclose(x1, y, rtol=1e-12, atol=0)     # numpy sums pairwise, the caller left to right


def test_mtx_loader_normalises_columns(callers, tmp_path):
    path = tmp_path / "g.mtx"
    path.write_text("%%MatrixMarket matrix coordinate pattern general\n% c\n4 4 6\n"
                    "1 2\n3 2\n2 1\n4 3\n1 4\n2 4\n")
    a, rowstr, colidx = callers.load_mtx(path, d=0.85)
    assert list(rowstr) == [1, 3, 5, 6, 7] and list(colidx) == [2, 4, 1, 4, 2, 3]
    assert np.allclose(a, 0.85 * np.array([0.5, 0.5, 1.0, 0.5, 0.5, 1.0]))


@pytest.mark.gpu
def test_bicg_history_bit_identical_on_gpu(callers, gen, oracle, libspmv):
    libspmv.invalidate()
    a, colidx, rowstr, n = gen.crsmat(30)
    cpu = callers.bicg(a, rowstr, colidx, oracle.harness_address())
    gpu = callers.bicg(a, rowstr, colidx, libspmv.harness_address())
    assert gpu["its"] == cpu["its"] and np.array_equal(gpu["hist"], cpu["hist"])
    assert np.array_equal(gpu["x"], cpu["x"])


@pytest.mark.gpu
def test_pagerank_bit_identical_on_gpu(callers, gen, oracle, libspmv):
    libspmv.invalidate()
    a, colidx, rowstr, x0 = gen.powerlaw_graph(n=1 << 16, seed=5, cap=200)   # below the long-row cap
    cpu, e0, _ = callers.pagerank(a, rowstr, colidx, x0, oracle.harness_address(), iters=30)
    gpu, e1, _ = callers.pagerank(a, rowstr, colidx, x0, libspmv.harness_address(), iters=30)
    lens = np.diff(rowstr)
    short = lens <= max(64, 4 * lens.mean())
    assert np.allclose(gpu, cpu, rtol=1e-12, atol=0)
    if short.all():
        assert np.array_equal(gpu, cpu) and e0 == e1


def _split_by_cap(rowstr, cap):
    lens = np.diff(rowstr)
    return lens <= cap


@pytest.mark.gpu
def test_crsmat170u_full_size_elementwise(gen, oracle, libspmv):
    """BASELINE config 3 at size: big_gen.py's 170^3 matrix (4.9 M rows, 24 M nonzeros,
    unsorted stream windows) through spmv_harness_, every y element against the OpenMP
    oracle.  Rows are short (<= 25), so every row is inside the order-preserving tiles:
    bit-exact."""
    libspmv.invalidate()
    a, colidx, rowstr, n = gen.crsmat(170)
    assert n == 170 ** 3 and np.diff(rowstr).max() <= 64
    rng = np.random.default_rng(170)
    for x in (np.ones(n), rng.standard_normal(n)):
        y = np.full(n, np.nan)
        libspmv.spmv_harness(y, a, x, rowstr, colidx, n)
        assert np.array_equal(y, oracle.spmv(a, x, rowstr, colidx, omp=True))
    libspmv.invalidate()


@pytest.mark.gpu
def test_powerlaw_2_22_full_size_elementwise(gen, oracle, libspmv):
    """BASELINE config 4 at size: the 2^22-vertex power-law graph with NO cap on the row
    lengths (rows up to 65536 entries go through the nnz-split long-row path).  Rows inside
    the order-preserving tiles: bit-exact; the long rows re-order the sum and are held to
    the north star's 1e-12 relative (pagerank values are positive: no cancellation)."""
    import torch
    libspmv.invalidate()
    a, colidx, rowstr, x0 = gen.powerlaw_graph()
    n = len(rowstr) - 1
    assert n == 1 << 22 and np.diff(rowstr).max() >= 30000
    rm = libspmv.ResidentMatrix(a, rowstr, colidx)
    dy = torch.empty(n, dtype=torch.float64, device="cuda")
    rm.exec(torch.from_numpy(x0).cuda(), dy)
    y = dy.cpu().numpy()
    y0 = oracle.spmv(a, x0, rowstr, colidx, omp=True)
    lens = np.diff(rowstr)
    tiled = lens <= 32                      # certainly below the cap (2.5 x mean = 40)
    assert np.array_equal(y[tiled], y0[tiled])
    assert np.all(np.abs(y - y0) <= 1e-12 * np.abs(y0))
    # and one pagerank step through the ABI with the C caller
    x1_cpu, _, _ = callers_pagerank(a, rowstr, colidx, x0, oracle.harness_address())
    x1_gpu, _, _ = callers_pagerank(a, rowstr, colidx, x0, libspmv.harness_address())
    assert np.all(np.abs(x1_gpu - x1_cpu) <= 1e-12 * np.abs(x1_cpu))
    rm.release()
    libspmv.invalidate()


def callers_pagerank(a, rowstr, colidx, x0, addr):
    from lilac_benchmarks_b200 import callers
    return callers.pagerank(a, rowstr, colidx, x0, addr, iters=1)
