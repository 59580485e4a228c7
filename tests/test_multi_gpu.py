"""Parity on SEVERAL GPUs (skipped on a box with one): the peer-memory exchange between
two processes, and the ABI mode of the drop-in symbols -- one process, B200_SPMV_DEVICES --
including the reference's own unit-test binary.  Element-wise against the oracle."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _device_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


needs_two = pytest.mark.skipif("_device_count() < 2", reason="needs two GPUs")


@needs_two
@pytest.mark.parametrize("cls", ["A"])
def test_two_rank_peer_exchange_elementwise(cls):
    """tests/peer_rank_script.py under torch.distributed.run with one rank per GPU."""
    env = dict(os.environ, OMP_NUM_THREADS="8")
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29611",
                           str(ROOT / "tests" / "peer_rank_script.py"), cls],
                          env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert proc.returncode == 0 and "peer ok world=2" in proc.stdout, proc.stdout[-4000:]


ABI_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
entry.load_package()
oracle = entry.load_oracle()
from lilac_benchmarks_b200 import libspmv, npb
import torch
# NPB class B through spmv_harness_: 13.7 M nonzeros, above the multi-device threshold
m = npb.NpbMatrix("B")
rng = np.random.default_rng(3)
for kind in ("pageable", "pinned"):
    for trial in range(3):
        x = rng.standard_normal(m.n + 2)
        y = np.full(m.n, np.nan)
        if kind == "pinned":
            xt, yt = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
            x, y = xt.numpy(), yt.numpy()
        libspmv.spmv_harness(y, m.a, x, m.rowstr, m.colidx, m.n)
        assert np.array_equal(y, oracle.spmv(m.a, x, m.rowstr, m.colidx, omp=True)), (kind, trial)
assert libspmv.devices_in_use() == {ndev}, libspmv.devices_in_use()
assert libspmv.stats()["uploads"] == 1
# whole NPB CG class A (1.85 M nonzeros: forced onto the devices with MULTI_MIN_NNZ=0 below)
res = npb.run_cg(npb.NpbMatrix("A"), libspmv.harness_address())
assert res["verified"], res["zeta"]
cpu = npb.run_cg(npb.NpbMatrix("A"), oracle.harness_address())
assert np.array_equal(res["zeta_hist"], cpu["zeta_hist"])
# unsorted / skewed / empty rows, fp32, more devices than some blocks have rows
lens = rng.poisson(6, 5000); lens[::7] = 0; lens[11] = 3000
rowstr = np.empty(5001, dtype=np.int32); rowstr[0] = 1; rowstr[1:] = 1 + np.cumsum(lens)
nnz = int(lens.sum())
colidx = rng.integers(1, 4000, nnz).astype(np.int32)
for dt in (np.float64, np.float32):
    a = (rng.random(nnz) + 0.1).astype(dt)
    x = (rng.random(4000) + 0.1).astype(dt)
    y = np.zeros(5000, dtype=dt)
    (libspmv.f_spmv_harness if dt == np.float32 else libspmv.spmv_harness)(y, a, x, rowstr, colidx, 5000)
    y0 = oracle.spmv(a, x, rowstr, colidx)
    short = lens <= 200
    assert np.array_equal(y[short], y0[short])
    assert np.allclose(y, y0, rtol=1e-12 if dt == np.float64 else 2e-5, atol=0)
print("abi multi ok")
"""


@needs_two
def test_abi_mode_spreads_the_drop_in_symbols_over_two_devices(tmp_path):
    """B200_SPMV_DEVICES=0,1: same symbols, same callers, row blocks on two GPUs,
    x pulled over two PCIe links and exchanged over NVLink, y blocks written back."""
    script = tmp_path / "abi_multi.py"
    script.write_text(ABI_SCRIPT.format(root=str(ROOT), ndev=2))
    env = dict(os.environ, B200_SPMV_DEVICES="0,1", B200_SPMV_MULTI_MIN_NNZ="0")
    proc = subprocess.run([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=900)
    assert proc.returncode == 0 and "abi multi ok" in proc.stdout, proc.stdout[-4000:]


@needs_two
def test_reference_test_binary_on_two_devices(libspmv, oracle):
    """libspmv/test.cpp (the reference's own unit test) against b200.so spread over two
    devices: 3 rows, so one device gets two rows and the other one."""
    if not oracle.REF_TEST_BIN.exists():
        pytest.skip("oracle/_ref/test not built")
    env = dict(os.environ, B200_SPMV_DEVICES="0,1", B200_SPMV_MULTI_MIN_NNZ="0", B200_SPMV_VERBOSE="1")
    proc = subprocess.run([str(oracle.REF_TEST_BIN), "b200"], cwd=str(libspmv.B200_SO.parent), env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert "success!" in proc.stderr, proc.stderr
    assert "spread over 2 devices" in proc.stderr, proc.stderr


def test_abi_mode_with_one_listed_device_is_the_single_device_path(tmp_path):
    """B200_SPMV_DEVICES=0 (runs on every box): the device list code path with one entry."""
    script = tmp_path / "abi_one.py"
    script.write_text(ABI_SCRIPT.format(root=str(ROOT), ndev=1))
    env = dict(os.environ, B200_SPMV_DEVICES="0", B200_SPMV_MULTI_MIN_NNZ="0")
    proc = subprocess.run([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=900)
    assert proc.returncode == 0 and "abi multi ok" in proc.stdout, proc.stdout[-4000:]


@needs_two
def test_npb_cg_binary_drives_two_gpus_unchanged(libspmv):
    """The compiled NPB CG caller (callers/npb: the C restatement of NPB3.3.1/CG/cg.f, the
    reference's primary caller of the ABI) dlopens libb200-spmv and, with nothing but
    B200_SPMV_DEVICES in its environment, runs class B on two GPUs: zeta verifies
    (cg.f:363-368; exit status 0) and the library reports two devices."""
    cg = ROOT / "lilac-benchmarks_b200" / "callers" / "cg"
    if not cg.exists():
        pytest.skip("callers/cg not built")
    env = dict(os.environ, B200_SPMV_DEVICES="0,1", B200_SPMV_VERBOSE="1")
    proc = subprocess.run([str(cg), "B", str(libspmv.B200_SO)], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "spread over 2 devices" in proc.stderr, proc.stderr[-2000:]
    assert "SUCCESSFUL" in proc.stdout.upper(), proc.stdout[-2000:]
