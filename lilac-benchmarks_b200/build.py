"""Build helper: compiles every native artefact in-tree with `make`.

  csrc/b200.so                 the product (nvcc, sm_100a)
  callers/libb200callers.so    NPB makea + CG driver (gcc)
  oracle/liboracle.so, oracle/_ref/*   the checker (gcc; never loaded by the product)
"""
import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
CALLERS = PKG / "callers"
ORACLE = ROOT / "oracle"

B200_SO = Path(os.environ.get("B200_SPMV_SO", CSRC / "b200.so"))   # override: experiment builds
CALLERS_SO = CALLERS / "libb200callers.so"
ORACLE_SO = ORACLE / "liboracle.so"
REF_NATIVE_SO = ORACLE / "_ref" / "native.so"
REF_TEST_BIN = ORACLE / "_ref" / "test"


def _make(directory, *targets, quiet=True):
    env = dict(os.environ)
    env.pop("CC", None)      # the image exports a CC without libgomp
    env.pop("CXX", None)
    proc = subprocess.run(["make", "-C", str(directory), *targets], env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"make -C {directory} {' '.join(targets)} failed:\n{proc.stdout}")
    if not quiet:
        print(proc.stdout)
    return proc.stdout


def build_product(quiet=True):
    return _make(CSRC, "b200.so", quiet=quiet)


def build_callers(quiet=True):
    return _make(CALLERS, "all", quiet=quiet)


def build_oracle(quiet=True):
    out = _make(ORACLE, "all", quiet=quiet)
    # when the reference tree is here, also relink its own C/C++ callers
    # (parboil spmv cpu, bfs), unchanged, against the b200 platform
    if Path("/root/reference/libspmv/native.c").exists() and B200_SO.exists():
        out += _make(ORACLE, "relink", f"B200_SO={B200_SO}", quiet=quiet)
    return out


def build_all(quiet=True):
    out = build_product(quiet) + build_callers(quiet) + build_oracle(quiet)
    return out
