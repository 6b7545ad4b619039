import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"

# The reference suite's own tolerances (tests_pytest/conftest.py:51-62 of GPUMD): FP32 accumulation
# and reduction-order noise.  Forces additionally get the atol the reference quotes for GPU(FP32)
# vs trainer output (examples/gpumd_static/check_force.m:9, "of the order of 1.0e-5").
TOL = {
    "energy": dict(rtol=1e-5, atol=1e-8),
    "force": dict(rtol=1e-4, atol=1e-5),
    "virial": dict(rtol=1e-4, atol=2e-5),
    "energy_per_atom": 1e-6,  # eV/atom, SURVEY.md 8(d)
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "all"], check=True, capture_output=True)
    return oracle_py


@pytest.fixture(scope="session")
def emu():
    """Host build of the libb200md kernel bodies (tests/emu, test infrastructure only)."""
    d = ROOT / "tests" / "emu"
    lib = d / "libb2emu.so"
    srcs = [d / "emu.cpp", ROOT / "gpumd_b200/csrc/b2_nep_model.cpp"]
    deps = srcs + list((ROOT / "gpumd_b200/csrc").glob("*.cuh")) + list((ROOT / "gpumd_b200/csrc").glob("*.h"))
    if not lib.exists() or any(p.stat().st_mtime > lib.stat().st_mtime for p in deps):
        subprocess.run(
            ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
             "-ffp-contract=fast", "-o", str(lib)] + [str(s) for s in srcs], check=True)
    from emu_py import Emu
    return Emu(str(lib))


@pytest.fixture(scope="session")
def b200md_lib():
    from gpumd_b200 import build
    return build.build_lib()


def assert_close(a, b, rtol, atol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    worst = np.argmax(err)
    assert err.flat[worst] <= 0, (
        f"{what}: |{a.flat[worst]} - {b.flat[worst]}| = {abs(a.flat[worst] - b.flat[worst]):.3e} "
        f"> atol {atol} + rtol {rtol}*|ref|")
