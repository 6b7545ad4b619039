"""The C-ABI library builds for sm_100a, loads, and exports exactly what include/b200md.h declares.
No compute calls here (no GPU in this container): creation must fail LOUDLY, never fall back."""
import ctypes as C
import re

import pytest

from conftest import GOLDEN, ROOT, has_gpu


def header_symbols():
    text = (ROOT / "include" / "b200md.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200md_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(b200md_lib):
    from gpumd_b200 import lib
    names = header_symbols()
    assert len(names) >= 20
    assert sorted(lib.SIGNATURES) == names, "python binding table and header disagree"
    L = C.CDLL(str(b200md_lib))
    for n in names:
        assert hasattr(L, n), f"{n} declared in b200md.h but not exported"


def test_mgpu_library_exports_every_declared_symbol(b200md_lib):
    """include/b200md_mgpu.h <-> libb200md_mgpu.so (C++ / CUDA / NCCL domain module) <-> gpumd_b200/mgpu.py."""
    from gpumd_b200 import build
    text = (ROOT / "include" / "b200md_mgpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(b200md_mgpu_[a-z_0-9]+)\s*\(", text)))
    assert len(names) >= 14
    so = build.build_mgpu()
    C.CDLL(str(b200md_lib), mode=C.RTLD_GLOBAL)
    L = C.CDLL(str(so))
    for n in names:
        assert hasattr(L, n), f"{n} declared in b200md_mgpu.h but not exported"


def test_library_targets_sm_100a(b200md_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(b200md_lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(b200md_lib):
    from gpumd_b200 import lib
    L = lib.load()
    h = C.c_void_p()
    rc = L.b200md_nep_create(str(GOLDEN / "nep_PbTe.txt").encode(), 512, C.byref(h))
    assert rc == 3 and b"no CPU fallback" in L.b200md_last_error()
    rc = L.b200md_lj_create(str(GOLDEN / "lj_Ar_10A.txt").encode(), 512, C.byref(h))
    assert rc == 3
    from gpumd_b200 import engine
    with pytest.raises(lib.B200mdError):
        engine.NEP(GOLDEN / "nep_PbTe.txt", 512)


def test_standalone_driver_refuses_to_run_without_a_gpu(tmp_path):
    """The C++ driver has no CPU path either: on a machine without a CUDA device it must say so and
    exit 1 before touching any input file (the reference's error convention, error.cuh:22-62)."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gpumd_b200 import build
    exe = build.build_host()
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 1
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr
