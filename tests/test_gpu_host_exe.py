"""The C++ host layer (gpumd_b200/host: Potential/Force/Ensemble adapters + the standalone `b200md`
driver) end to end: GPUMD's own input files in, thermo.out in the reference's format out, compared
with the thermo.out the unmodified reference gpumd wrote for the same inputs on a B200."""
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN
from gpumd_b200.structures import init_velocities, nep_type_order, rocksalt_pbte, write_xyz

pytestmark = pytest.mark.gpu


def read_thermo(path):
    return np.array([ln.split() for ln in open(path) if not ln.startswith("#")], dtype=np.float64)


def test_b200md_executable_reproduces_reference_thermo(tmp_path):
    from gpumd_b200 import build
    build.build_lib()
    exe = build.build_host()
    s = rocksalt_pbte(20, rattle=0.02, seed=1)  # 64 000 atoms: the md_pbte reference case
    vel = init_velocities(s["mass"], 300.0, seed=42)
    write_xyz(tmp_path / "model.xyz", s, nep_type_order(GOLDEN / "nep_PbTe.txt"), vel)
    shutil.copyfile(GOLDEN / "nep_PbTe.txt", tmp_path / "potential.txt")
    (tmp_path / "run.in").write_text(
        "potential potential.txt\nensemble nve\ntime_step 1\ndump_thermo 10\nrun 200\n")
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Speed of this run" in r.stdout
    mine = read_thermo(tmp_path / "thermo.out")
    ref = read_thermo(GOLDEN / "refgpu_md_pbte_thermo.out")
    assert mine.shape == ref.shape == (20, 18)
    n = 64000
    for k in range(20):
        tol = 2e-5 * (1 + k)
        assert abs(mine[k, 0] - ref[k, 0]) < tol * 3000
        assert abs(mine[k, 2] - ref[k, 2]) < tol * abs(ref[k, 2])
        assert np.allclose(mine[k, 3:6], ref[k, 3:6], rtol=1e-3, atol=2e-3)
        assert np.array_equal(mine[k, 9:], ref[k, 9:])  # the box columns
    assert abs(mine[0, 0] - ref[0, 0]) < 2e-3 and abs(mine[0, 2] - ref[0, 2]) / n < 2e-7


def test_b200md_rejects_unknown_keyword(tmp_path):
    from gpumd_b200 import build
    exe = build.build_host()
    s = rocksalt_pbte(4, rattle=0.02, seed=1)
    write_xyz(tmp_path / "model.xyz", s, nep_type_order(GOLDEN / "nep_PbTe.txt"))
    shutil.copyfile(GOLDEN / "nep_PbTe.txt", tmp_path / "potential.txt")
    (tmp_path / "run.in").write_text("potential potential.txt\nensemble npt_scr 300 300 100 0 0 0 100 100 100 1000\n")
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Input Error" in r.stderr
