"""The C++ host layer (gpumd_b200/host: Potential/Force/Ensemble adapters + the standalone `b200md`
driver) end to end: GPUMD's own input files in, thermo.out in the reference's format out, compared
with the thermo.out the unmodified reference gpumd wrote for the same inputs on a B200."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN
from gpumd_b200.structures import (diamond, fcc, init_velocities, nep_type_order, rocksalt_pbte,
                                   write_xyz)

pytestmark = pytest.mark.gpu


def read_thermo(path):
    return np.array([ln.split() for ln in open(path) if not ln.startswith("#")], dtype=np.float64)


HOST_CASES = {
    # case: (structure, symbols, potential file, T0, run.in ensemble line, time step fs)
    "md_pbte": (lambda: rocksalt_pbte(20, rattle=0.02, seed=1), None, "nep_PbTe.txt", 300.0, "nve", 1),
    "md_pbte_bdp": (lambda: rocksalt_pbte(20, rattle=0.02, seed=1), None, "nep_PbTe.txt", 300.0,
                    "nvt_bdp 300 300 100", 1),
    "md_pbte_nhc": (lambda: rocksalt_pbte(20, rattle=0.02, seed=1), None, "nep_PbTe.txt", 300.0,
                    "nvt_nhc 300 300 100", 1),
    "md_si": (lambda: diamond(20, a=5.431, rattle=0.0, seed=1), ["Si"], "tersoff_Si_1989.txt", 300.0, "nve", 1),
    "md_lj": (lambda: fcc(25, 5.30, rattle=0.0, seed=1), ["Ar"], "lj_Ar_10A.txt", 80.0, "nve", 5),
}


@pytest.mark.parametrize("case", list(HOST_CASES))
def test_b200md_executable_reproduces_reference_thermo(tmp_path, case):
    """run.in / model.xyz / potential file in, thermo.out out, for NEP (NVE and two thermostats),
    Tersoff-1989 and LJ -- against the reference gpumd's thermo.out for the same inputs."""
    from gpumd_b200 import build
    build.build_lib()
    exe = build.build_host()
    make, symbols, potfile, T0, ensemble, dt = HOST_CASES[case]
    s = make()
    n = s["type"].shape[0]
    vel = init_velocities(s["mass"], T0, seed=42)
    write_xyz(tmp_path / "model.xyz", s, symbols or nep_type_order(GOLDEN / potfile), vel)
    shutil.copyfile(GOLDEN / potfile, tmp_path / "potential.txt")
    (tmp_path / "run.in").write_text(
        f"potential potential.txt\nensemble {ensemble}\ntime_step {dt}\ndump_thermo 10\nrun 200\n")
    # the fixture was written by a -DDEBUG build of the reference: std::mt19937(12345678) for nvt_bdp
    env = dict(os.environ, B200MD_DEBUG_SEED="1")
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Speed of this run" in r.stdout
    mine = read_thermo(tmp_path / "thermo.out")
    ref = read_thermo(GOLDEN / f"refgpu_{case}_thermo.out")
    assert mine.shape == ref.shape == (20, 18)
    for k in range(20):
        tol = 2e-5 * (1 + k)
        assert abs(mine[k, 0] - ref[k, 0]) < tol * 10 * T0
        assert abs(mine[k, 2] - ref[k, 2]) < tol * abs(ref[k, 2])
        assert np.allclose(mine[k, 3:6], ref[k, 3:6], rtol=1e-3, atol=2e-3)
        assert np.array_equal(mine[k, 9:], ref[k, 9:])  # the box columns
    assert abs(mine[0, 0] - ref[0, 0]) < 2e-3 and abs(mine[0, 2] - ref[0, 2]) / n < 2e-7


def test_b200md_ensemble_is_built_at_run_with_the_final_time_step(tmp_path):
    """`ensemble nvt_nhc ...` BEFORE `time_step 0.5` (the usual run.in order): the chain masses
    Q = kT (dt Tc)^2 must use the 0.5 fs step in force at `run` (Integrate::initialize,
    integrate.cu:76-280), and a second `run` starts from a fresh chain.  Checked against the Python
    mirror driving the same library with the thermostat constructed from dt = 0.5 fs."""
    from gpumd_b200 import build, engine
    from test_gpu_md import run_nve
    exe = build.build_host()
    s = rocksalt_pbte(8, rattle=0.02, seed=1)
    vel = init_velocities(s["mass"], 300.0, seed=42)
    write_xyz(tmp_path / "model.xyz", s, nep_type_order(GOLDEN / "nep_PbTe.txt"), vel)
    shutil.copyfile(GOLDEN / "nep_PbTe.txt", tmp_path / "potential.txt")
    (tmp_path / "run.in").write_text(
        "potential potential.txt\nensemble nvt_nhc 300 300 100\ntime_step 0.5\ndump_thermo 10\nrun 100\n")
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    mine = read_thermo(tmp_path / "thermo.out")
    _, _, rows = run_nve(engine, s, GOLDEN / "nep_PbTe.txt", 100, 0.5, 300.0, seed=42, every=10,
                         ensemble="nvt_nhc")
    ref = rows[1:-1]
    assert mine.shape[0] == ref.shape[0] == 10
    # model.xyz carries 17 significant digits, so both runs start from the same state; the same
    # kernels in the same order give the same trajectory up to the text round trip
    assert np.allclose(mine[:, 0], ref[:, 0], rtol=1e-7, atol=0)
    # with the 1 fs default (the bug) T(100 steps) differs in the 4th digit
    _, _, wrong = run_nve(engine, s, GOLDEN / "nep_PbTe.txt", 100, 0.5, 300.0, seed=42, every=10,
                          ensemble="nvt_nhc_dt1")
    assert abs(wrong[-2, 0] - ref[-1, 0]) > 1e-5 * ref[-1, 0]


def test_b200md_compute_hac_matches_reference(tmp_path):
    """compute_hac through the C++ driver against hac.out of the unmodified reference gpumd for the same
    run.in / model.xyz (8000 Si atoms, Tersoff-1989 FP64, nvt_ber 100 steps then NVE 400 steps with
    `compute_hac 2 50 1`; tests/golden/refgpu_hac_si.out, scripts/run_reference_gpumd.py)."""
    from gpumd_b200 import build
    if not (GOLDEN / "refgpu_hac_si.out").exists():
        pytest.skip("refgpu_hac_si.out not generated yet")
    exe = build.build_host()
    s = diamond(10, a=5.431, rattle=0.08, seed=24)
    vel = init_velocities(s["mass"], 300.0, seed=42)
    write_xyz(tmp_path / "model.xyz", s, ["Si"], vel)
    shutil.copyfile(GOLDEN / "tersoff_Si_1989.txt", tmp_path / "potential.txt")
    (tmp_path / "run.in").write_text(
        "potential potential.txt\nensemble nvt_ber 300 300 100\ntime_step 1\nrun 100\n"
        "ensemble nve\ncompute_hac 2 50 1\nrun 400\n")
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    mine = np.loadtxt(tmp_path / "hac.out")
    ref = np.loadtxt(GOLDEN / "refgpu_hac_si.out")
    assert mine.shape == ref.shape == (50, 11)
    assert np.array_equal(mine[:, 0], ref[:, 0])  # correlation times
    # FP64 force field, same integrator: 500 steps of chaotic divergence from 1e-13 stay far below this
    scale = np.abs(ref[:, 1:6]).max()
    assert np.allclose(mine[:, 1:6], ref[:, 1:6], rtol=1e-5, atol=1e-6 * scale)
    assert np.allclose(mine[:, 6:], ref[:, 6:], rtol=1e-5, atol=1e-6 * np.abs(ref[:, 6:]).max())


def test_b200md_rejects_unknown_keyword(tmp_path):
    from gpumd_b200 import build
    exe = build.build_host()
    s = rocksalt_pbte(4, rattle=0.02, seed=1)
    write_xyz(tmp_path / "model.xyz", s, nep_type_order(GOLDEN / "nep_PbTe.txt"))
    shutil.copyfile(GOLDEN / "nep_PbTe.txt", tmp_path / "potential.txt")
    (tmp_path / "run.in").write_text("potential potential.txt\nensemble npt_scr 300 300 100 0 0 0 100 100 100 1000\n")
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Input Error" in r.stderr


def test_b200md_dump_xyz_single_point_matches_reference(tmp_path):
    """The reference's own single-point recipe (time_step 0, dump_xyz ... precision double force):
    the extended-XYZ frame written by b200md must parse with the same reader and carry the energy,
    virial and forces the reference gpumd wrote for the same input (refgpu_sp_pbte.npz)."""
    from gpumd_b200 import build
    from gpumd_b200.structures import read_xyz
    exe = build.build_host()
    d = np.load(GOLDEN / "refgpu_sp_pbte.npz")
    sym = nep_type_order(GOLDEN / "nep_PbTe.txt")
    s = dict(type=d["type"], pos=d["pos"], h=d["h"], pbc=d["pbc"],
             mass=np.where(d["type"] == sym.index("Pb"), 207.2, 127.6))
    n = s["type"].shape[0]
    write_xyz(tmp_path / "model.xyz", s, sym)
    shutil.copyfile(GOLDEN / "nep_PbTe.txt", tmp_path / "potential.txt")
    (tmp_path / "run.in").write_text(
        "potential potential.txt\nvelocity 1\nensemble nve\ntime_step 0\n"
        "dump_xyz 1 dump.xyz precision double force potential\ndump_restart 1\nrun 1\n")
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = read_xyz(tmp_path / "dump.xyz", sym)
    assert abs(out["energy"] - float(d["energy"])) / n < 1e-6
    assert np.allclose(out["forces"], d["force"], rtol=1e-4, atol=1e-5)
    assert np.allclose(out["virial"], d["virial"], rtol=1e-4, atol=2e-3)
    L = np.diag(d["h"].reshape(3, 3))[:, None]  # orthogonal box: wrapped positions are pos mod L
    wrapped = np.mod(d["pos"], L)
    assert np.allclose(out["pos"], wrapped, rtol=0, atol=1e-9)
    # restart.xyz is a model.xyz (%g columns, like the reference's): positions and masses round-trip
    rs = read_xyz(tmp_path / "restart.xyz", sym)
    assert np.array_equal(rs["type"], s["type"])
    assert np.allclose(rs["pos"], wrapped, rtol=1e-5, atol=1e-5)  # %g keeps 6 significant digits


def test_b200md_replicate(tmp_path):
    """replicate 2 1 2 before the potential: 4x the atoms, lattice vectors a and c doubled, and for a
    periodic crystal 4x the energy of the unreplicated cell."""
    from gpumd_b200 import build
    exe = build.build_host()
    s = rocksalt_pbte(8, rattle=0.03, seed=3)  # 52.6 A: a large box already
    sym = nep_type_order(GOLDEN / "nep_PbTe.txt")
    write_xyz(tmp_path / "model.xyz", s, sym)
    shutil.copyfile(GOLDEN / "nep_PbTe.txt", tmp_path / "potential.txt")
    energies = []
    for rep in ("", "replicate 2 1 2\n"):
        for f in ("thermo.out",):
            (tmp_path / f).unlink(missing_ok=True)
        (tmp_path / "run.in").write_text(
            rep + "potential potential.txt\nvelocity 1\nensemble nve\ntime_step 0\ndump_thermo 1\nrun 1\n")
        r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        th = read_thermo(tmp_path / "thermo.out")
        energies.append(th[0, 2])
        if rep:
            assert "Number of atoms is %d" % (4 * s["type"].shape[0]) in r.stdout
            L = s["h"][0]
            assert np.allclose(th[0, 9:], [2 * L, 0, 0, 0, L, 0, 0, 0, 2 * L])
    assert abs(energies[1] - 4 * energies[0]) < 1e-5 * abs(energies[1])


def test_b200md_reproduces_the_references_carbon_golden(tmp_path):
    """tests/gpumd/carbon of the reference: its own regression golden for the large-box NEP path --
    64 000 carbon atoms (C_2022_NEP4, 4- and 5-body terms), `velocity 300` from the unseeded rand()
    stream, NVE, 100 steps, thermo1.out every 10 steps, neighbor1.out maxima at step 0.  The whole chain
    (model.xyz reader, velocity initialisation incl. angular-momentum removal, force field, integrator,
    thermo writer) must land on the checked-in numbers as far as they do not depend on the platform's
    rand() stream."""
    from gpumd_b200 import build
    exe = build.build_host()
    d = np.load(GOLDEN / "carbon_model.npz")
    pos = d["pos"]
    with open(tmp_path / "model.xyz", "w") as f:
        f.write(f"{pos.shape[0]}\n{str(d['header'])}\n")
        for x, y, z in pos:
            f.write(f"C {x:.11f} {y:.11f} {z:.11f}\n")
    shutil.copyfile(GOLDEN / "nep_C_2022_NEP4.txt", tmp_path / "potential.txt")
    (tmp_path / "run.in").write_text(
        "potential potential.txt\nvelocity 300\ntime_step 1.0\nensemble nve\ndump_thermo 10\nrun 100\n")
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    mine = read_thermo(tmp_path / "thermo.out")
    ref = read_thermo(GOLDEN / "carbon_thermo1.out")
    assert mine.shape == ref.shape == (10, 18)
    n = pos.shape[0]
    # The golden's random velocities cannot be reproduced bit for bit (it was written on a platform
    # whose rand() stream differs: T(step 10) = 300.70 K there, 300.58 K with glibc), so the pin is on
    # what does not depend on the stream: both runs start from the same positions with the kinetic
    # energy scaled to exactly 300 K, hence the same conserved total energy -- which checks the force
    # field's energy at the golden configuration, the velocity scaling and the integrator -- and the
    # same thermodynamic state within finite-size fluctuations (1/sqrt(N) = 0.4 %).
    e_mine = mine[:, 1] + mine[:, 2]
    e_ref = ref[:, 1] + ref[:, 2]
    # NVE conservation.  The total energy does not drift here, it WANDERS by ~1e-6 eV/atom from
    # output to output: FP32 force noise.  Measured on one B200 from identical positions and
    # velocities (scripts/diag_r02.py, profiles/r02_diag_carbon.json), deviation from the first
    # row in ueV/atom over the 10 rows: this library rms 0.85 / max 2.08 (bit-identical with the
    # SIMT or the tensor-core hidden layer, with rsqrtf or the exact 1/sqrt, +-0.02 with the round-1
    # kernels -- so neither 3xTF32 nor rsqrt is the cause); the unmodified reference gpumd rms 0.68 /
    # max 1.29; the checked-in golden itself rms 0.93 / max 1.99.  Three samples of the same noise.
    # Bounds: rms and extreme within 2x of the golden's own (3.98e-6 eV/atom; round 1 asserted 5e-6).
    w_mine, w_ref = (e_mine - e_mine[0]) / n, (e_ref - e_ref[0]) / n
    assert np.sqrt(np.mean(w_mine ** 2)) < 2.0 * np.sqrt(np.mean(w_ref ** 2)), (w_mine, w_ref)
    assert np.abs(w_mine).max() < 2.0 * np.abs(w_ref).max(), (w_mine, w_ref)
    # same conserved energy as the golden (same positions, KE scaled to exactly 300 K)
    assert np.all(np.abs(e_mine - e_ref) / n < 4e-6), np.abs(e_mine - e_ref).max() / n
    assert np.all(np.abs(mine[:, 0] - ref[:, 0]) < 1.5)            # T, K
    assert np.all(np.abs(mine[:, 2] - ref[:, 2]) / n < 3e-4)       # PE, eV/atom
    assert np.allclose(mine[:, 3:6], ref[:, 3:6], rtol=0, atol=0.3)   # diagonal stress, GPa (row noise 0.1)
    assert np.allclose(mine[:, 3:6].mean(axis=0), ref[:, 3:6].mean(axis=0), rtol=0, atol=0.1)
    assert np.array_equal(mine[:, 9:], ref[:, 9:])                 # box columns


def test_carbon_golden_neighbor_maxima():
    """tests/gpumd/carbon/neighbor1.out: "radial(max=100,actual=58), angular(max=59,actual=43)"."""
    import torch
    from gpumd_b200 import engine
    d = np.load(GOLDEN / "carbon_model.npz")
    pos = np.ascontiguousarray(d["pos"].T)
    n = pos.shape[1]
    txt = (GOLDEN / "carbon_neighbor1.out").read_text()
    assert "actual=58" in txt and "actual=43" in txt
    pot = engine.NEP(GOLDEN / "nep_C_2022_NEP4.txt", n)
    atom = engine.Atom(np.zeros(n, np.int32), pos, np.full(n, 12.011))
    box = engine.Box(np.diag([75.2] * 3).reshape(9), np.array([1, 1, 1], np.int32))
    pot.compute(box, atom.type, atom.position_per_atom, atom.potential_per_atom, atom.force_per_atom,
                atom.virial_per_atom)
    pot.check()
    NNr, _, NNa, _ = pot.export_neighbors(100, 59)
    assert NNr.max() == 58 and NNa.max() == 43
