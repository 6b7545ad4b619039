"""Langevin / BAOAB / Berendsen-barostat kernels against their restatements
(oracle/oracle_py.py) and, for the cuRAND-driven noise, against the statistics it must have.  The
trajectory-level pins against the unmodified reference gpumd are in test_gpu_md.py
(refgpu_md_pbte_{lan,bao,npt}_thermo.out)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
K_B = 8.617343e-5


@pytest.fixture(scope="module")
def eng():
    import torch
    from gpumd_b200 import build, engine
    build.build_lib()
    assert torch.cuda.is_available()
    return engine


def test_langevin_noise_statistics_and_momentum(eng):
    import torch
    n = 1_000_000
    rng = np.random.default_rng(2)
    mass = rng.uniform(1.0, 200.0, n)
    atom = eng.Atom(np.zeros(n, np.int32), np.zeros((3, n)), mass, np.zeros((3, n)))
    T = 300.0
    ens = eng.Ensemble_LAN(n, T, 1e-9)  # c1 = exp(-0.5e9) = 0: pure noise, v = sqrt(kT/m) xi
    assert ens.c1 == 0.0
    ens._kick(atom)
    v = atom.velocity_per_atom.cpu().numpy().reshape(3, n)
    z = v * np.sqrt(mass / (K_B * T))[None, :]
    # 3e6 standard normals: mean 0 +- 5.8e-4, variance 1 +- 8e-4 (1 sigma); allow 5 sigma
    assert abs(z.mean()) < 3e-3 and abs(z.var() - 1.0) < 4e-3
    assert abs((z ** 4).mean() - 3.0) < 0.03  # kurtosis of a normal
    # momentum correction (gpu_find_momentum / gpu_correct_momentum): total momentum is zero
    p = (v * mass[None, :]).sum(axis=1)
    assert np.abs(p).max() < 1e-7 * np.sqrt(n)
    # the stream is a function of (seed, atom index): same seed, same kick; other seed, other kick
    a2 = eng.Atom(np.zeros(n, np.int32), np.zeros((3, n)), mass, np.zeros((3, n)))
    e2 = eng.Ensemble_LAN(n, T, 1e-9)
    e2._kick(a2)
    assert torch.equal(atom.velocity_per_atom, a2.velocity_per_atom)
    a3 = eng.Atom(np.zeros(n, np.int32), np.zeros((3, n)), mass, np.zeros((3, n)))
    e3 = eng.Ensemble_LAN(n, T, 1e-9, seed=7)
    e3._kick(a3)
    assert not torch.equal(atom.velocity_per_atom, a3.velocity_per_atom)
    # c1 = 1 (infinite coupling time): velocities untouched apart from the momentum shift
    e4 = eng.Ensemble_LAN(n, T, 1e30)
    before = atom.velocity_per_atom.clone()
    e4._kick(atom)
    assert torch.allclose(atom.velocity_per_atom, before, rtol=0, atol=1e-12)


def test_baoab_operators_match_restatement(oracle, eng):
    import torch
    rng = np.random.default_rng(3)
    n = 100_003
    mass = rng.uniform(1, 200, n)
    pos, vel, f = rng.normal(size=(3, n)) * 5, rng.normal(size=(3, n)), rng.normal(size=(3, n))
    label = rng.integers(0, 3, n).astype(np.int32)
    for fixed in (-1, 1):
        atom = eng.Atom(np.zeros(n, np.int32), pos, mass, vel)
        atom.force_per_atom.copy_(torch.as_tensor(f.reshape(-1)))
        ens = eng.Ensemble_BAO(n, 300.0, 100.0)
        if fixed >= 0:
            ens.set_groups(label, fixed_group=fixed)
        p, v = pos, vel
        for which in (1, 0, 0, 1):
            ens._op(which, 0.098, atom)
            p, v = oracle.baoab_operator(which, 0.098, mass, p, v, f, label, fixed)
            assert np.allclose(atom.position_per_atom.cpu().numpy().reshape(3, n), p, rtol=1e-15, atol=1e-15)
            assert np.allclose(atom.velocity_per_atom.cpu().numpy().reshape(3, n), v, rtol=1e-15, atol=1e-15)


@pytest.mark.parametrize("num", [1, 3, 6])
def test_berendsen_barostat_matches_restatement(oracle, eng, num):
    import torch
    from gpumd_b200 import lib as L
    lib = L.load()
    rng = np.random.default_rng(4)
    n = 50_001
    h = np.diag([40.0, 41.0, 42.0]).reshape(9)
    if num == 6:
        h = np.array([40.0, 1.5, 0.5, 0.0, 41.0, 2.0, 0.0, 0.0, 42.0])
    pbc = np.array([1, 1, 0 if num == 3 else 1], np.int32)
    pos = rng.uniform(0, 40, (3, n))
    thermo = np.array([300.0, -1.0, 0.010, 0.012, 0.008, 0.001, -0.002, 0.0005])
    p0 = np.array([0.003, 0.002, 0.001, 0.0004, 0.0002, 0.0001])
    pc = np.array([0.5, 0.4, 0.3, 0.2, 0.1, 0.05])
    d_pos = torch.as_tensor(pos.reshape(-1).copy(), device="cuda")
    d_th = torch.as_tensor(thermo, device="cuda")
    hh = (C.c_double * 9)(*h)
    z3i, z3d = (C.c_int * 3)(0, 0, 0), (C.c_double * 3)(0, 0, 0)
    L.check(lib.b200md_berendsen_pressure(
        n, n, num, p0.ctypes.data_as(C.POINTER(C.c_double)), pc.ctypes.data_as(C.POINTER(C.c_double)), z3i, z3d,
        pbc.ctypes.data_as(C.POINTER(C.c_int)), hh, C.c_void_p(d_th.data_ptr()), C.c_void_p(d_pos.data_ptr()), None))
    h_ref, mu = oracle.berendsen_pressure(h, pbc, thermo, p0, pc, num)
    assert np.allclose(np.array(hh[:]), h_ref, rtol=1e-15, atol=1e-15)
    want = mu.reshape(3, 3) @ pos
    assert np.allclose(d_pos.cpu().numpy().reshape(3, n), want, rtol=1e-14, atol=1e-13)
    if num == 3:
        assert hh[8] == h[8]  # the open direction is left alone


@pytest.mark.parametrize("ensemble", ["nvt_lan", "nvt_bao"])
def test_langevin_ensembles_thermalise_to_the_target(eng, ensemble):
    """From 50 K to the 300 K bath: after 1500 steps with tau_T = 100 steps the kinetic temperature sits
    at the target within the canonical fluctuation sqrt(2/(3N)) T = 2.8 K for 4096 atoms (allow 4 sigma)."""
    from conftest import GOLDEN
    from gpumd_b200.structures import rocksalt_pbte
    from test_gpu_md import run_nve
    s = rocksalt_pbte(8, rattle=0.02, seed=1)
    _, _, rows = run_nve(eng, s, GOLDEN / "nep_PbTe.txt", 1500, 1.0, 50.0, seed=42, every=100, ensemble=ensemble)
    T = rows[8:-1, 0]
    assert abs(T.mean() - 300.0) < 6.0, T


def test_hac_matches_restatement(oracle, eng):
    """compute_hac (hac.cu:32-280): the sampled heat-current series against compute_heat's restatement
    summed on the host, then hac / rtc against the numpy restatement of gpu_find_hac + find_rtc."""
    import torch
    rng = np.random.default_rng(5)
    n, steps, interval, Nc = 20_011, 120, 3, 25
    atom = eng.Atom(np.zeros(n, np.int32), np.zeros((3, n)), np.ones(n), np.zeros((3, n)))
    hac = eng.HAC(steps, interval, Nc)
    want = np.zeros((5, steps // interval))
    for step in range(steps):
        if (step + 1) % interval == 0:
            v, w = rng.normal(size=(3, n)), rng.normal(size=(9, n))
            atom.velocity_per_atom.copy_(torch.as_tensor(v.reshape(-1)))
            atom.virial_per_atom.copy_(torch.as_tensor(w.reshape(-1)))
            want[:, (step + 1) // interval - 1] = oracle.compute_heat(w, v).sum(axis=1)
        hac.process(step, atom)
    got = hac.series()
    assert np.allclose(got, want, rtol=1e-12, atol=1e-9)
    dt, T, V = 0.0982, 300.0, 5.0e4
    h, r = hac.postprocess(dt, T, V)
    h0, r0 = oracle.find_hac(got, Nc, dt * interval, T, V)
    assert np.allclose(h, h0, rtol=1e-12, atol=1e-9) and np.allclose(r, r0, rtol=1e-12, atol=1e-9)
