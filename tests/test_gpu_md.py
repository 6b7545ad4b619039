"""MD-level and full-size checks on the GPU: NVE conservation (the reference's own bound,
tests_pytest/test_md_conservation.py:27,57-64), parity after many steps, and the size-independent
properties at BASELINE.json's 1M-atom size."""
import numpy as np
import pytest

from conftest import GOLDEN, TOL
from gpumd_b200.structures import TIME_UNIT_CONVERSION, fcc, init_velocities, rocksalt_pbte
from test_kernel_bodies_cpu import check_fv

pytestmark = pytest.mark.gpu


def run_nve(eng, s, pot_file, steps, dt_fs, temperature, seed=42, every=None, ensemble="nve", T_target=300.0):
    import torch
    n = s["type"].shape[0]
    vel = init_velocities(s["mass"], temperature, seed)
    atom = eng.Atom(s["type"], s["pos"], s["mass"], vel)
    box = eng.Box(s["h"], s["pbc"])
    force = eng.Force()
    pot = force.parse_potential(pot_file, n)
    if ensemble == "nvt_ber":
        ens = eng.Ensemble_BER(n, T_target, 100.0)
    elif ensemble == "nvt_nhc":
        ens = eng.Ensemble_NHC(n, T_target, 100.0, dt_fs / TIME_UNIT_CONVERSION)
    elif ensemble == "nvt_nhc_dt1":  # chain masses from a 1 fs step whatever dt is (a former driver bug)
        ens = eng.Ensemble_NHC(n, T_target, 100.0, 1.0 / TIME_UNIT_CONVERSION)
    elif ensemble == "nvt_lan":
        ens = eng.Ensemble_LAN(n, T_target, 100.0)
    elif ensemble == "nvt_bao":
        ens = eng.Ensemble_BAO(n, T_target, 100.0)
    elif ensemble == "npt_ber":  # `ensemble npt_ber 300 300 100 0 50 1000` (isotropic, 50 GPa modulus)
        ens = eng.Ensemble_NPT_BER(n, T_target, 100.0, [0.0], [50.0], 1000.0)
    elif ensemble == "nvt_bdp":
        ens = eng.Ensemble_BDP(n, T_target, 100.0)
    else:
        ens = eng.Ensemble_NVE(n)
    thermo = torch.zeros(8, dtype=torch.float64, device="cuda")
    dt = dt_fs / TIME_UNIT_CONVERSION
    args = (box, atom.position_per_atom, atom.type, atom.potential_per_atom, atom.force_per_atom,
            atom.virial_per_atom)
    force.compute(*args)
    ens.find_thermo(box.get_volume(), atom, thermo)
    rows = [thermo.cpu().numpy().copy()]
    for step in range(steps):
        ens.compute1(dt, box, atom, thermo)
        force.compute(*args)
        ens.compute2(dt, box, atom, thermo)
        if every and (step + 1) % every == 0:
            rows.append(thermo.cpu().numpy().copy())
    pot.check()
    rows.append(thermo.cpu().numpy().copy())
    return atom, pot, np.array(rows)


def total_energy(rows, n):
    return 1.5 * n * 8.617343e-5 * rows[:, 0] + rows[:, 1]


def test_nve_conservation_pbte(oracle, eng_mod):
    s = rocksalt_pbte(6, rattle=0.02, seed=1)  # 1728 atoms
    n = s["type"].shape[0]
    atom, pot, rows = run_nve(eng_mod, s, GOLDEN / "nep_PbTe.txt", 200, 1.0, 300.0, every=10)
    e = total_energy(rows, n)
    dt = 1.0
    assert np.abs(e - e[0]).max() < 2e-3 * dt * dt * n  # reference bound: 2e-3*dt^2*N eV
    assert np.abs(e - e[0]).max() < 5e-5 * n  # and far tighter in practice
    assert pot.num_rebuilds >= 1
    # after 200 steps the device state still matches an oracle evaluation at the same positions
    pos = atom.position_per_atom.cpu().numpy().reshape(3, n)
    r = oracle.NepOracle(GOLDEN / "nep_PbTe.txt").compute(s["type"], s["h"], s["pbc"], pos)
    check_fv(dict(force=atom.force_per_atom.cpu().numpy().reshape(3, n),
                  virial=atom.virial_per_atom.cpu().numpy().reshape(9, n)), r)
    # total momentum stays zero
    v = atom.velocity_per_atom.cpu().numpy().reshape(3, n)
    assert np.abs((v * s["mass"]).sum(axis=1)).max() < 1e-6 * np.sqrt(n)


def test_nve_conservation_lj(oracle, eng_mod):
    # The reference's LJ is truncated, NOT shifted (lj.cu:67-75,128): a pair crossing the cutoff
    # changes the energy by U(rc) = -6.4e-5 eV.  At a = 5.30 A the k=7 fcc shell sits 0.085 A inside
    # rc = 10 A, so energy is not conserved there by construction; a = 5.60 A keeps the nearest
    # shells 0.30 / 0.48 A away from the cutoff.
    s = fcc(8, 5.60, rattle=0.0, seed=1)  # 2048 atoms, 44.8 A box
    n = s["type"].shape[0]
    atom, pot, rows = run_nve(eng_mod, s, GOLDEN / "lj_Ar_10A.txt", 300, 5.0, 40.0, every=10)
    e = total_energy(rows, n)
    assert np.abs(e - e[0]).max() < 1e-4 * n  # O(dt^2) fluctuation at dt = 5 fs, no drift
    assert abs(e[-1] - e[0]) < 3e-5 * n
    assert pot.num_rebuilds >= 1
    pos = atom.position_per_atom.cpu().numpy().reshape(3, n)
    r = oracle.lj_compute(np.array([[[1.032e-2, 3.405, 10.0]]]), s["type"], s["h"], s["pbc"], pos)
    check_fv(dict(force=atom.force_per_atom.cpu().numpy().reshape(3, n),
                  virial=atom.virial_per_atom.cpu().numpy().reshape(9, n)), r)


def read_thermo(path):
    rows = [ln.split() for ln in open(path) if not ln.startswith("#")]
    return np.array(rows, dtype=np.float64)


@pytest.mark.parametrize("case", ["md_pbte", "md_lj", "md_si", "md_pbte_nhc", "md_pbte_ber", "md_pbte_bdp",
                                  "md_pbte_lan", "md_pbte_bao", "md_pbte_npt"])
def test_nve_trajectory_matches_reference_gpu(eng_mod, case):
    """tests/golden/refgpu_md_*_thermo.out: thermo.out (every 10 steps, 200 steps) written by the
    unmodified reference gpumd on a B200 from the same positions and velocities
    (scripts/run_reference_gpumd.py).  T, kinetic, potential energy and the diagonal stress of our
    trajectory must track it; the two FP32 force fields differ in summation order, so the
    trajectories separate slowly (Lyapunov) -- tolerances widen with time accordingly."""
    from gpumd_b200.structures import diamond
    if not (GOLDEN / f"refgpu_{case}_thermo.out").exists():
        pytest.skip("reference-GPU fixture not generated yet")
    ensemble = "nve"
    if case.startswith("md_pbte"):
        s, pot_file, dt, T0 = rocksalt_pbte(20, rattle=0.02, seed=1), GOLDEN / "nep_PbTe.txt", 1.0, 300.0
        if case == "md_pbte_npt":
            ensemble = "npt_ber"
        elif case != "md_pbte":
            ensemble = "nvt_" + case.split("_")[-1]
    elif case == "md_si":
        s, pot_file, dt, T0 = diamond(20, a=5.431, rattle=0.0, seed=1), GOLDEN / "tersoff_Si_1989.txt", 1.0, 300.0
    else:
        s, pot_file, dt, T0 = fcc(25, 5.30, rattle=0.0, seed=1), GOLDEN / "lj_Ar_10A.txt", 5.0, 80.0
    n = s["type"].shape[0]
    ref = read_thermo(GOLDEN / f"refgpu_{case}_thermo.out")  # T KE PE sxx syy szz syz sxz sxy ...
    atom, pot, rows = run_nve(eng_mod, s, pot_file, 200, dt, T0, seed=42, every=10, ensemble=ensemble)
    mine = rows[1:-1]  # rows[0] is step 0, the last row repeats step 200
    assert mine.shape[0] == ref.shape[0] == 20
    T, U = mine[:, 0], mine[:, 1]
    PRESSURE_UNIT_CONVERSION = 1.602177e+2  # eV/A^3 -> GPa, common.cuh:25
    for k in range(20):
        tol = 2e-5 * (1 + k)  # relative; grows with time
        assert abs(T[k] - ref[k, 0]) < tol * T0 * 10, (k, T[k], ref[k, 0])
        assert abs(U[k] - ref[k, 2]) < tol * abs(ref[k, 2]) , (k, U[k], ref[k, 2])
        for c in range(3):
            assert abs(mine[k, 2 + c] * PRESSURE_UNIT_CONVERSION - ref[k, 3 + c]) < 2e-3 + 1e-3 * abs(ref[k, 3 + c])
    # first output (step 10) is still essentially round-off limited
    assert abs(T[0] - ref[0, 0]) < 2e-3 and abs(U[0] - ref[0, 2]) / n < 2e-7


@pytest.fixture(scope="module")
def eng_mod():
    import torch
    from gpumd_b200 import build, engine
    build.build_lib()
    assert torch.cuda.is_available()
    return engine


def test_full_size_properties_pbte_1m(oracle, eng_mod):
    """BASELINE config C3 size (1 000 000 atoms): properties that do not need an O(N) oracle run --
    Newton's third law (sum of forces = 0), symmetric neighbour relation, virial symmetry of the
    pair part, agreement of a spatial sub-block with an oracle evaluation of that block's
    neighbourhood, determinism."""
    import torch
    eng = eng_mod
    s = rocksalt_pbte(50, rattle=0.02, seed=1)
    n = s["type"].shape[0]
    assert n == 1_000_000
    pot = eng.NEP(GOLDEN / "nep_PbTe.txt", n)
    atom = eng.Atom(s["type"], s["pos"], s["mass"])
    box = eng.Box(s["h"], s["pbc"])
    args = (box, atom.type, atom.position_per_atom, atom.potential_per_atom, atom.force_per_atom,
            atom.virial_per_atom)
    pot.compute(*args)
    pot.check()
    f = atom.force_per_atom.cpu().numpy().reshape(3, n)
    pe = atom.potential_per_atom.cpu().numpy()
    assert np.isfinite(f).all() and np.isfinite(pe).all()
    # FP32 per-atom sums: residual ~ sqrt(N) * 1e-6 * |f|
    assert np.abs(f.sum(axis=1)).max() < 2e-3
    NNr, NLr, NNa, NLa = pot.export_neighbors()
    assert NNr.min() >= 40 and NNr.max() <= 80 and NNa.max() <= 10
    # symmetric relation: j in list(i) <=> i in list(j), checked on a sample
    rng = np.random.default_rng(0)
    for i in rng.integers(0, n, 200):
        for j in NLr[i, :NNr[i]][::7]:
            assert i in NLr[j, :NNr[j]]
    # a sub-block: oracle on all atoms within rc_r + rc_r of a probe region, compare inner atoms
    pos = s["pos"]
    centre = np.array([100.0, 120.0, 140.0])
    d = pos - centre[:, None]
    inner = np.nonzero((np.abs(d) < 6.0).all(axis=0))[0]
    shell = np.nonzero((np.abs(d) < 6.0 + 16.5).all(axis=0))[0]
    sub = dict(pos=np.ascontiguousarray(pos[:, shell]), type=s["type"][shell],
               h=s["h"], pbc=np.array([0, 0, 0], np.int32))
    r = oracle.NepOracle(GOLDEN / "nep_PbTe.txt").compute(sub["type"], sub["h"], sub["pbc"], sub["pos"])
    where = np.searchsorted(shell, inner)
    assert np.allclose(f[:, inner], r["force"][:, where], rtol=1e-4, atol=1e-5)
    assert np.allclose(pe[inner], r["pe"][where], rtol=1e-5, atol=1e-6)
    # determinism: a second evaluation is bit-identical
    atom.force_per_atom.zero_(); atom.potential_per_atom.zero_(); atom.virial_per_atom.zero_()
    pot.compute(*args)
    assert np.array_equal(f, atom.force_per_atom.cpu().numpy().reshape(3, n))
