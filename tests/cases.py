"""Seeded input structures shared by the CPU and GPU parity tests."""
import numpy as np

from conftest import GOLDEN
from gpumd_b200.structures import diamond, fcc, nep_type_order, read_xyz, rocksalt_pbte


def shear(s, xy=0.15, yz=-0.1, xz=0.05):
    """Same fractional coordinates in a triclinic box."""
    Hm = s["h"].reshape(3, 3).copy()
    Hn = Hm.copy()
    Hn[0, 1] = xy * Hm[0, 0]
    Hn[1, 2] = yz * Hm[1, 1]
    Hn[0, 2] = xz * Hm[0, 0]
    frac = np.linalg.solve(Hm, s["pos"])
    s = dict(s)
    s["pos"] = np.ascontiguousarray(Hn @ frac)
    s["h"] = Hn.reshape(9)
    return s


def bazro3_supercell(reps=3, seed=5):
    order = nep_type_order(GOLDEN / "nep_BaZrO3.txt")
    b = read_xyz(GOLDEN / "BaZrO3-nat40-rattled.xyz", order)
    L = b["h"][0]
    shifts = np.array([[i, j, k] for i in range(reps) for j in range(reps) for k in range(reps)], float) * L
    pos = (b["pos"].T[None, :, :] + shifts[:, None, :]).reshape(-1, 3)
    pos += np.random.default_rng(seed).normal(0, 0.02, pos.shape)
    pos = np.mod(pos, L * reps)
    return dict(pos=np.ascontiguousarray(pos.T), type=np.tile(b["type"], reps ** 3).astype(np.int32),
                h=np.diag([L * reps] * 3).reshape(9).astype(np.float64), pbc=np.array([1, 1, 1], np.int32))


def slab(s):
    """Open boundary along z: widen the box by 12 A of vacuum and switch pbc_z off."""
    s = dict(s)
    h = s["h"].copy()
    h[8] += 12.0
    s["h"] = h
    s["pbc"] = np.array([1, 1, 0], np.int32)
    return s


# name -> (model file, structure factory)
NEP_CASES = {
    # the headline model: 2 types, D=30, rc 8/4, 3-body only
    "PbTe": ("nep_PbTe.txt", lambda: rocksalt_pbte(4, rattle=0.05, seed=1)),
    "PbTe_triclinic": ("nep_PbTe.txt", lambda: shear(rocksalt_pbte(4, rattle=0.05, seed=2))),
    "PbTe_slab": ("nep_PbTe.txt", lambda: slab(rocksalt_pbte(4, rattle=0.05, seed=4))),
    # 1 type, basis 10/8 (padded to 13/9), n_max 10/8, 4- and 5-body terms, D=65
    "carbon": ("nep_C_2022_NEP4.txt", lambda: diamond(5, a=3.57, rattle=0.05, seed=3, symbol="C")),
    # 3 types (shared-memory accumulator path), n_max 8/6, D=44, ~170 radial neighbours
    "BaZrO3": ("nep_BaZrO3.txt", lambda: bazro3_supercell()),
    # 16 types, universal ZBL, 4- and 5-body terms (config C4's model)
    "UNEP": ("nep_UNEP_v1.txt", lambda: fcc(
        6, 3.9, rattle=0.08, seed=7, num_types=16,
        symbols=nep_type_order(GOLDEN / "nep_UNEP_v1.txt"))),
}


def _random_alloy(model, cells, a, seed):
    """fcc lattice with the species of `model` drawn i.i.d. (masses are irrelevant for single points)."""
    s = fcc(cells, a, rattle=0.08, seed=seed)
    order = nep_type_order(GOLDEN / model)
    s["type"] = np.random.default_rng(seed).integers(0, len(order), s["type"].shape[0]).astype(np.int32)
    s["symbols"] = order
    return s


# synthetic models (tests/golden/make_synthetic_models.py) for paths no shipped model reaches
NEP_CASES_SYNTH = {
    # 50 species: beyond the shared-memory radial accumulators -> per-pair contraction (the NEP89 situation)
    "synth50": ("nep_synth_50types.txt", lambda: _random_alloy("nep_synth_50types.txt", 5, 4.2, 9)),
    # `cutoff` line in its 2*Nt+2 form: per-type radial / angular cutoffs, pair cutoff = mean of the two
    "pertype_cutoff": ("nep_synth_pertype_cutoff.txt",
                       lambda: _random_alloy("nep_synth_pertype_cutoff.txt", 5, 4.0, 10)),
}


def _golden_static():
    order = nep_type_order(GOLDEN / "nep_PbTe_static.txt")
    return read_xyz(GOLDEN / "gpumd_static_model.xyz", order)


def _golden_bazro3():
    order = nep_type_order(GOLDEN / "nep_BaZrO3.txt")
    return read_xyz(GOLDEN / "BaZrO3-nat40-rattled.xyz", order)


# small periodic boxes (the reference's explicit-image path, nep_small_box.cuh): evaluated through a
# supercell by the library; the oracle sums explicit images like the reference
NEP_SMALL_CASES = {
    # config C1: 216-atom PbTe, 19.7 A box (< 2.5 * 9 A)
    "PbTe_C1": ("nep_PbTe.txt", lambda: rocksalt_pbte(3, rattle=0.05, seed=1)),
    # examples/gpumd_static: 250 atoms, triclinic
    "PbTe_static_golden": ("nep_PbTe_static.txt", _golden_static),
    # tests_pytest bulk_bazro3 fixture: 40 atoms, 8.4 A box -> 3 replicas per direction
    "BaZrO3_40": ("nep_BaZrO3.txt", _golden_bazro3),
    # thin slab: periodic in x,y (one of them short), open in z
    "PbTe_thin": ("nep_PbTe.txt", lambda: slab(rocksalt_pbte((5, 2, 2), rattle=0.05, seed=6))),
}


def check_reference_goldens(dev_factory, assert_close, TOL):
    """The reference's own known answers for the path, applied DIRECTLY to an implementation
    (dev_factory(model_file, n) -> object with .compute(type, h, pbc, pos) -> (rc, dict)):
    examples/gpumd_static/dump.xyz (250-atom PbTe: energy, 9-virial, per-atom forces) and
    tests_pytest/fixtures/golden/bulk_bazro3.npz (40-atom BaZrO3: energy, forces, stress).  Both are
    small-box inputs."""
    order = nep_type_order(GOLDEN / "nep_PbTe_static.txt")
    s = read_xyz(GOLDEN / "gpumd_static_model.xyz", order)
    gold = read_xyz(GOLDEN / "gpumd_static_dump.xyz", order)
    rc, r = dev_factory("nep_PbTe_static.txt", s["type"].shape[0]).compute(s["type"], s["h"], s["pbc"], s["pos"])
    assert rc == 0
    assert_close(r["pe"].sum(), gold["energy"], **TOL["energy"], what="energy")
    assert_close(r["force"], gold["forces"], **TOL["force"], what="force")
    v = r["virial"].sum(axis=1)  # GPUMD order xx yy zz xy xz yz yx zx zy -> row-major 3x3
    v33 = np.array([v[0], v[3], v[4], v[6], v[1], v[5], v[7], v[8], v[2]])
    assert_close(v33, gold["virial"], rtol=1e-4, atol=2e-4, what="virial")

    order = nep_type_order(GOLDEN / "nep_BaZrO3.txt")
    s = read_xyz(GOLDEN / "BaZrO3-nat40-rattled.xyz", order)
    g = np.load(GOLDEN / "bulk_bazro3.npz")
    rc, r = dev_factory("nep_BaZrO3.txt", s["type"].shape[0]).compute(s["type"], s["h"], s["pbc"], s["pos"])
    assert rc == 0
    assert_close(r["pe"].sum(), g["energy"], **TOL["energy"], what="energy")
    assert_close(r["force"].T, g["forces"], rtol=1e-4, atol=2e-5, what="forces")
    vol = abs(np.linalg.det(s["h"].reshape(3, 3)))
    v = r["virial"].sum(axis=1)
    stress = -np.array([v[0], v[1], v[2], v[5], v[4], v[3]]) / vol  # Voigt xx yy zz yz xz xy
    assert_close(stress, g["stress"], rtol=1e-4, atol=1e-6, what="stress")
