#!/usr/bin/env python
"""Synthetic NEP4 model files for code paths no shipped fixture reaches (random but fixed parameters;
physics-free, only the arithmetic matters):

  nep_synth_50types.txt          50 species -- more than the shared-memory radial accumulators can hold, the
                                 situation of the reference's NEP89 model (potentials/nep/nep89_20250409, 89
                                 species, 8 MB).  Global cutoffs, so the reference's own NEP_CPU reads it.
  nep_synth_pertype_cutoff.txt   3 species with the `cutoff` line in its 2*Nt+2 form (nep.cu:197-237).

File layout as NEP::NEP reads it (nep.cu:88-420): header lines, ANN parameters per type (w0[nneu][dim], b0[nneu],
w1[nneu]), the common bias, c_radial then c_angular in the order [(n*(K+1)+k)*Nt^2 + t1*Nt + t2], q_scaler[dim].

usage: python tests/golden/make_synthetic_models.py   (writes next to itself)
"""
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ELEMENTS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb "
            "Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn").split()


def write(path, symbols, cutoff_line, n_max, basis, l_max, nneu, seed):
    rng = np.random.default_rng(seed)
    nt = len(symbols)
    nr1, na1 = n_max[0] + 1, n_max[1] + 1
    kr1, ka1 = basis[0] + 1, basis[1] + 1
    num_l = l_max[0] + (1 if l_max[1] else 0) + (1 if l_max[2] else 0)
    dim = nr1 + na1 * num_l
    vals = []
    for _ in range(nt):
        vals += list(rng.normal(0.0, 0.4, nneu * dim))   # w0
        vals += list(rng.normal(0.0, 0.3, nneu))         # b0
        vals += list(rng.normal(0.0, 0.5, nneu))         # w1
    vals.append(rng.normal(0.0, 1.0))                    # b1
    vals += list(rng.normal(0.0, 0.3, nt * nt * nr1 * kr1))
    vals += list(rng.normal(0.0, 0.3, nt * nt * na1 * ka1))
    vals += list(rng.uniform(0.05, 0.6, dim))            # q_scaler
    with open(path, "w") as f:
        f.write("nep4 %d %s\n" % (nt, " ".join(symbols)))
        f.write(cutoff_line + "\n")
        f.write("n_max %d %d\nbasis_size %d %d\nl_max %d %d %d\nANN %d 0\n" % (*n_max, *basis, *l_max, nneu))
        for v in vals:
            f.write("%.7e\n" % v)
    print(path.name, nt, "types, dim", dim, len(vals), "parameters,", path.stat().st_size // 1024, "KB")


write(HERE / "nep_synth_50types.txt", ELEMENTS[:50], "cutoff 5 4 80 40", (1, 1), (3, 3), (4, 2, 1), 4, seed=11)
write(HERE / "nep_synth_pertype_cutoff.txt", ["Cu", "O", "H"], "cutoff 5.0 4.0 4.4 3.6 3.8 3.2 80 40",
      (3, 2), (6, 5), (4, 2, 0), 10, seed=12)
