"""Every potential-energy NEP model the reference ships under potentials/nep (read where it lies, so this
module only runs in the build container; skipped where /root/reference is absent): the model loader accepts
it and the kernel bodies (host build) reproduce the oracle on a small random structure.  Includes the
89-species NEP89, whose radial descriptor takes the per-pair contraction path."""
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN  # noqa: F401
from gpumd_b200.structures import diamond, fcc, nep_type_order
from test_kernel_bodies_cpu import check_nep

REF = Path("/root/reference/potentials/nep")
MODELS = {
    "Si_3body": ("Si_2022_NEP4_3body.txt", "diamond", 5.43),
    "Si_4body": ("Si_2022_NEP4_4body.txt", "diamond", 5.43),
    "Si_5body": ("Si_2022_NEP4_5body.txt", "diamond", 5.43),
    "C_2024": ("C_2024_NEP4.txt", "diamond", 3.57),
    "UNEP_v1": ("Song-2024-UNEP-v1-AgAlAuCrCuMgMoNiPbPdPtTaTiVWZr.txt", "alloy", 3.9),
    "NEP89": ("nep89_20250409/nep89_20250409.txt", "alloy89", 3.9),
}


@pytest.mark.parametrize("name", list(MODELS))
def test_shipped_nep_models_match_oracle(oracle, emu, name):
    fname, kind, a = MODELS[name]
    path = REF / fname
    if not path.exists():
        pytest.skip("reference tree not present")
    order = nep_type_order(path)
    if kind == "diamond":
        s = diamond(3, a=a, rattle=0.05, seed=2, symbol="Si")
        s["symbols"] = order
    else:
        s = fcc(5, a, rattle=0.08, seed=4)
        rng = np.random.default_rng(4)
        if kind == "alloy89":  # a dozen of the 89 species, metals so that a = 3.9 A is not absurd
            pick = [order.index(x) for x in ("Al", "Ti", "Fe", "Co", "Ni", "Cu", "Zr", "Mo", "Pd", "Ag", "Pt", "Au")]
            s["type"] = rng.choice(pick, s["type"].shape[0]).astype(np.int32)
        else:
            s["type"] = rng.integers(0, len(order), s["type"].shape[0]).astype(np.int32)
        s["symbols"] = order
    n = s["type"].shape[0]
    check_nep(oracle, emu.nep(path, n), str(path), s, n)
