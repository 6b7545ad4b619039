#!/usr/bin/env python
"""Workload for the ncu capture of the kernels the headline bench does not exercise with real work:
the EAM density / force passes (1 M-atom fcc CuFeNi, eam_zhou_2004) and the neighbour REBUILD kernels
(cell sort, skin list, type tiles) of the 1 M-atom PbTe NEP case, forced with invalidate().
Prints CUDA-event times so the same script doubles as a stand-alone measurement.

usage: python scripts/prof_eam_rebuild.py [cells]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from gpumd_b200 import build, engine  # noqa: E402
from gpumd_b200.structures import fcc, rocksalt_pbte  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
build.build_lib()
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 63


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# --- EAM
s = fcc(cells, 3.61, rattle=0.02, seed=3, num_types=3, symbols=["Cu", "Fe", "Ni"])
n = s["type"].shape[0]
pot = engine.EAM(GOLDEN / "eam_zhou_2004_CuFeNi.txt", n)
atom = engine.Atom(s["type"], s["pos"], s["mass"])
box = engine.Box(s["h"], s["pbc"])
args = (box, atom.type, atom.position_per_atom, atom.potential_per_atom, atom.force_per_atom, atom.virial_per_atom)
for _ in range(3):
    pot.compute(*args)
pot.check()
ms = timed(lambda: pot.compute(*args), 20)
print("EAM zhou_2004 CuFeNi: %d atoms, %.3f ms per force call, %.3g atom-calls/s" % (n, ms, n / ms * 1e3))

# --- neighbour rebuild of the headline case
s = rocksalt_pbte(50, rattle=0.02, seed=1)
n = s["type"].shape[0]
nep = engine.NEP(GOLDEN / "nep_PbTe.txt", n)
atom = engine.Atom(s["type"], s["pos"], s["mass"])
box = engine.Box(s["h"], s["pbc"])
args = (box, atom.type, atom.position_per_atom, atom.potential_per_atom, atom.force_per_atom, atom.virial_per_atom)
for _ in range(3):
    nep.compute(*args)
nep.check()
keep = timed(lambda: nep.compute(*args), 10)


def rebuilt():
    nep.invalidate(n)
    nep.compute(*args)


reb = timed(rebuilt, 5)
print("PbTe NEP: %d atoms, force call %.3f ms with lists kept, %.3f ms with a full rebuild (+%.3f ms)" % (n, keep, reb, reb - keep))
