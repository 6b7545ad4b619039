#!/usr/bin/env python
"""Round-2 GPU diagnostics (run on the B200 box; writes gpurun_out/r02_diag_*.json).

  props   the reference's four fixture structures through the C-ABI with the tensor-core and the SIMT
          hidden layer and with the exact-rsqrt build: errors against the FP32 / FP64 oracle
  carbon  energy-conservation trace of the reference's 64 000-atom carbon case from GIVEN velocities,
          ours (tc / simt / exact-rsqrt build) next to the unmodified reference gpumd on the same box

Each variant runs in its own process (the library reads its switches at creation).
"""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
GOLDEN = ROOT / "tests" / "golden"
OUT = ROOT / "gpurun_out"

VARIANTS = {
    "tc": {},
    "simt": {"B200MD_NEP_MLP": "simt"},
    "radial_v1": {"B200MD_NEP_RADIAL": "v1"},
    "exact_rsqrt": {"B200MD_LIB": str(ROOT / "gpumd_b200" / "libb200md_exact.so")},
    "exact_rsqrt_simt": {"B200MD_LIB": str(ROOT / "gpumd_b200" / "libb200md_exact.so"),
                         "B200MD_NEP_MLP": "simt"},
}


def carbon_system():
    d = np.load(GOLDEN / "carbon_model.npz")
    pos = np.ascontiguousarray(d["pos"].T)
    n = pos.shape[1]
    return dict(type=np.zeros(n, np.int32), pos=pos, mass=np.full(n, 12.011),
                h=np.diag([75.2] * 3).reshape(9).astype(np.float64), pbc=np.array([1, 1, 1], np.int32))


def child_props():
    import test_reference_properties_cpu as P
    from gpumd_b200 import engine
    from oracle import oracle_py
    from test_gpu_parity import GpuNep
    res = {}
    for name in P.PAIRS:
        model, s = P.load(name)
        n = s["type"].shape[0]
        orc = oracle_py.NepOracle(GOLDEN / model)
        r32 = orc.compute(s["type"], s["h"], s["pbc"], s["pos"], precision=32)
        r64 = orc.compute(s["type"], s["h"], s["pbc"], s["pos"], precision=64)
        _, out = GpuNep(engine, model, n).compute(s["type"], s["h"], s["pbc"], s["pos"])
        res[name] = dict(
            n=n, sum_abs_e=float(np.abs(r64["pe"]).sum()), fmax=float(np.abs(r64["force"]).max()),
            dE_r32=float(out["pe"].sum() - r32["pe"].sum()), dE_r64=float(out["pe"].sum() - r64["pe"].sum()),
            dE_r32_r64=float(r32["pe"].sum() - r64["pe"].sum()),
            dF_r32=float(np.abs(out["force"] - r32["force"]).max()),
            dF_r64=float(np.abs(out["force"] - r64["force"]).max()),
            dF_r32_r64=float(np.abs(r32["force"] - r64["force"]).max()))
    print("RESULT " + json.dumps(res))


def child_carbon():
    from gpumd_b200 import engine
    from test_gpu_md import run_nve, total_energy
    s = carbon_system()
    n = s["type"].shape[0]
    _, _, rows = run_nve(engine, s, GOLDEN / "nep_C_2022_NEP4.txt", 100, 1.0, 300.0, seed=42, every=10)
    e = total_energy(rows[:-1], n)
    print("RESULT " + json.dumps(dict(n=n, T=rows[:-1, 0].tolist(), U=rows[:-1, 1].tolist(), E=e.tolist())))


def run_children(what):
    res = {}
    for v, env in VARIANTS.items():
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, __file__, "child", what], env=e, capture_output=True, text=True,
                           timeout=900)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
        res[v] = json.loads(line[-1][7:]) if line else dict(error=(r.stdout + r.stderr)[-2000:])
    return res


def reference_carbon():
    """the unmodified reference binary from the same positions and velocities"""
    from gpumd_b200.structures import init_velocities, write_xyz
    exe = ROOT / "oracle" / "_ref" / "gpumd_ref"
    if not exe.exists():
        return dict(error="oracle/_ref/gpumd_ref missing")
    s = carbon_system()
    vel = init_velocities(s["mass"], 300.0, seed=42)
    d = OUT / "r02_ref_carbon"
    d.mkdir(parents=True, exist_ok=True)
    write_xyz(d / "model.xyz", s, ["C"], vel)
    (d / "potential.txt").write_bytes((GOLDEN / "nep_C_2022_NEP4.txt").read_bytes())
    (d / "run.in").write_text("potential potential.txt\nensemble nve\ntime_step 1\ndump_thermo 10\nrun 100\n")
    r = subprocess.run([str(exe)], cwd=d, capture_output=True, text=True, timeout=900)
    (d / "model.xyz").unlink()
    if r.returncode != 0:
        return dict(error=(r.stdout + r.stderr)[-1500:])
    th = np.array([ln.split()[:12] for ln in open(d / "thermo.out") if not ln.startswith("#")],
                  dtype=np.float64)
    return dict(T=th[:, 0].tolist(), K=th[:, 1].tolist(), U=th[:, 2].tolist(), E=(th[:, 1] + th[:, 2]).tolist())


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        {"props": child_props, "carbon": child_carbon}[sys.argv[2]]()
        sys.exit(0)
    OUT.mkdir(exist_ok=True)
    what = sys.argv[1:] or ["props", "carbon"]
    if "props" in what:
        res = run_children("props")
        (OUT / "r02_diag_props.json").write_text(json.dumps(res, indent=1))
        for v, r in res.items():
            for name, x in r.items():
                print(v, name, x if not isinstance(x, dict) else
                      {k: (f"{val:.3e}" if isinstance(val, float) else val) for k, val in x.items()})
    if "carbon" in what:
        res = run_children("carbon")
        res["reference_gpumd"] = reference_carbon()
        (OUT / "r02_diag_carbon.json").write_text(json.dumps(res, indent=1))
        for v, r in res.items():
            if "E" in r:
                e = np.array(r["E"])
                n = 64000
                print(f"{v:18s} E0/N {e[0] / n:.9f}  max|E-E0|/N {np.abs(e - e[0]).max() / n:.3e}  "
                      f"trace {np.array2string((e - e[0]) / n * 1e6, precision=2)} ueV/atom")
            else:
                print(v, r)
