"""Multi-GPU (needs >= 2 devices: run with `gpurun --gpus 2`): the slab-decomposed NVE run must
track the single-GPU run of the same system -- same thermo trajectory up to FP32 summation order."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, steps, ensemble):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from gpumd_b200.domain import DomainMD, SlabDomain
        from gpumd_b200.structures import TIME_UNIT_CONVERSION, init_velocities, rocksalt_pbte
        s = rocksalt_pbte((12, 6, 6), rattle=0.02, seed=1)  # 3456 atoms, 78.8 x 39.4 x 39.4 A
        vel = init_velocities(s["mass"], 600.0, seed=42)
        dom = SlabDomain(s["h"], s["pbc"], 8.0, rank, world, "cuda")
        dom.distribute(s["type"], s["pos"], s["mass"], vel)
        dt = 2.0 / TIME_UNIT_CONVERSION
        md = DomainMD(dom, GOLDEN / "nep_PbTe.txt", ensemble=ensemble, temperature=450.0,
                      temperature_coupling=50.0, time_step=dt)
        md.compute_force()
        md.find_thermo()
        rows = [md.read_thermo()]
        for k in range(steps):
            if k == 30:
                md.exchange()  # a forced migration/re-order in the middle of the run
            else:
                md.maybe_exchange(5)
            md.step(dt)
            if (k + 1) % 10 == 0:
                rows.append(md.read_thermo())
        md.pot.check()
        if rank == 0:
            np.save(os.path.join(out_dir, "multi.npy"), np.array(rows))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ensemble", ["nve", "nvt_bdp", "nvt_nhc"])
def test_two_domains_track_single_gpu(tmp_path, ensemble):
    """NVE, and two thermostats whose state (BDP generator, NHC chain) is replicated on every rank and
    advanced from the all-reduced temperature."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from gpumd_b200 import build, engine
    from gpumd_b200.structures import TIME_UNIT_CONVERSION, init_velocities, rocksalt_pbte
    build.build_lib()
    steps = 100
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), steps, ensemble), nprocs=2, join=True)
    multi = np.load(tmp_path / "multi.npy")
    # single GPU, full periodic box
    s = rocksalt_pbte((12, 6, 6), rattle=0.02, seed=1)
    n = s["type"].shape[0]
    atom = engine.Atom(s["type"], s["pos"], s["mass"], init_velocities(s["mass"], 600.0, seed=42))
    box = engine.Box(s["h"], s["pbc"])
    force = engine.Force()
    pot = force.parse_potential(GOLDEN / "nep_PbTe.txt", n)
    dt = 2.0 / TIME_UNIT_CONVERSION
    if ensemble == "nvt_bdp":
        ens = engine.Ensemble_BDP(n, 450.0, 50.0)
    elif ensemble == "nvt_nhc":
        ens = engine.Ensemble_NHC(n, 450.0, 50.0, dt)
    else:
        ens = engine.Ensemble_NVE(n)
    thermo = torch.zeros(8, dtype=torch.float64, device="cuda")
    args = (box, atom.position_per_atom, atom.type, atom.potential_per_atom, atom.force_per_atom,
            atom.virial_per_atom)
    force.compute(*args)
    ens.find_thermo(box.get_volume(), atom, thermo)
    single = [thermo.cpu().numpy().copy()]
    for k in range(steps):
        ens.compute1(dt, box, atom, thermo)
        force.compute(*args)
        ens.compute2(dt, box, atom, thermo)
        if (k + 1) % 10 == 0:
            single.append(thermo.cpu().numpy().copy())
    pot.check()
    single = np.array(single)
    assert multi.shape == single.shape
    # step 0: identical system, only the summation order differs
    assert abs(multi[0, 0] - single[0, 0]) < 1e-9 * 600
    assert abs(multi[0, 1] - single[0, 1]) / n < 1e-7
    assert np.allclose(multi[0, 2:], single[0, 2:], rtol=1e-5, atol=1e-7)
    for k in range(1, multi.shape[0]):
        tol = 3e-5 * (1 + k)
        assert abs(multi[k, 0] - single[k, 0]) < tol * 600, (k, multi[k, 0], single[k, 0])
        assert abs(multi[k, 1] - single[k, 1]) < tol * abs(single[k, 1])
