"""N>1 path on CPU: world_size-2 gloo processes run the slab decomposition (gpumd_b200/domain.py,
backend="torch") with the oracle as the force function and must reproduce the single-domain result
for the atoms they own -- before and after atoms migrate across the slab boundary."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpumd_b200.domain import SlabDomain
        from gpumd_b200.structures import rocksalt_pbte
        from oracle import oracle_py
        s = rocksalt_pbte((6, 4, 4), rattle=0.05, seed=3)  # 768 atoms, 39.4 x 26.3 x 26.3 A
        orc = oracle_py.NepOracle(GOLDEN / "nep_PbTe.txt")
        dom = SlabDomain(s["h"], s["pbc"], orc.rc_radial, rank, world, "cpu", backend="torch")
        dom.distribute(s["type"], s["pos"], s["mass"])
        results = {}
        rng = np.random.default_rng(100)  # same stream on both ranks
        pos_global = s["pos"].copy()
        for phase in range(3):
            n, m = dom.n_own, dom.n_loc
            loc = dom.pos.view(3, m).numpy()
            r = orc.compute(dom.type.numpy(), dom.local_h, dom.local_pbc, loc, precision=32, lists=True)
            ids = dom.own_id.numpy()
            results[phase] = dict(ids=ids.copy(), force=r["force"][:, :n].copy(), pe=r["pe"][:n].copy(),
                                  virial=r["virial"][:, :n].copy(), NN=r["NN_radial"][:n].copy(),
                                  n_loc=m, pos=pos_global.copy())
            # move every atom (same global displacement on both ranks); phase 1 pushes a layer of
            # atoms across the slab boundaries so that they must migrate
            disp = rng.normal(0, 0.05, pos_global.shape)
            if phase == 1:
                disp[0] += 0.4
            pos_global = pos_global + disp
            pos_global[1] = np.mod(pos_global[1], s["h"][4])
            pos_global[2] = np.mod(pos_global[2], s["h"][8])
            dom.pos.view(3, m)[:, :n] += torch.as_tensor(disp[:, ids])
            dom.sync_owned_views()
            dom.exchange()
        # every atom is owned by exactly one rank after the migrations
        counts = torch.zeros(s["type"].shape[0], dtype=torch.int64)
        counts[dom.own_id] += 1
        dist.all_reduce(counts)
        results["owned_once"] = bool((counts == 1).all().item())
        torch.save(results, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_slab_decomposition_matches_single_domain(oracle, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from gpumd_b200.structures import rocksalt_pbte
    s = rocksalt_pbte((6, 4, 4), rattle=0.05, seed=3)
    orc = oracle.NepOracle(GOLDEN / "nep_PbTe.txt")
    res = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(world)]
    assert all(r["owned_once"] for r in res)
    for phase in range(3):
        pos = res[0][phase]["pos"]
        pos_w = pos.copy()
        pos_w[0] = np.mod(pos_w[0], s["h"][0])
        ref = orc.compute(s["type"], s["h"], s["pbc"], pos_w, precision=32, lists=True)
        seen = np.zeros(s["type"].shape[0], bool)
        for r in res:
            ids = r[phase]["ids"]
            assert not seen[ids].any()
            seen[ids] = True
            # neighbour counts are exact; forces/energies agree up to FP32 summation order
            assert np.array_equal(r[phase]["NN"], ref["NN_radial"][ids])
            assert np.allclose(r[phase]["force"], ref["force"][:, ids], rtol=1e-4, atol=1e-5)
            assert np.allclose(r[phase]["pe"], ref["pe"][ids], rtol=1e-5, atol=2e-6)
            assert np.allclose(r[phase]["virial"], ref["virial"][:, ids], rtol=1e-4, atol=2e-5)
            assert r[phase]["n_loc"] > len(ids)  # there are ghosts
        assert seen.all()
    # phase 2 follows the +0.4 A drift: ownership must have changed for some atoms
    assert not np.array_equal(np.sort(res[0][0]["ids"]), np.sort(res[0][2]["ids"]))
