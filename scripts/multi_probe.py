"""Where does the multi-GPU step time go?  torchrun --nproc-per-node N scripts/multi_probe.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gpumd_b200.domain import DomainMD, SlabDomain
from gpumd_b200.structures import TIME_UNIT_CONVERSION, init_velocities, rocksalt_pbte

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
cells = int(os.environ.get("CELLS", "50"))
s = rocksalt_pbte((cells * world, cells, cells), rattle=0.02, seed=1)
vel = init_velocities(s["mass"], 300.0, seed=42)
dom = SlabDomain(s["h"], s["pbc"], 8.0, rank, world, "cuda", skin=3.0)
dom.distribute(s["type"], s["pos"], s["mass"], vel)
md = DomainMD(dom, os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "nep_PbTe.txt"))
dt = 1.0 / TIME_UNIT_CONVERSION
md.compute_force()


def run(label, steps, fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(steps):
        fn()
    e1.record(); t_issue = time.time() - t0
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label:34s} {ms.item():8.3f} ms/step   host issue {t_issue / steps * 1e3:6.3f} ms/step", flush=True)


k = [0]
def step_plain():
    md.step(dt)
def step_nothermo():
    md.step(dt, reduce_thermo=False)
def step_check():
    k[0] += 1
    if k[0] % 5 == 0:
        dom.needs_exchange()
    md.step(dt)
def force_only():
    md.compute_force()
def halo_only():
    dom.halo_update()

run("force only (no comm)", 40, force_only)
run("halo only", 40, halo_only)
run("step, no thermo all-reduce", 40, step_nothermo)
run("step (thermo all-reduce)", 40, step_plain)
run("step + displacement check / 5", 40, step_check)
dist.barrier()
