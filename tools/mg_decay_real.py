"""Tuning aid (torchrun, 2+ GPUs): does the slab step slow down over a run?  Rank 0's stage times per window."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import __graft_entry__ as g
from bench import make_workload
import importlib
pkg = g.load_package()
mg = importlib.import_module(pkg.__name__ + ".multigpu")
lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rank, world = dist.get_rank(), dist.get_world_size()
w = make_workload("c4", 126 ** 3)
every = int(os.environ.get("NB200_MIGRATE_EVERY", "5"))
sim = mg.SlabSimulation(pkg, w, rank, world, lr, dist, migrate_every=every)
for win in range(6):
    sim.step_async(40); sim.sync()
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    sim.step_async(20); sim.sync()
    ms = (time.perf_counter() - t0) * 1e3 / 20
    sim.h.set_profiling(True)
    sim.step_async(10); sim.sync()
    st = sim.h.get_stage_times(); sim.h.set_profiling(False)
    extra = int(os.environ.get("NB200_EXTRA", "0"))
    if extra:
        sim.step_async(extra); sim.sync()
    stats = sim.h.get_stats()
    if rank == 0:
        print("decay every", every, "steps", (win + 1) * 70, "ms/step %.3f" % ms, "owned", sim.h.mg_owned_count(), "ghosts", sim.n_ghost,
              "slots", stats["n_atoms"], "tiles", stats["n_slots"] // 64, {a: round(b[0] / 10, 3) for a, b in st.items() if b[1] > 0}, flush=True)
sim.close(); dist.barrier(); dist.destroy_process_group()
