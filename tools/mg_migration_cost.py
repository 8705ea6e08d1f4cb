"""Tuning aid (torchrun, 2+ GPUs): wall time of single slab steps with a sync after each, migration steps marked."""
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
m = 126
w = make_workload("c4", m ** 3)
every = int(os.environ.get("NB200_MIGRATE_EVERY", "5"))
sim = mg.SlabSimulation(pkg, w, rank, world, lr, dist, migrate_every=every)
sim.step_async(50); sim.sync()
ts = []
for s in range(40):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.step_async(1); sim.sync()
    ts.append((time.perf_counter() - t0) * 1e3)
# chunks: 4 pipelined normal steps, then the migration step alone; then 10 pipelined steps spanning two migrations
ca, cb, cc = [], [], []
for rep in range(6):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    sim.step_async(4); sim.sync(); ca.append((time.perf_counter() - t0) * 1e3 / 4)
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    sim.step_async(1); sim.sync(); cb.append((time.perf_counter() - t0) * 1e3)
for rep in range(4):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    sim.step_async(10); sim.sync(); cc.append((time.perf_counter() - t0) * 1e3 / 10)
if rank == 0:
    print("every: 4 pipelined normal steps ms/step", " ".join("%.3f" % t for t in ca))
    print("every: migration step alone ms", " ".join("%.3f" % t for t in cb))
    print("every: 10 pipelined steps (2 migrations) ms/step", " ".join("%.3f" % t for t in cc))
dist.barrier(); torch.cuda.synchronize()
sim.h.mg_set_migration(None, 0)
t0 = time.perf_counter()
sim.step_async(12)              # returns when enqueued (the host may run 16 submissions ahead)
t_enq = (time.perf_counter() - t0) * 1e3 / 12
sim.sync()
t_all = (time.perf_counter() - t0) * 1e3 / 12
if rank == 0:
    print("every: host enqueue %.3f ms/step, enqueue+execute %.3f ms/step" % (t_enq, t_all))
if rank == 0:
    print("every", every, "per-step ms:", " ".join("%.2f" % t for t in ts))
sim.close(); dist.barrier(); dist.destroy_process_group()
