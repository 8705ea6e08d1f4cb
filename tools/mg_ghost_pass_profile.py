"""Two virtual ranks (1 M atoms each) on one GPU for an ncu look at the owned pass and the ghost pass of the slab step:
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio \
      -k regex:traverse_kernel -s 4 -c 8 python tools/mg_ghost_pass_profile.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from bench import make_workload  # noqa: E402

pkg = graft.load_package()
mg = importlib.import_module(pkg.__name__ + ".multigpu")
world = int(os.environ.get("WORLD", "2"))
w = make_workload("c4", 1_000_000 * world)
vc = mg.VirtualCluster(pkg, w, world, migrate_every=20, headroom=1.6)
vc.step_async(int(os.environ.get("STEPS", "6")))
print("ghosts", [s.n_ghost for s in vc.sims], "entries", [s.n_entries for s in vc.sims])
