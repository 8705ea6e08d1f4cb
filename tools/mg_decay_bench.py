"""Tuning aid: does the slab step slow down as the run proceeds?  VirtualCluster on one device, stage times per window.
usage: python tools/mg_decay_bench.py [world] [atoms_total] [windows] [steps_per_window]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import __graft_entry__ as g
from bench import make_workload
pkg = g.load_package()
mg = importlib.import_module(pkg.__name__ + ".multigpu")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ntot = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
windows = int(sys.argv[3]) if len(sys.argv) > 3 else 4
per = int(sys.argv[4]) if len(sys.argv) > 4 else 100
m = int(round(ntot ** (1 / 3)))
while (m ** 3) % world:
    m += 1
w = make_workload("c4", m ** 3)
vc = mg.VirtualCluster(pkg, w, world, migrate_every=int(os.environ.get("NB200_MIGRATE_EVERY", "0")))
done = 0
for k in range(windows):
    vc.step_async(per - 10)
    for s in vc.sims:
        s.h.set_profiling(True)
    vc.step_async(10)
    done += per
    s = vc.sims[0]
    st = s.h.get_stage_times()
    stats = s.h.get_stats()
    print("after", done, "steps: slab 0 owned", s.h.mg_owned_count(), "ghosts", s.n_ghost, "entries", s.n_entries, "tiles", stats["n_slots"] // 64, "groups", stats["n_segments"],
          {a: round(b[0] / 10, 4) for a, b in st.items() if b[1] > 0}, flush=True)
    if os.environ.get("NB200_MIGRATE_EVERY", "0") != "0":   # rows of mg_get_owned are the current curve order once atoms migrated
        import numpy as np
        x = vc.sims[0].h.mg_get_owned(0)
        nl = len(x) // 32
        xb = x[: nl * 32].reshape(nl, 32, 3)
        ext = (xb.max(1) - xb.min(1)).max(1) / w["cutoff"]
        print("   slab 0 leaves: extent/cutoff mean %.2f max %.2f, wider than 3 cutoffs: %d, than 10: %d" % (ext.mean(), ext.max(), (ext > 3).sum(), (ext > 10).sum()), flush=True)
    for s in vc.sims:
        s.h.set_profiling(False)
