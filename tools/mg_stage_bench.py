"""Tuning aid: stage times of the multi-GPU slab path with all slabs on ONE device (VirtualCluster, asynchronous steps).
usage: NAIVEB200_LIB=... python tools/mg_stage_bench.py [world] [atoms_total] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import __graft_entry__ as g
from bench import make_workload
pkg = g.load_package()
mg = importlib.import_module(pkg.__name__ + ".multigpu")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ntot = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
m = int(round(ntot ** (1 / 3)))
while (m ** 3) % world:
    m += 1
w = make_workload("c4", m ** 3)
vc = mg.VirtualCluster(pkg, w, world)
vc.step_async(5)
for s in vc.sims:
    s.h.set_profiling(True)
vc.step_async(steps)
for k, s in enumerate(vc.sims[:2]):
    st = s.h.get_stage_times()
    print(os.path.basename(os.environ.get("NAIVEB200_LIB", "default")), "slab", k, "n_own", s.n_own, "ghosts", s.n_ghost,
          {a: round(b[0] / steps, 4) for a, b in st.items() if b[1] > 0}, "sum %.4f" % (sum(b[0] for b in st.values()) / steps))
