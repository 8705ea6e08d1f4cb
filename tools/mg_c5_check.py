"""BASELINE config 5 (clustered gas) on EIGHT virtual ranks of one GPU, at a reduced size: the ragged-slab path of the
asynchronous step (occupancy grid, ghost capacity sized by the first search) with migration, checked against the oracle.
Usage: python tools/mg_c5_check.py [n_atoms]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from bench import make_workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
world = int(os.environ.get("WORLD", "8"))
pkg = graft.load_package()
mg = importlib.import_module(pkg.__name__ + ".multigpu")
O = graft.load_oracle()
w = make_workload("c5", n)
vc = mg.VirtualCluster(pkg, w, world, migrate_every=20, headroom=2.0)
print("n", w["n"], "cutoff", w["cutoff"], "owned", [s.n_own for s in vc.sims], "ghosts(sync)", [s.n_ghost for s in vc.sims], flush=True)
t0 = time.perf_counter()
vc.step_async(45)
print("45 async steps ok in %.2f s; ghosts" % (time.perf_counter() - t0), [s.n_ghost for s in vc.sims], "owned", [s.h.mg_owned_count() for s in vc.sims], flush=True)
x = vc.gather(0)
vc._exchange()
ref = O.cellgrid_digest(x, w["cutoff"])
owner = np.full(w["n"], -1, np.int64)
for g, s in enumerate(vc.sims):
    owner[s.owned_original_ids()] = g
assert np.all(owner >= 0)
cnt, xo, su = 0, 0, 0
for g, (a, b, d) in enumerate(vc.entries()):
    keep = owner[np.minimum(a, b)] == g
    dg = O.digest_pairs(a[keep] + 1, b[keep] + 1, d[keep])
    cnt += dg["count"]; xo ^= dg["xor"]; su = (su + dg["sum"]) & 0xffffffffffffffff
print("pairs", cnt, "oracle", ref["count"], "match", cnt == ref["count"], xo == ref["xor"], su == ref["sum"])
