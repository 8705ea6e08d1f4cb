"""Tuning aid: per-warp profile of the traversal kernel on a bench workload (run on the GPU box)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from bench import make_workload
pkg = g.load_package()
w = make_workload(sys.argv[1] if len(sys.argv) > 1 else "c3", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
h = pkg.Handle(w["n"])
h.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
h.set_system(w["pos"], w["vel"], w["mass"], w["charge"])
p = h.debug_traverse_profile()
cyc, cand, rounds, targ = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
def q(a): return [float(np.percentile(a, x)) for x in (50, 90, 99, 99.9, 100)]
print(json.dumps({"leaves": len(p), "cycles_p50_90_99_999_max": q(cyc), "cand": q(cand), "rounds": q(rounds), "targets": q(targ),
                  "sum_cycles": float(cyc.sum()), "mean_cycles": float(cyc.mean()), "mean_cand": float(cand.mean()),
                  "mean_targets": float(targ.mean()),
                  "cyc_per_cand_p50": float(np.median(cyc / np.maximum(cand, 1)))}))
worst = np.argsort(-cyc)[:8]
print("worst leaves", [(int(i), int(cyc[i]), int(cand[i]), int(rounds[i]), int(targ[i])) for i in worst])
