"""Times nb200_collect_objects (device-side system setup) at the C3 size: 1M atoms, rho* = 0.8,
minimum distance 0.7 sigma.  Wall clock around the blocking call (draws, re-draw rounds with one search each, first
forces).  Usage: python tools/setup_bench.py [n] [mindist_in_sigma] > gpurun_out/setup_bench.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
sigma = (0.8 / n) ** (1.0 / 3.0)
h = pkg.Handle(n)
h.set_box((0, 0, 0), (1, 1, 1))
h.set_forcefield(eps=1.0, sigma=sigma, kcoul=0.0, cutoff=2.5 * sigma, shift=True)
out = []
for seed in (1, 2, 3):
    l0 = h.get_stats()["kernel_launches"]
    t0 = time.perf_counter()
    mass, charge, rounds, redrawn = h.collect_objects(n, seed, 1.0, 1.0, -1.0, 1.0, 0.72, True, frac * sigma)
    dt = time.perf_counter() - t0
    out.append(dict(seed=seed, seconds=dt, rounds=rounds, redrawn=redrawn, pairs_at_cutoff=h.pair_count(),
                    launches=h.get_stats()["kernel_launches"] - l0))
print(json.dumps(dict(n=n, minimumdistance_sigma=frac, sigma_box=sigma, runs=out)))
