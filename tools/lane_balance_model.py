"""CPU model of the lane balance of the fused force evaluation on the c3 workload (no GPU needed).

For a sample of query leaves it builds the half-list hit rows (pairs with later leaves + the symmetric self tile) and
reports how many loop trips different ways of dealing the pairs out to the 32 lanes of the warp need:
rows (lane = query atom), fixed lane pairs helping each other, count-sorted lane pairs, and the exact even split.
Usage: python tools/lane_balance_model.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from test_gpu_stages import hilbert30_numpy
from scipy.spatial import cKDTree
w = bench.make_workload("c3")
x = w["pos"].astype(np.float32); r = float(w["cutoff"]); n = len(x)
order = np.argsort(hilbert30_numpy(x), kind="stable")
xs = x[order].astype(np.float64)
nL = (n + 31) // 32
rng = np.random.default_rng(0)
sample = rng.choice(nL - 1, 400, replace=False)
tree = cKDTree(xs)
r2 = r * r
acc = {k: [] for k in ("mean", "rows", "blocks", "pair31", "pair1", "pair16", "sorted", "even", "quad_sorted")}
for A in sample:
    a0 = A * 32; q = xs[a0:a0 + 32]
    lo, hi = q.min(0), q.max(0)
    c = (lo + hi) / 2; R = np.linalg.norm((hi - lo) / 2) + r
    idx = np.sort(np.array(tree.query_ball_point(c, R))); idx = idx[idx >= a0 + 32]
    p = xs[idx]
    g = np.maximum(0, np.maximum(lo - p, p - hi)); near = (g * g).sum(-1) <= r2
    p = p[near]
    hit = ((q[:, None, :] - p[None, :, :]) ** 2).sum(-1) < r2          # [32, T]
    self_hit = ((q[:, None, :] - q[None, :, :]) ** 2).sum(-1) < r2
    np.fill_diagonal(self_hit, False)
    rows = hit.sum(1) + self_hit.sum(1)
    acc["mean"].append(rows.mean())
    acc["rows"].append(rows.max())
    # block by block (targets in gather order = sorted slot order, 32 per block), self tile first
    tb = self_hit.sum(1).max()
    for t0 in range(0, hit.shape[1], 32):
        tb += hit[:, t0:t0 + 32].sum(1).max()
    acc["blocks"].append(tb)
    for name, partner in (("pair31", 31 - np.arange(32)), ("pair1", np.arange(32) ^ 1), ("pair16", np.arange(32) ^ 16)):
        acc[name].append(np.ceil((rows + rows[partner]) / 2).max())
    s = np.sort(rows)
    acc["sorted"].append(np.ceil((s + s[::-1]) / 2).max())
    acc["quad_sorted"].append(np.ceil((s[0:8] + s[15:7:-1] + s[16:24] + s[31:23:-1]) / 4).max())
    acc["even"].append(np.ceil(rows.sum() / 32))
m = np.mean(acc["mean"])
print("pairs per lane (mean row): %.1f" % m)
for k in ("blocks", "rows", "pair31", "pair1", "pair16", "sorted", "quad_sorted", "even"):
    print("%-12s trips/leaf %.1f  lane utilisation %.0f%%" % (k, np.mean(acc[k]), 100 * m / np.mean(acc[k])))
