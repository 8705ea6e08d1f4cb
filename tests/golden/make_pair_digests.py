"""Regenerates tests/golden/pair_digests.json: order-independent digests of the exact neighbour pair set
(reference predicate fl(fl(fl(dx^2)+fl(dy^2))+fl(dz^2)) < fl(r*r), BVHTraverse.jl:1026-1027,1248) for seeded inputs.

The reference cannot run here (no Julia) and ships no pair-list fixture of its own (its fixture file is git-ignored,
SURVEY.md 8c), so these vectors come from the CPU oracle: the all-pairs loop up to 20k atoms, the independent
cell-grid search at 1M.  digest = (count, xor and sum of a 64-bit hash of (min id, max id, bits(d)), ids 1-based).

Usage (from the repo root):  python tests/golden/make_pair_digests.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

CASES = {
    # name: (n, seed, cutoff, method)
    "c1_10k_uniform_r0.1": (10_000, 20250313, 0.1, "brute"),
    "uniform_5000_r0.05": (5000, 5000, 0.05, "brute"),
    "uniform_20000_r0.03": (20_000, 21, 0.03, "brute"),
    "c3_size_1m_uniform_seed3": (1_000_000, 3, float(np.float32(2.5 * (0.8 / 1_000_000) ** (1 / 3))), "cellgrid"),
}


def positions(n, seed):
    """uniform [0,1)^3: Float64 draw -> Float32, as generate_positions (MDInput.jl:179-187); = tests/conftest.py"""
    return np.random.default_rng(seed).random((n, 3)).astype(np.float32)


def digest(oracle, n, seed, cutoff, method):
    x = positions(n, seed)
    if method == "brute":
        d = oracle.digest_pairs(*oracle.brute_force(x, cutoff, "d2"))
    else:
        d = oracle.cellgrid_digest(x, np.float32(cutoff))
    return {"n": n, "seed": seed, "cutoff": float(np.float32(cutoff)), "method": method,
            "count": int(d["count"]), "xor": int(d["xor"]), "sum": int(d["sum"])}


if __name__ == "__main__":
    O = graft.load_oracle()
    out = {name: digest(O, *case) for name, case in CASES.items()}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pair_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
