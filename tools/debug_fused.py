"""Debug aid: fused vs tile-kernel forces against the fp64 oracle after each step (prints the worst atoms)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package(); O = g.load_oracle()
m = 24; rng = np.random.default_rng(11)
gg = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / m
a = 0.8 / m
x = (0.1 + 0.8 * gg + 0.1 * a * (rng.random(gg.shape) - 0.5)).astype(np.float32)
n = len(x); sigma = a / 1.1; rc = 2.5 * sigma
rng = np.random.default_rng(12)
v = (rng.standard_normal((n, 3)) * 0.5 * sigma).astype(np.float32)
for fused in (True, False):
    h = pkg.Handle(n); h.set_fused_force(fused)
    h.set_forcefield(eps=1.0, sigma=sigma, kcoul=0.0, cutoff=rc, shift=True)
    h.set_system(x, v, np.full(n, 1.0 / sigma ** 2, np.float32), None)
    for step in range(3):
        h.step(1, 0.002)
        xs = h.get_positions(); f = h.get_forces()
        pa, pb, pd = h.get_pairs()
        f64, pe64, scale = O.forces_physical_f64(xs, None, pa, pb, 1.0, sigma, 0.0, rc, True)
        err = np.abs(f - f64).max(axis=1) / np.maximum(scale, 1e-30)
        bad = np.nonzero(~(err < 1e-5))[0]
        print("fused", fused, "step", step, "pairs", len(pa), "nan pos", int(np.isnan(xs).sum()), "zero-scale", int((scale == 0).sum()),
              "max err", float(np.nanmax(err)), "bad atoms", len(bad), bad[:8], flush=True)
        for i in bad[:4]:
            print("   atom", i, "x", xs[i], "f", f[i], "f64", f64[i], "scale", scale[i])
    h.close()
