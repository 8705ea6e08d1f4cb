"""Debug aid: NVE energy over 1000 steps for {graph, no graph} x {fused, unfused}."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
m = 24; n = m ** 3; rho = 0.8442
L = (n / rho) ** (1 / 3); sigma = 0.8 / L
rng = np.random.default_rng(11)
gg = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / m
a = 0.8 / m
x = (0.1 + 0.8 * gg + 0.02 * a * (rng.random(gg.shape) - 0.5)).astype(np.float32)
rng = np.random.default_rng(1)
v = rng.standard_normal((n, 3)) * np.sqrt(0.72); v -= v.mean(0); v = (v * sigma).astype(np.float32)
mass = np.full(n, 1.0 / sigma ** 2, np.float32)
for graph in (True, False):
    for fused in (True, False):
        if graph: os.environ.pop("NB200_NO_GRAPH", None)
        else: os.environ["NB200_NO_GRAPH"] = "1"
        h = pkg.Handle(n); h.set_fused_force(fused)
        h.set_forcefield(1.0, sigma, 0.0, 2.5 * sigma, True)
        h.set_system(x, v, mass, None)
        ke0, pe0 = h.get_energies(); e0 = ke0 + pe0
        out = []
        for _ in range(10):
            h.step(100, 0.005)
            ke, pe = h.get_energies()
            out.append((ke + pe - e0) / e0)
        print("graph", graph, "fused", fused, "e0 %.6g" % e0, "steps", h.get_stats()["steps_done"], "launches", h.get_stats()["kernel_launches"],
              " ".join("%.2e" % o for o in out), flush=True)
        h.close()
