import sys, os
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
from bench import make_workload
pkg = g.load_package()
w = make_workload("c3")
h = pkg.Handle(w["n"])
h.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
h.set_system(w["pos"], w["vel"], w["mass"], w["charge"])
h.step(10, w["dt"])
for prof in (False, True, False, True):
    h.set_profiling(prof)
    h.timer_start(); h.step_async(200, w["dt"]); ms = h.timer_stop(); h.sync()
    print("profiling", prof, "ms/step %.4f" % (ms / 200))
