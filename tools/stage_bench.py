"""Tuning aid: per-stage CUDA-event times of the device step loop for one library build.
usage: NAIVEB200_LIB=path/to/variant.so python tools/stage_bench.py [workload] [steps] [n_atoms]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from bench import make_workload
pkg = g.load_package()
w = make_workload(sys.argv[1] if len(sys.argv) > 1 else "c3", int(sys.argv[3]) if len(sys.argv) > 3 else 0)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
h = pkg.Handle(w["n"])
h.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
fused = os.environ.get("NB200_FUSED", "1") != "0"
h.set_fused_force(fused)
h.set_system(w["pos"], w["vel"], w["mass"], w["charge"])
h.step(int(os.environ.get("NB200_PRESTEPS", "5")), w["dt"])
h.set_profiling(True)
h.timer_start(); h.step_async(steps, w["dt"]); ms = h.timer_stop(); h.sync()
st = h.get_stage_times()
h.set_profiling(False)
h.timer_start(); h.step_async(steps, w["dt"]); ms0 = h.timer_stop(); h.sync()
print(os.path.basename(os.environ.get("NAIVEB200_LIB", "default")), "fused" if fused else "unfused", "ms/step %.4f (no events %.4f)" % (ms / steps, ms0 / steps),
      {k: round(v[0] / steps, 4) for k, v in st.items() if v[1] > 0}, "segments", h.get_stats()["n_segments"], "words/pair %.2f" % (h.get_stats()["n_slots"] / max(1, h.get_stats()["n_entries"])), flush=True)
