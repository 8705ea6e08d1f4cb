#!/usr/bin/env python
"""bench.py — atom-steps/sec of the MD hot path (BVH neighbour search + LJ/Coulomb force + Verlet).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1|random]

One "step" = kick-drift(+wall reflection) -> 30-bit Morton -> radix sort -> gather -> LBVH build ->
traversal (neighbour list rebuilt) -> pair forces, on the workload BASELINE.json quotes the metric on:
config 3, 1,000,000 LJ + Coulomb point charges, rho* = 0.8, cutoff 2.5 sigma, Float32, one B200.
Prints ONE JSON line (see the keys at the bottom).  N > 1 is launched by torchrun (one rank per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

METRIC = "atom-steps/sec (BVH neighbour search + LJ force + Verlet) at 1M particles"
UNIT = "atom-steps/s"


# ---------------------------------------------------------------------------------------------------
# workloads (synthetic)
# ---------------------------------------------------------------------------------------------------
def make_workload(name: str, n_override: int = 0, seed: int = 3):
    """Returns dict(pos, vel, mass, charge, sigma, cutoff, eps, kcoul, dt, desc)."""
    rng = np.random.default_rng(seed)
    if name in ("c3", "c2", "c4"):
        m = {"c3": 100, "c2": 46, "c4": 200}[name]
        if n_override:
            m = int(round(n_override ** (1 / 3)))
        n = m ** 3
        rho = 0.8 if name != "c2" else 0.8442
        margin = 0.02
        a = (1 - 2 * margin) / m                 # lattice constant in box units
        sigma = a * rho ** (1 / 3)               # rho* = sigma^3 / a^3
        g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / m
        pos = (margin + (1 - 2 * margin) * g + 0.05 * a * (rng.random(g.shape) - 0.5)).astype(np.float32)
        perm = rng.permutation(n)                # callers do not hand over Morton-sorted atoms
        pos = pos[perm]
        temp = 0.72
        vel = rng.standard_normal((n, 3)) * np.sqrt(temp)
        vel -= vel.mean(0)
        vel = (vel * sigma).astype(np.float32)   # box units per tau
        mass = np.full(n, 1.0 / sigma ** 2, np.float32)  # m* = 1 (lengths in box units, time in tau)
        if name == "c2":
            charge, kcoul = None, 0.0
        else:
            q = rng.uniform(-1, 1, n)
            q -= q.mean()
            charge = (0.1 * q).astype(np.float32)
            kcoul = float(sigma)                 # q* = 0.1: kc q q / r in eps units with r in box units
        return dict(name=name, pos=pos, vel=vel, mass=mass, charge=charge, sigma=float(sigma), cutoff=float(2.5 * sigma),
                    eps=1.0, kcoul=kcoul, dt=0.005, n=n,
                    desc=f"{n}-atom jittered simple-cubic LJ{'+Coulomb' if charge is not None else ''} box, rho*={rho}, "
                         f"rc=2.5sigma, T*=0.72, dt=0.005tau, reflective walls, neighbour rebuild every step")
    if name in ("c1", "random"):
        n = n_override or (10_000 if name == "c1" else 1_000_000)
        pos = np.random.default_rng(20250313).random((n, 3)).astype(np.float32)
        cutoff = 0.1 if name == "c1" else float(2.5 * (0.8 / n) ** (1 / 3))
        sigma = cutoff / 2.5
        return dict(name=name, pos=pos, vel=np.zeros((n, 3), np.float32), mass=np.full(n, 1.0 / sigma ** 2, np.float32),
                    charge=None, sigma=sigma, cutoff=cutoff, eps=0.0, kcoul=0.0, dt=0.0, n=n,
                    desc=f"{n} uniform-random points in the unit box, r={cutoff:.5f}: search + force-free step")
    if name == "c5":
        # BASELINE config 5: dilute/clustered random gas — half uniform background, half in Gaussian clusters
        # (sigma_cluster = 0.01, centres uniform, seed 5), clamped to [0,1)^3, r chosen for a mean of ~20 neighbours;
        # search + force, a stress test of traversal imbalance.  64M atoms at full size (n_override scales it down;
        # the cluster count scales with n so the per-cluster population stays ~7800).
        n = n_override or 64_000_000
        r5 = np.random.default_rng(5)
        nc = max(1, n // 15625)                      # 4096 clusters at 64M atoms
        nb = n // 2
        pos = np.empty((n, 3), np.float32)
        pos[:nb] = r5.random((nb, 3), dtype=np.float32)
        centres = r5.random((nc, 3))
        which = r5.integers(0, nc, n - nb)
        pos[nb:] = (centres[which] + 0.01 * r5.standard_normal((n - nb, 3))).astype(np.float32)
        np.clip(pos, 0.0, np.float32(1.0 - 2 ** -24), out=pos)
        # atom-weighted mean density: background n/2 everywhere, clusters add N_c / (8 pi^1.5 s^3) on average to their own atoms
        rho_b = n / 2
        rho_c = ((n - nb) / nc) / (8 * np.pi ** 1.5 * 0.01 ** 3)
        rho_mean = 0.5 * rho_b + 0.5 * (rho_b + rho_c)
        cutoff = float((20.0 / (4.0 / 3.0 * np.pi * rho_mean)) ** (1 / 3))
        sigma = cutoff / 2.5
        vel = np.zeros((n, 3), np.float32)
        # soft pair model for the force pass: the random gas has overlapping atoms, so a pure Coulomb-like 1/r term with
        # a tiny coupling keeps the numbers finite (eps = 0 switches the r^-12 core off)
        charge = r5.uniform(-1, 1, n).astype(np.float32)
        return dict(name=name, pos=pos, vel=vel, mass=np.ones(n, np.float32), charge=charge, sigma=sigma, cutoff=cutoff,
                    eps=0.0, kcoul=1e-6, dt=1e-6, n=n,
                    desc=f"{n}-atom dilute/clustered random gas ({nc} Gaussian clusters of sigma 0.01 + uniform background), "
                         f"r={cutoff:.5f} (~20 neighbours per atom on average), search + Coulomb force every step")
    raise SystemExit(f"unknown workload {name}")


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled every ~10 ms through NVML while the timed region runs
    (nvidia-smi's own query path; the subprocess form is the fallback when pynvml is unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, device_index: int):
        self.dev = device_index
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv = self._nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)))
        mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [c.strip() for c in out.split(",")]
        self.sm.append(float(r[0]))
        self.mx.append(float(r[1]))
        for k, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                self.reasons.add(nm)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml is not None else 0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle's reference-shaped step on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_reference_rate(w, budget_s: float, steps: int = 1, warmup: int = 0):
    """Times the reference's CPU algorithm (oracle restatement: 10-bit key, serial sort, threaded
    Apetrei build, leaf traversal with atomsperleaf=4, serial pair forces, Verlet + reflect) on the box's
    host cores.  Every step is a BOUNDED SAMPLE of the workload: the traversal visits every `qstride`-th query leaf,
    chosen so that all `warmup + steps` sample steps together stay inside `budget_s`; traversal and force time are scaled
    back by qstride for the estimate of a full step, build/sort/Verlet are timed in full.
    Returns (cpu_baseline dict, estimated seconds of one full step, measured seconds of one sample step)."""
    O = graft.load_oracle()
    threads = O.hardware_threads()
    n = w["n"]
    apl = 4
    while n % apl:
        apl -= 1
    apl = max(apl, 2)
    pos, vel = w["pos"].copy(), w["vel"].copy()
    force = np.zeros_like(pos)
    from math import gcd

    def coprime(q):
        # the oracle hands query leaf i to thread i mod T (the reference's static round-robin, BVHTraverse.jl:1255-1256)
        # and samples leaves 0, q, 2q, ...: a stride sharing a factor with T would leave threads idle and inflate the time
        while gcd(q, threads) != 1:
            q += 1
        return q

    def one(q):
        t1 = time.perf_counter()
        npairs, tm = O.cpu_step(pos.copy(), vel.copy(), force.copy(), w["mass"], w["charge"], w["dt"], w["cutoff"], apl, threads, q,
                                w["eps"], w["sigma"], w["kcoul"], (0, 0, 0), (1, 1, 1))
        return time.perf_counter() - t1, tm
    # probe with a sparse sample to pick qstride
    probe = coprime(max(1, (n // apl) // 4096))
    t0 = time.time()
    _, tm = one(probe)
    per_leaf = (tm[1] + tm[2]) * probe  # estimated full traverse+force seconds
    fixed = tm[0] + tm[3]
    per_step = budget_s / max(steps + warmup, 1)
    qstride = int(max(1, np.ceil(per_leaf / max(per_step - fixed, 0.5 if steps + warmup <= 3 else 0.15))))
    qstride = coprime(min(qstride, max(1, (n // apl) // (64 * threads))))  # at least ~64 sampled query leaves per thread
    for _ in range(warmup):
        one(qstride)
    est, walls = [], []
    for _ in range(steps):
        wall, tm = one(qstride)
        walls.append(wall)
        est.append(tm[0] + tm[3] + (tm[1] + tm[2]) * qstride)
    step_s = float(np.median(est))
    return dict(value=n / step_s, unit=UNIT, cores=threads, kind="port", estimated=qstride > 1,
                sample=f"oracle C++ restatement of the reference CPU path (no Julia in the image), {threads} threads, "
                       f"atomsperleaf={apl}: {warmup} + {steps} sample steps, each = tree build + Verlet in full, traversal+force over "
                       f"every {qstride}-th query leaf (scaled x{qstride} in the estimate); est. {step_s:.2f} s per full step, "
                       f"{np.mean(walls):.2f} s per sample step; wall {time.time() - t0:.1f} s"), step_s, float(np.mean(walls))


def cpu_measured_suite(budget_s: float = 25.0):
    """MEASURED (unsampled) CPU numbers of the oracle port beside the estimated 1M-atom figure:
    the published sanity anchor (1024 atoms, r = 0.03, 10 force-free steps: 5.010 ms on the author's 8-thread laptop,
    assets/BVHBenchSuite.jl:188-193), BASELINE config 1 (10k uniform points, r = 0.1: build + traverse) and config 2
    (97 336-atom LJ fluid, 3 full steps), with T = all host threads and T = 8 (the author's thread count)."""
    O = graft.load_oracle()
    t_all = O.hardware_threads()
    out = {"host_threads": t_all}
    t_begin = time.perf_counter()
    rng = np.random.default_rng(1)
    xa = rng.random((1024, 3)).astype(np.float32)
    va = (rng.standard_normal((1024, 3)) * 0.01).astype(np.float32)
    ma = np.ones(1024, np.float32)
    c1 = make_workload("c1")
    for T in sorted({t_all, min(8, t_all)}, reverse=True):
        best = 1e30
        for _ in range(5):
            p, v, f = xa.copy(), va.copy(), np.zeros_like(xa)
            t0 = time.perf_counter()
            for _s in range(10):
                O.cpu_step(p, v, f, ma, None, 1.0, 0.03, 4, T, 1, 0.0, 1.0, 0.0, (0, 0, 0), (1, 1, 1))
            best = min(best, time.perf_counter() - t0)
        out[f"anchor_1024x10_r0.03_T{T}"] = {"ms": best * 1e3, "atom_steps_per_s": 10240 / best, "published_ms": 5.010,
                                             "published_on": "i5-9300H, 8 Julia threads (assets/BVHBenchSuite.jl:188-193)"}
        best = 1e30
        for _ in range(6):
            t0 = time.perf_counter()
            a, _, _ = O.leafbuild_traverse_bvh(c1["pos"], c1["cutoff"], 4, T)
            best = min(best, time.perf_counter() - t0)
        out[f"c1_search_10k_T{T}"] = {"ms_per_search": best * 1e3, "searches_per_s": 1.0 / best, "unique_pairs": int(len(a)),
                                      "note": "leafbuild_traverse_bvh restatement, atomsperleaf = 4, best of 5 after 1 warm-up, unsampled"}
    w2 = make_workload("c2")
    p, v, f = w2["pos"].copy(), w2["vel"].copy(), np.zeros_like(w2["pos"])
    steps, ts = 0, []
    while steps < 3 or (steps < 6 and time.perf_counter() - t_begin < budget_s):
        t0 = time.perf_counter()
        npairs, _ = O.cpu_step(p, v, f, w2["mass"], None, w2["dt"], w2["cutoff"], 4, t_all, 1, w2["eps"], w2["sigma"], 0.0, (0, 0, 0), (1, 1, 1))
        ts.append(time.perf_counter() - t0)
        steps += 1
    out[f"c2_100k_md_T{t_all}"] = {"s_per_step": float(np.median(ts)), "atom_steps_per_s": w2["n"] / float(np.median(ts)), "steps": steps,
                                   "n_atoms": w2["n"], "unique_pairs": int(npairs), "note": "full reference-shaped steps, nothing sampled or scaled"}
    out["wall_s"] = time.perf_counter() - t_begin
    return out


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch  # device plumbing for the pinned buffers and, for N > 1, torch.distributed

    pkg = graft.load_package()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from importlib import import_module
        mg = import_module(graft.PKG_NAME + ".multigpu")
        return mg.bench_multi(args, make_workload, METRIC, UNIT, ClockSampler, measured_peak_hbm, emit)

    torch.cuda.set_device(0)
    w = make_workload(args.workload, args.n)
    n = w["n"]
    h = pkg.Handle(n, device=0)
    h.set_box((0, 0, 0), (1, 1, 1))
    h.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
    h.set_fused_force(bool(args.fused))
    variants = {}

    def timed_loop(hh, steps):
        """K steps enqueued between two CUDA events on the library's stream -> (ms, kernel launches)."""
        l0 = hh.get_stats()["kernel_launches"]
        torch.cuda.synchronize()
        hh.timer_start()
        hh.step_async(steps, w["dt"])
        ms_ = hh.timer_stop()
        hh.sync()
        torch.cuda.synchronize()
        return ms_, hh.get_stats()["kernel_launches"] - l0

    # ---- the lattice start (what round 1 timed) as a variant; then MELT: the headline is timed on an equilibrated,
    # disordered liquid, not on the friendliest possible box ----
    h.set_system(w["pos"], w["vel"], w["mass"], w["charge"])
    melt = args.melt if w["eps"] != 0.0 else 0
    if melt > 0:
        h.step(args.warmup, w["dt"])
        ms_l, _ = timed_loop(h, min(args.steps, 100))
        variants["lattice_start"] = {"value": n * min(args.steps, 100) / (ms_l * 1e-3), "ms_per_step": ms_l / min(args.steps, 100),
                                     "note": "the same loop timed from the 5 %-jittered simple-cubic lattice, before melting"}
        h.step(melt, w["dt"])
    npairs0 = h.pair_count()

    # ---- device-resident loop: `value` ----
    # Steady-state steps are replayed as a CUDA graph (two captured steps), no events inside the timed region.
    h.step(args.warmup, w["dt"])
    with ClockSampler(0) as clk:
        ms, launches = timed_loop(h, args.steps)
    ms_per_step = ms / args.steps
    value = n * args.steps / (ms * 1e-3)
    # ---- the same K steps again with the dominant kernel bracketed by CUDA events on the launching stream (an event
    # pair costs a few us of GPU idle per step and rules the graph out, so this is not the headline run) ----
    h.set_profiling(True, only_stage="traverse")
    ms_inst, _ = timed_loop(h, args.steps)
    dom_stage = h.get_stage_times()["traverse"]
    split_steps = max(20, min(args.steps, 100))
    h.set_profiling(True)
    ms_split, _ = timed_loop(h, split_steps)
    stages = h.get_stage_times()
    h.set_profiling(False)
    st = h.get_stats()
    npairs = st["n_pairs"]
    nentries = st["n_entries"]  # set mask bits of the tile list (half list: one per pair)
    nwords = st["n_slots"]      # 4-byte words of the tile list (64 per tile)
    ke, pe = h.get_energies()
    fused = bool(args.fused) and w["eps"] != 0.0

    # ---- the same loop with the Hilbert re-sort every 8th step and a leaf-box refresh in between (the reference's
    # TreeData! update path; the list is still rebuilt from scratch every step): reported next to, not as, `value`
    h.set_resort_interval(8)
    h.step(8, w["dt"])
    ms8, _ = timed_loop(h, args.steps)
    h.set_resort_interval(1)
    h.step(1, w["dt"])
    variants["resort_every_8_steps"] = {"value": n * args.steps / (ms8 * 1e-3), "ms_per_step": ms8 / args.steps,
                                        "note": "atoms re-sorted along the Hilbert curve every 8th step, leaf boxes refreshed "
                                                "and tree + list rebuilt every step (nb200_set_resort_interval)"}
    # ---- forces by the separate tile kernel instead of inside the traversal
    h.set_fused_force(False)
    h.step(2, w["dt"])
    msu, _ = timed_loop(h, min(args.steps, 100))
    h.set_fused_force(bool(args.fused))
    h.step(1, w["dt"])
    variants["unfused_force_kernel"] = {"value": n * min(args.steps, 100) / (msu * 1e-3), "ms_per_step": msu / min(args.steps, 100),
                                        "note": "nb200_set_fused_force(0): the traversal only writes the tile list, a second kernel reads it back"}

    # ---- adjacent component (list reuse across steps): skin list rebuilt every 8th step, exact predicate re-applied
    # by the force kernel; NOT the headline (which rebuilds every step), reported for orientation
    try:
        hr = pkg.Handle(n, device=0, pair_capacity_hint=int(npairs * 2.2))
        hr.set_box((0, 0, 0), (1, 1, 1))
        hr.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
        hr.set_list_reuse(0.5 * w["sigma"], 8)
        hr.set_system(h.get_positions(), h.get_velocities(), w["mass"], w["charge"])
        hr.step(8, w["dt"])
        torch.cuda.synchronize()
        hr.timer_start()
        hr.step_async(args.steps // 8 * 8, w["dt"])
        msr = hr.timer_stop()
        hr.sync()  # raises if an atom moved more than skin/2 while a list was reused
        variants["list_reuse_every_8_steps_skin_0.5sigma"] = {
            "value": n * (args.steps // 8 * 8) / (msr * 1e-3), "ms_per_step": msr / (args.steps // 8 * 8),
            "note": "nb200_set_list_reuse: list built with cutoff + skin every 8th step, forces every step over the exact pair set"}
        hr.close()
    except Exception as exc:  # a variant must never cost the headline line
        variants["list_reuse_every_8_steps_skin_0.5sigma"] = {"error": str(exc)[:200]}

    # ---- BASELINE config 1 (the reference's own CPU-runnable case, BVHBenchSuite-style): neighbour searches per second
    # through the one-call entry point with HOST positions in, pair count out (10k uniform points, r = 0.1)
    c1 = make_workload("c1")
    hc = pkg.Handle(c1["n"], device=0)
    for _ in range(3):
        hc.neighbors(c1["pos"], c1["cutoff"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        c1_pairs = hc.neighbors(c1["pos"], c1["cutoff"])
    c1_dt = (time.perf_counter() - t0) / reps
    hc.close()
    variants["c1_search_10k"] = {"searches_per_s": 1.0 / c1_dt, "ms_per_search": c1_dt * 1e3, "unique_pairs": int(c1_pairs),
                                 "note": "nb200_neighbors (H2D positions + Morton + sort + LBVH + traversal + count readback), host wall clock; "
                                         "the CPU port's time for the same search is in cpu_baseline.measured"}

    # ---- the system drawn on the device (collect_objects with its minimum-distance re-draw), same size and density
    try:
        hs = pkg.Handle(n, device=0, pair_capacity_hint=int(npairs * 1.2))
        hs.set_box((0, 0, 0), (1, 1, 1))
        hs.set_forcefield(w["eps"], w["sigma"], 0.0, w["cutoff"], True)
        best = None
        for seed in (1, 2, 3):
            t0 = time.perf_counter()
            _, _, rounds, redrawn = hs.collect_objects(n, seed, 1.0, 1.0, -1.0, 1.0, 0.72, True, 0.7 * w["sigma"])
            dts = time.perf_counter() - t0
            if best is None or dts < best[0]:
                best = (dts, rounds, redrawn)
        hs.close()
        variants["device_setup_collect_objects"] = {
            "seconds": best[0], "redraw_rounds": best[1], "atoms_redrawn": best[2], "minimumdistance": "0.7 sigma",
            "note": "nb200_collect_objects: Philox draws + BVH-search re-draw rounds + first forces, host wall clock, best of 3 seeds"}
    except Exception as exc:
        variants["device_setup_collect_objects"] = {"error": str(exc)[:200]}

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peak_hbm()
    dom = max((s for s in stages if stages[s][1] > 0), key=lambda s: stages[s][0])
    ntiles = nwords / 64.0
    nl = n / 32.0
    # algorithmic bytes per launch.  The fused traversal reads the positions (16 B/atom), every tree
    # node and leaf box once (64 + 32 B per leaf), writes the tile list (256 B per tile + 16 B per group) and adds the
    # forces (16 B read-modify-write per atom); the un-fused traversal stops before the forces.
    trav_bytes = 16.0 * n + 96.0 * nl + 256.0 * ntiles + 16.0 * st["n_segments"]
    per_launch_bytes = {"integrate": 88.0 * n, "sort": (4.0 + 16.0 * 2) * n / 3, "reorder": 92.0 * n, "build": 136.0 * nl,
                        "traverse": trav_bytes + (32.0 * n if fused else 0.0), "force": 256.0 * ntiles + 16.0 * n + 32.0 * n}
    if dom == "traverse":  # bracketed alone inside a timed region of the same K steps
        dom_ms = dom_stage[0] / max(dom_stage[1], 1)
        dom_share = dom_stage[0] / ms_inst
    else:
        dom_ms = stages[dom][0] / max(stages[dom][1], 1)
        dom_share = stages[dom][0] / ms_split
    achieved = per_launch_bytes.get(dom, 0.0) / (dom_ms * 1e-3) / 1e9
    step_bytes = 440.0 * n + 16.0 * npairs
    survey_kernel_bytes = 112.0 * n + 16.0 * npairs if (dom == "traverse" and fused) else None  # SURVEY 8(d) S5 + S6 constants
    traffic, traffic_note = None, "no ncu capture of this build committed"
    try:  # DRAM bytes per launch of that kernel from the committed ncu --set full capture — only if it was taken on THIS source
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        import hashlib
        src = open(os.path.join(ROOT, "naivedynamics.jl_b200", "csrc", "traverse.cu"), "rb").read()
        if w["name"] == "c3" and tj.get("traverse_cu_sha1") == hashlib.sha1(src).hexdigest():
            traffic = tj.get(dom + "_kernel")
            traffic_note = tj.get("note")
        else:
            traffic_note = "profiles/r2_traffic.json was captured on a different traverse.cu or workload: not reported"
    except Exception:
        pass
    dom_kernel = "traverse_kernel<fused forces>" if (dom == "traverse" and fused) else dom + "_kernel"
    roofline = {"bound": "hbm", "kernel": dom_kernel, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_launch_bytes.get(dom),
                "kernel_ms_per_launch": round(dom_ms, 4),
                "kernel_share_of_step": round(dom_share, 3),
                "limiter": "instruction issue, not HBM (see profiles/): the roofline fraction of this kernel stays small by nature",
                "frac_with_survey_8d_constants": (round(survey_kernel_bytes / (dom_ms * 1e-3) / 1e9 / peak, 4) if survey_kernel_bytes else None),
                "whole_step": {"algorithmic_bytes": step_bytes, "achieved": round(step_bytes / (ms_per_step * 1e-3) / 1e9, 1),
                               "frac": round(step_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4)},
                "stage_ms_per_step": {s_: round(stages[s_][0] / split_steps, 4) for s_ in stages if stages[s_][1] > 0},
                "instrumented_ms_per_step": round(ms_inst / args.steps, 4),
                "stage_split_note": f"headline region: graph replay, no events. kernel_ms_per_launch: a second region of the same {args.steps} steps "
                                    f"with the traversal bracketed by events; stage split: a third run of {split_steps} steps with every stage bracketed "
                                    f"({ms_split / split_steps:.4f} ms/step with that instrumentation)"}

    # ---- end to end through the C ABI with HOST buffers: `e2e` ----
    # ONE system (the headline): nb200_leapfrog_host_async + nb200_sync, positions-only exchange — every step uploads
    # x(t) from pinned host memory, rebuilds the list, evaluates the forces, integrates, and downloads x(t+dt); the
    # velocities stay resident (what simulate!'s poslog contract moves per step, Simulator.jl:245).  Strictly serial:
    # the next step starts from the array the previous one wrote.
    xh = torch.from_numpy(h.get_positions()).pin_memory()
    vh = torch.from_numpy(h.get_velocities()).pin_memory()

    def e2e_single(k, with_vel):
        torch.cuda.synchronize()
        t0_ = time.perf_counter()
        for _ in range(k):
            h.leapfrog_host_async(xh.data_ptr(), vh.data_ptr() if with_vel else 0, 3, n, w["dt"], True)
            h.sync()
        torch.cuda.synchronize()
        return time.perf_counter() - t0_

    e2e_steps = max(6, min(args.steps, 60))
    e2e_single(3, False)
    wall = e2e_single(e2e_steps, False)
    e2e_val = n * e2e_steps / wall
    vh.copy_(torch.from_numpy(h.get_velocities()))  # synchronised v(t); the first x+v call moves them to the half step
    h.leapfrog_host_async(xh.data_ptr(), vh.data_ptr(), 3, n, w["dt"], False)
    h.sync()
    e2e_single(2, True)
    wall_xv = e2e_single(max(6, e2e_steps // 2), True)
    # copy bandwidth beside it: the PCIe bound of the positions-only step
    tb = torch.empty(n * 3, dtype=torch.float32, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(3):  # warm-up: first copies carry one-time costs
        tb.copy_(xh.view(-1), non_blocking=True)
        xh.view(-1).copy_(tb, non_blocking=True)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(20):
        tb.copy_(xh.view(-1), non_blocking=True)
    ev[1].record()
    for _ in range(20):
        xh.view(-1).copy_(tb, non_blocking=True)
    ev[2].record()
    torch.cuda.synchronize()
    h2d = 20 * 12.0 * n / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
    d2h = 20 * 12.0 * n / (ev[1].elapsed_time(ev[2]) * 1e-3) / 1e9
    # R independent replicas round-robin: copies of one replica run under the kernels of the others (aggregate, NOT the headline)
    R = 3
    reps = [h]
    for r in range(1, R):
        hr = pkg.Handle(n, device=0, pair_capacity_hint=int(npairs * 1.15))
        hr.set_box((0, 0, 0), (1, 1, 1))
        hr.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
        hr.set_system(xh.numpy(), np.roll(vh.numpy(), r, axis=0), w["mass"], w["charge"])
        reps.append(hr)
    xr = [xh] + [torch.from_numpy(hr.get_positions()).pin_memory() for hr in reps[1:]]

    def e2e_loop(k0, k1):
        for it in range(k0, k1):
            r = it % R
            reps[r].sync()  # this replica's previous step (and its D2H) is complete
            reps[r].leapfrog_host_async(xr[r].data_ptr(), 0, 3, n, w["dt"], True)
        for hr in reps:
            hr.sync()

    agg_steps = e2e_steps // R * R
    e2e_loop(0, 2 * R)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_loop(2 * R, 2 * R + agg_steps)
    torch.cuda.synchronize()
    wall_agg = time.perf_counter() - t0
    for hr in reps[1:]:
        hr.close()
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 12 * n, "d2h_bytes_per_step": 12 * n,
           "steps": e2e_steps, "ms_per_step": wall / e2e_steps * 1e3, "timer": "host wall clock around the loop, device idle on both sides",
           "api": "nb200_leapfrog_host_async(xyz, vel = NULL) + nb200_sync (C ABI, pinned host buffer), ONE system, strictly serial",
           "pcie_measured_gbs": {"h2d": round(h2d, 1), "d2h": round(d2h, 1)},
           "pcie_bound_ms_per_step": round(12.0 * n / (h2d * 1e9) * 1e3 + 12.0 * n / (d2h * 1e9) * 1e3, 4),
           "single_system_positions_and_velocities": {"value": n * max(6, e2e_steps // 2) / wall_xv, "h2d_bytes_per_step": 24 * n,
                                                      "d2h_bytes_per_step": 24 * n},
           "aggregate_3_replicas": {"value": n * agg_steps / wall_agg, "note": "three independent systems round-robin; NOT the headline"},
           "note": "every step uploads x, rebuilds the neighbour list, evaluates forces, integrates and downloads x; velocities stay on the device"}

    # ---- CPU baseline (oracle port) ----
    cpu, _, _ = cpu_reference_rate(w, budget_s=args.cpu_budget)
    try:
        cpu["measured"] = cpu_measured_suite()
    except Exception as exc:
        cpu["measured"] = {"error": str(exc)[:200]}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": w["desc"] + (f"; timed after {melt} steps of melting (equilibrated, disordered liquid)" if melt else ""),
                      "name": w["name"], "n_atoms": n, "unique_pairs": int(npairs),
                      "pairs_per_atom": round(npairs / n, 2), "cutoff_box_units": w["cutoff"],
                      "list": ("half" if st["list_half"] else "directed") + ", 32x32 hit-mask tiles, %.2f words per pair" % (nwords / max(nentries, 1)),
                      "forces": "evaluated inside the traversal kernel (tile list still written)" if fused else "separate tile kernel",
                      "step_loop": "two steps captured as a CUDA graph and replayed",
                      "l2_policy": "working set (state 96 MB + tree/keys 23 MB + list %d MB) exceeds the 126 MB L2" % (4 * nwords // 2**20),
                      "parallelism": "single GPU"},
           "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "variants": variants, "gpu_launches": int(launches),
           "clocks": clk.summary(), "energy": {"ke": ke, "pe": pe}, "groups": st["n_segments"], "leaves": st["n_leaves"]}
    emit(json.dumps(out))
    h.close()


def run_reference(args):
    """The reference arm: the reference's own CPU algorithm (oracle port — the reference is pure Julia and there is no Julia
    here) on rank 0's host cores.  `--warmup W` untimed + `--steps K` timed steps, each a bounded sample of the workload
    (cpu_reference_rate); `ms_per_step` is the measured wall time of one SAMPLE step, so steps x ms_per_step is the timed region;
    `value` is the throughput of the whole workload estimated from the samples (cpu_baseline.sample says how)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload, args.n)
    cpu, step_s, sample_s = cpu_reference_rate(w, budget_s=max(60.0, 3.0 * args.cpu_budget), steps=max(1, args.steps), warmup=args.warmup)
    cpu["estimated_full_step_ms"] = step_s * 1e3
    try:
        cpu["measured"] = cpu_measured_suite()
    except Exception as exc:
        cpu["measured"] = {"error": str(exc)[:200]}
    out = {"metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": sample_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "impl": "reference",
           "config": {"workload": w["desc"], "name": w["name"], "n_atoms": w["n"], "parallelism": f"{cpu['cores']} host threads",
                      "note": "the CPU arm always runs the 1M-atom config-3 box on rank 0's host cores, whatever --gpus says (the GPU arm's "
                              "weak-scaling box grows with N); every step is a bounded SAMPLE of that workload (ms_per_step is the time of a "
                              "sample step), the 1M figure is an ESTIMATE from the sampled traversal (cpu_baseline.estimated, "
                              "estimated_full_step_ms); the unsampled measurements (sanity anchor, config 1, config 2) are in "
                              "cpu_baseline.measured"},
           "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(out))


def emit(line: str):
    """The ONE JSON line goes to the process's original stdout (see main)."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1


def main():
    # Libraries (NCCL's version banner, torchrun's notices) print to fd 1: keep stdout for the JSON line alone.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--n", type=int, default=0, help="override the atom count (lattice: rounded to a cube)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--fused", type=int, default=1, help="1: pair forces inside the traversal kernel (default); 0: separate tile force kernel")
    ap.add_argument("--melt", type=int, default=600, help="untimed MD steps before the timed region (lattice -> disordered liquid)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
