"""Multi-GPU driver: Morton-slab partition + per-step halo exchange.

One process per GPU.  The reference has no multi-process code at all; this is the B200-native
extension of its hot path:

  * the atoms are ordered by a 30-bit Morton key and cut into `world` contiguous slabs of equal atom
    count; rank g owns slab g for the whole run (atoms do not migrate between ranks in this version;
    the slab boxes simply grow as atoms diffuse);
  * every step each rank kick-drifts its owned atoms and PUBLISHES their positions and the boxes of its
    32-atom publication leaves in a peer-mapped buffer, then every rank pulls as ghosts the foreign atoms
    within the cutoff of its slab's bounding box, builds a local LBVH over owned + ghosts and traverses it;
  * exchange="peer" (default): the pull is the library's own kernel reading the peers' GPU memory over
    NVLink/NVSwitch (CUDA IPC mappings), synchronised only by per-rank step flags — no collective, no
    barrier, ~1.5 MB per rank per step at 1M atoms per GPU;
    exchange="nccl": the packed float4 positions of all ranks are all-gathered with torch.distributed
    (16 B x N per rank per step) and the ghosts are selected from the gathered array;
  * the neighbour list is a half list (every pair with at least one owned atom, once; a pair of two ghosts
    belongs to other ranks and is dropped): the force kernel adds the reaction to the partner, forces that
    land on ghosts are discarded, so every rank has complete forces for its own atoms and there is no
    reverse force reduction.  The union of the ranks' entries is exactly the single-GPU pair set
    (tests/test_multigpu.py checks it against the oracle).

torch is plumbing only: process group, exchange of the IPC handles, the optional all_gather.  All compute
is in libnaiveb200.so.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np


# ----------------------------------------------------------------------------------------------------
# host-side partition logic (pure numpy: testable on CPU with gloo)
# ----------------------------------------------------------------------------------------------------
def _spread10(v):
    v = v.astype(np.uint32) & np.uint32(0x3FF)
    v = (v | (v << np.uint32(16))) & np.uint32(0x030000FF)
    v = (v | (v << np.uint32(8))) & np.uint32(0x0300F00F)
    v = (v | (v << np.uint32(4))) & np.uint32(0x030C30C3)
    v = (v | (v << np.uint32(2))) & np.uint32(0x09249249)
    return v


def morton30(pos, box_min=(0.0, 0.0, 0.0), box_max=(1.0, 1.0, 1.0)):
    """Same key as morton_kernel (csrc/atoms.cu): 10 bits per axis, x in bit 0."""
    lo = np.asarray(box_min, np.float32)
    scale = np.float32(1024.0) / (np.asarray(box_max, np.float32) - lo)
    q = np.clip(np.floor((pos.astype(np.float32) - lo) * scale), 0, 1023).astype(np.uint32)
    return _spread10(q[:, 0]) | (_spread10(q[:, 1]) << np.uint32(1)) | (_spread10(q[:, 2]) << np.uint32(2))


def morton_slab_partition(pos, world, box_min=(0.0, 0.0, 0.0), box_max=(1.0, 1.0, 1.0), equal=True):
    """Returns `owner_order`: atom ids in global Morton order, and `bounds`: world+1 cut positions.
    Rank g owns owner_order[bounds[g]:bounds[g+1]].  equal=True (the NCCL exchange: fixed-size all_gather) requires
    len(pos) to be divisible by world; the peer exchange has no such constraint (equal=False: counts differ by <= 1)."""
    n = len(pos)
    if equal and n % world:
        raise ValueError(f"atom count {n} must be divisible by the number of ranks {world}")
    order = np.argsort(morton30(pos, box_min, box_max), kind="stable")
    if equal:
        bounds = np.arange(world + 1, dtype=np.int64) * (n // world)
    else:
        bounds = np.round(np.linspace(0, n, world + 1)).astype(np.int64)
    return order, bounds


def select_ghosts_reference(all_pos, own_begin, n_own, cutoff):
    """numpy twin of ghost_select_kernel: indices (into the gathered array) of foreign atoms within the
    cutoff of the slab's bounding box."""
    own = all_pos[own_begin:own_begin + n_own, :3]
    lo, hi = own.min(0), own.max(0)
    p = all_pos[:, :3]
    gap = np.maximum(0, np.maximum(lo - p, p - hi)).astype(np.float32)
    r2 = np.float32(cutoff) * np.float32(cutoff)
    near = (gap * gap).sum(1) <= r2 * np.float32(1 + 4e-6) + np.float32(1e-37)
    near[own_begin:own_begin + n_own] = False
    return np.nonzero(near)[0]


GRID = 64  # csrc/slab_grid.cuh


def occupancy_grid_reference(own_pos, cutoff, box_min=(0.0, 0.0, 0.0), box_max=(1.0, 1.0, 1.0)):
    """numpy twin of mg_grid_mark_kernel + mg_grid_dilate_kernel (csrc/peer_exchange.cu): boolean [z, y, x] grid of the
    cells within R = ceil(cutoff / cell) cells (Chebyshev) of a cell that holds an owned atom."""
    lo = np.asarray(box_min, np.float32)
    scale = np.float32(GRID) / (np.asarray(box_max, np.float32) - lo)
    c = np.clip(np.floor((own_pos[:, :3].astype(np.float32) - lo) * scale), 0, GRID - 1).astype(np.int64)
    raw = np.zeros((GRID, GRID, GRID), bool)
    raw[c[:, 2], c[:, 1], c[:, 0]] = True
    R = int(np.ceil(np.float32(cutoff) * scale.max()))
    R = max(R, 1)
    if R > 8:
        return np.ones_like(raw)
    out = np.zeros_like(raw)
    for dz in range(-R, R + 1):
        for dy in range(-R, R + 1):
            for dx in range(-R, R + 1):
                src = raw[max(0, -dz):GRID - max(0, dz), max(0, -dy):GRID - max(0, dy), max(0, -dx):GRID - max(0, dx)]
                out[max(0, dz):GRID - max(0, -dz), max(0, dy):GRID - max(0, -dy), max(0, dx):GRID - max(0, -dx)] |= src
    return out


def grid_lookup_reference(grid, pos, box_min=(0.0, 0.0, 0.0), box_max=(1.0, 1.0, 1.0)):
    """grid_point (csrc/slab_grid.cuh): is the (clamped) cell of each position set?"""
    lo = np.asarray(box_min, np.float32)
    scale = np.float32(GRID) / (np.asarray(box_max, np.float32) - lo)
    c = np.clip(np.floor((pos[:, :3].astype(np.float32) - lo) * scale), 0, GRID - 1).astype(np.int64)
    return grid[c[:, 2], c[:, 1], c[:, 0]]


# ----------------------------------------------------------------------------------------------------
# per-rank simulation object (GPU)
# ----------------------------------------------------------------------------------------------------
class SlabSimulation:
    def __init__(self, pkg, workload, rank, world, device, dist=None, headroom=1.6, defer=False, exchange="peer",
                 list_mode=1):
        import torch

        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.exchange = exchange
        w = workload
        pos = w["pos"]
        n = len(pos)
        self.n_total = n
        order, bounds = morton_slab_partition(pos, world, equal=(exchange == "nccl"))
        mine = order[bounds[rank]:bounds[rank + 1]]
        self.owned_ids = mine
        self.order, self.bounds = order, bounds
        self.n_own = len(mine)
        self.own_begin = int(bounds[rank])
        self.h = pkg.Handle(int(self.n_own * headroom) + 4096, device=device)
        self.h.set_box((0, 0, 0), (1, 1, 1))
        self.h.set_list_mode(list_mode)
        self.h.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
        self.dev = torch.device("cuda", device)
        if exchange == "nccl":
            self.h.set_stream(torch.cuda.current_stream().cuda_stream)  # kernels and the collective share a stream
            self.all_pos = torch.empty((n, 4), dtype=torch.float32, device=self.dev)
        q = None if w["charge"] is None else w["charge"][mine]
        self.h.mg_set_owned(pos[mine], w["vel"][mine], w["mass"][mine], q)
        self.dt = w["dt"]
        self.n_ghost = 0
        self.n_entries = 0
        if exchange == "peer" and dist is not None and world > 1:
            # hand every rank the CUDA IPC handle of every publication (setup only; the step loop has no collective)
            _, _, handle = self.h.mg_publication()
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            self.h.mg_connect(world, rank, bounds[:-1], np.diff(bounds), ipc_handles=handles)
            dist.barrier()
        if not defer:
            self.exchange_and_search()

    def _wrap(self, ptr, rows):
        torch = self.torch

        class _Arr:  # __cuda_array_interface__ shim: lets torch view foreign device memory
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (rows, 4), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
        return torch.as_tensor(a, device=self.dev)

    def send_buffer(self):
        """zero-copy torch view of the library's owned positions of the current step (all-gather send buffer)"""
        return self._wrap(self.h.mg_owned_pos_device(), self.n_own)

    def exchange_and_search(self):
        if self.exchange == "nccl":
            if self.world > 1:
                self.dist.all_gather_into_tensor(self.all_pos, self.send_buffer())
            else:
                self.all_pos.copy_(self.send_buffer())
            self.search(self.all_pos)
        else:
            self.search(None)

    def search(self, all_pos=None):
        if all_pos is None:
            self.n_ghost, self.n_entries = self.h.mg_search_force()
        else:
            self.n_ghost, self.n_entries = self.h.mg_search_force(all_pos.data_ptr(), self.n_total, self.own_begin)

    def integrate(self):
        self.h.mg_integrate(self.dt)

    def step(self, nsteps=1):
        for _ in range(nsteps):
            self.integrate()
            self.exchange_and_search()

    def step_async(self, nsteps=1):
        """Peer exchange only: nothing in the loop waits for the GPU (launches are sized for n_own + ghost capacity);
        call sync() afterwards — it reports a ghost/list overflow or a silent peer for all steps since the last sync."""
        assert self.exchange == "peer"
        for _ in range(nsteps):
            self.h.mg_integrate(self.dt)
            self.h.mg_search_force_async()

    def sync(self):
        self.n_ghost, self.n_entries = self.h.mg_sync()

    def entries_global(self):
        """(a, b, d) with ORIGINAL atom ids (0-based) for this rank's list entries."""
        a, b, d = self.h.mg_get_entries(self.n_entries)
        return self.order[a], self.order[b], d

    def close(self):
        self.h.close()


class VirtualCluster:
    """All `world` slabs on ONE device in one process: exercises exactly the code path of the real multi-GPU run
    (publication, flags, pull kernel or gathered-array ghost selection, half/directed list, force scatter), so
    the parity test does not need several GPUs.  The slabs are stepped in lockstep — all publish, then all pull —
    because a pull kernel waits for flags that only the other slabs' (same-device) kernels can set."""

    def __init__(self, pkg, workload, world, device=0, exchange="peer", list_mode=1, headroom=1.6):
        import torch
        self.torch = torch
        self.exchange = exchange
        self.sims = [SlabSimulation(pkg, workload, g, world, device, dist=None, defer=True, exchange=exchange, list_mode=list_mode,
                                    headroom=headroom) for g in range(world)]
        self.n_total = self.sims[0].n_total
        if exchange == "peer":
            bases = [s.h.mg_publication()[0] for s in self.sims]
            bounds = self.sims[0].bounds
            for g, s in enumerate(self.sims):
                s.h.mg_connect(world, g, bounds[:-1], np.diff(bounds), direct_base=bases)
        else:
            self.all_pos = torch.empty((self.n_total, 4), dtype=torch.float32, device=self.sims[0].dev)
        self._exchange()

    def _exchange(self):
        if self.exchange == "peer":
            for s in self.sims:
                s.search(None)
        else:
            self.torch.cat([s.send_buffer() for s in self.sims], out=self.all_pos)
            for s in self.sims:
                s.search(self.all_pos)

    def step(self, nsteps=1):
        for _ in range(nsteps):
            for s in self.sims:
                s.integrate()
            self._exchange()

    def step_async(self, nsteps=1):
        """the asynchronous step of every slab, in lockstep (all publish, then all pull), one sync at the end"""
        for _ in range(nsteps):
            for s in self.sims:
                s.h.mg_integrate(s.dt)
            for s in self.sims:
                s.h.mg_search_force_async()
        for s in self.sims:
            s.sync()

    def gather(self, mode):
        """positions (0) / velocities (1) / forces (2) of all atoms in ORIGINAL order."""
        out = np.empty((self.n_total, 3), np.float32)
        for s in self.sims:
            out[s.owned_ids] = s.h.mg_get_owned(mode)
        return out

    def entries(self):
        """per-rank (a, b, d) entry lists, ORIGINAL atom ids"""
        return [s.entries_global() for s in self.sims]

    def energies(self):
        e = np.array([s.h.mg_get_energies() for s in self.sims])
        return e[:, 0].sum(), e[:, 1].sum()

    def close(self):
        for s in self.sims:
            s.close()


# ----------------------------------------------------------------------------------------------------
# bench entry (called by bench.py under torchrun)
# ----------------------------------------------------------------------------------------------------
def bench_multi(args, make_workload, METRIC, UNIT, ClockSampler, measured_peak_hbm, emit=print):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as graft
    pkg = graft.load_package()
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # weak scaling: ~1M atoms per GPU (BASELINE config 4 = 8M atoms on 8 GPUs)
    if args.workload == "c5":  # BASELINE config 5: 64M-atom clustered gas on 8 GPUs (8M per GPU), traversal-imbalance stress test
        per_gpu = args.n or 8_000_000
        w = make_workload("c5", per_gpu * world)
    else:
        per_gpu = args.n or 1_000_000
        m = int(round((per_gpu * world) ** (1 / 3)))
        while (m ** 3) % world:
            m += 1
        w = make_workload("c4", m ** 3)
    n = w["n"]
    exchange = os.environ.get("NB200_EXCHANGE", "peer")
    sim = SlabSimulation(pkg, w, rank, world, local_rank, dist, exchange=exchange)
    def run(k):
        if exchange == "peer":
            sim.step_async(k)
            sim.sync()
        else:
            sim.step(k)
    run(args.warmup)
    l0 = sim.h.get_stats()["kernel_launches"]
    if not os.environ.get("NB200_NO_PROFILE"):  # (tuning aid: A/B runs against library builds with the older profiling levels)
        sim.h.set_profiling(True, only_stage="traverse")  # six bracketed stages would cost ~5 % of the step (see bench.py)
    with ClockSampler(local_rank) as clk:
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(torch.cuda.current_stream()) if exchange != "peer" else sim.h.timer_start()
        if exchange == "peer":
            sim.step_async(args.steps)
            ms_lib = sim.h.timer_stop()  # CUDA events on the library's own stream
            sim.sync()
        else:
            sim.step(args.steps)
            ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dist.barrier()
    ms = ms_lib if exchange == "peer" else ev0.elapsed_time(ev1)
    split_steps = 20  # the full stage split comes from a short extra run with every stage bracketed
    sim.h.set_profiling(True)
    run(split_steps)
    stages = sim.h.get_stage_times()
    sim.h.set_profiling(False)
    l1 = sim.h.get_stats()["kernel_launches"]
    t = torch.tensor([ms, wall * 1e3, float(sim.n_ghost), float(sim.n_entries), float(l1 - l0)], dtype=torch.float64, device=sim.dev)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    ke, pe = sim.h.mg_get_energies()
    e = torch.tensor([ke, pe], dtype=torch.float64, device=sim.dev)
    dist.all_reduce(e)
    if rank == 0:
        ms_max = float(tmax[0])
        npairs = float(tsum[3])  # half lists: cross-slab pairs are counted by both owners
        peak, peak_src = measured_peak_hbm()
        step_bytes = 440.0 * n + 16.0 * npairs
        out = {"metric": METRIC, "value": n * args.steps / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": w["desc"], "name": "c5" if args.workload == "c5" else "c4-weak", "n_atoms": n, "atoms_per_gpu": n // world,
                          "ghosts_per_gpu_max": int(tmax[2]),
                          "parallelism": (f"morton-slab x{world}, halo pulled from peer memory over NVLink by mg_pull_kernel (no collective)"
                                          if exchange == "peer" else f"morton-slab x{world}, all_gather of float4 positions per step (NCCL)"),
                          "exchange": exchange, "list_entries_sum_over_ranks": int(npairs),
                          "l2_policy": "per-GPU working set exceeds the 126 MB L2"},
               "roofline": {"bound": "hbm", "kernel": "traverse_kernel", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                            "achieved": round(step_bytes / world / (ms_max / args.steps * 1e-3) / 1e9, 1),
                            "frac": round(step_bytes / world / (ms_max / args.steps * 1e-3) / 1e9 / peak, 4), "traffic": None,
                            "note": "whole-step algorithmic bytes (440 N + 16 P) per GPU over the step time; rank 0 stage split below",
                            "stage_ms_per_step_rank0": {s: round(stages[s][0] / split_steps, 4) for s in stages if stages[s][1] > 0},
                            "stage_split_note": f"separate run of {split_steps} steps with all stages bracketed by events"},
               "e2e": {"value": n * args.steps / (float(tmax[1]) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                       "note": "wall clock of the same loop incl. the host driver; state is device resident across steps in the "
                               "multi-GPU driver (peer exchange: no host round trip inside the loop)"},
               "gpu_launches": int(tsum[4]), "clocks": clk.summary(), "energy": {"ke": float(e[0]), "pe": float(e[1])}}
        emit(json.dumps(out))
    sim.close()
    dist.barrier()
    dist.destroy_process_group()
