"""Multi-GPU driver: Morton-slab partition + per-step halo exchange.

One process per GPU.  The reference has no multi-process code at all; this is the B200-native
extension of its hot path:

  * the atoms are ordered by a 30-bit Morton key and cut into `world` contiguous key ranges of (nearly) equal atom
    count, the cuts snapped to octree boundaries; rank g owns the atoms whose key lies in range g.  With
    `migrate_every = k` ownership follows the atoms: every k-th step the atoms whose key left the range are handed to
    their new owner through outboxes in the published region (nb200_mg_set_migration); with 0 atoms never change rank;
  * every step each rank kick-drifts its owned atoms and PUBLISHES their positions and the boxes of its
    32-atom publication leaves in a peer-mapped buffer, then every rank pulls as ghosts the foreign atoms
    within the cutoff of its slab's bounding box (and occupancy grid, for ragged slabs).  The owned atoms stay resident
    in curve order with their own LBVH; the ghosts form a second sorted segment with a tree of their own, and the owned
    leaves query both (owned pass + ghost pass);
  * exchange="peer" (default): the pull is the library's own kernel reading the peers' GPU memory over
    NVLink/NVSwitch (CUDA IPC mappings), synchronised only by per-rank step flags — no collective, no
    barrier, ~1.5 MB per rank per step at 1M atoms per GPU;
    exchange="nccl": the packed float4 positions of all ranks are all-gathered with torch.distributed
    (16 B x N per rank per step) and the ghosts are selected from the gathered array;
  * the neighbour list is a half list (every pair with at least one owned atom, once; a pair of two ghosts
    belongs to other ranks and is never generated): reactions go to owned partners only, a ghost's force is
    its owner's business, so every rank has complete forces for its own atoms and there is no reverse force
    reduction.  The union of the ranks' entries is exactly the single-GPU pair set
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


def morton_slab_partition(pos, world, box_min=(0.0, 0.0, 0.0), box_max=(1.0, 1.0, 1.0), equal=True, with_splitters=False, snap=0.01):
    """Returns `owner_order`: atom ids in global Morton order, and `bounds`: world+1 cut positions.
    Rank g owns owner_order[bounds[g]:bounds[g+1]].  equal=True (the NCCL exchange: fixed-size all_gather) requires
    len(pos) to be divisible by world; the peer exchange has no such constraint (equal=False).
    with_splitters: also the world+1 Morton keys that bound the ranks' key ranges (nb200_mg_set_migration).
    equal=False snaps every cut to the coarsest octree boundary (a key that is a multiple of 2^b, b as large as possible)
    that moves it by at most `snap` of a slab's atoms: a key range that ends a little past such a boundary owns a sliver of
    a far-away octree cell, its bounding box then spans both, and with migration every atom that wanders into the sliver
    drags the box (and the ghost set) across the domain.  Snapped ranges are unions of few aligned cells: compact boxes."""
    n = len(pos)
    if equal and n % world:
        raise ValueError(f"atom count {n} must be divisible by the number of ranks {world}")
    keys = morton30(pos, box_min, box_max)
    # stable argsort of the 30-bit keys as two 15-bit LSD passes: numpy sorts 16-bit integers with a radix sort, which is
    # twice as fast as its merge sort of the 32-bit keys at 64M atoms (identical permutation)
    o1 = np.argsort((keys & np.uint32(0x7FFF)).astype(np.uint16), kind="stable")
    order = o1[np.argsort((keys >> np.uint32(15)).astype(np.uint16)[o1], kind="stable")]
    del o1
    skeys = keys[order].astype(np.int64)
    split = np.empty(world + 1, np.uint32)
    split[0] = 0
    split[world] = 1 << 30
    if equal:
        bounds = np.arange(world + 1, dtype=np.int64) * (n // world)
        for g in range(1, world):
            split[g] = skeys[bounds[g]]  # atoms that share this key with the end of slab g-1 move up at the first migration
    else:
        bounds = np.round(np.linspace(0, n, world + 1)).astype(np.int64)
        tol = int(snap * n / world)
        for g in range(1, world):
            b = int(bounds[g])
            cut = int(skeys[min(b, n - 1)])
            for bit in range(29, -1, -1):
                c = ((cut + (1 << bit >> 1)) >> bit) << bit
                p = int(np.searchsorted(skeys, c, side="left"))
                if abs(p - b) <= tol and 0 < c < (1 << 30):
                    cut, b = c, p
                    break
            else:  # (bit 0 always fits unless c left the key space)
                b = int(np.searchsorted(skeys, cut, side="left"))
            bounds[g], split[g] = b, cut
        for g in range(1, world + 1):  # degenerate data: keep the cuts ordered
            bounds[g] = max(bounds[g], bounds[g - 1])
    if not with_splitters:
        return order, bounds
    return order, bounds, split


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
                 list_mode=1, migrate_every=0):
        import torch

        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.exchange = exchange
        w = workload
        pos = w["pos"]
        n = len(pos)
        self.n_total = n
        order, bounds, split = morton_slab_partition(pos, world, equal=(exchange == "nccl"), with_splitters=True)
        self.split, self.migrate_every = split, (migrate_every if exchange == "peer" and world > 1 else 0)
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
            if self.migrate_every:
                self.h.mg_set_migration(split, self.migrate_every)
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
        self.h.mg_step_async(nsteps, self.dt)  # (two-step CUDA graph in steady state)

    def sync(self):
        self.n_ghost, self.n_entries = self.h.mg_sync()

    def entries_global(self):
        """(a, b, d) with ORIGINAL atom ids (0-based) for this rank's list entries."""
        a, b, d = self.h.mg_get_entries(self.n_entries)
        return self.order[a], self.order[b], d

    def owned_original_ids(self):
        """ORIGINAL atom id of every row of mg_get_owned (the owned set changes when atoms migrate)."""
        return self.order[self.h.mg_get_owned_ids()]

    def close(self):
        self.h.close()


class VirtualCluster:
    """All `world` slabs on ONE device in one process: exercises exactly the code path of the real multi-GPU run
    (publication, flags, pull kernel or gathered-array ghost selection, half/directed list, force scatter), so
    the parity test does not need several GPUs.  The slabs are stepped in lockstep — all publish, then all pull —
    because a pull kernel waits for flags that only the other slabs' (same-device) kernels can set."""

    def __init__(self, pkg, workload, world, device=0, exchange="peer", list_mode=1, headroom=1.6, migrate_every=0):
        import torch
        self.torch = torch
        self.exchange = exchange
        self.sims = [SlabSimulation(pkg, workload, g, world, device, dist=None, defer=True, exchange=exchange, list_mode=list_mode,
                                    headroom=headroom, migrate_every=migrate_every) for g in range(world)]
        self.n_total = self.sims[0].n_total
        if exchange == "peer":
            bases = [s.h.mg_publication()[0] for s in self.sims]
            bounds = self.sims[0].bounds
            for g, s in enumerate(self.sims):
                s.h.mg_connect(world, g, bounds[:-1], np.diff(bounds), direct_base=bases)
                if s.migrate_every:
                    s.h.mg_set_migration(s.split, s.migrate_every)
        else:
            self.all_pos = torch.empty((self.n_total, 4), dtype=torch.float32, device=self.sims[0].dev)
        self._exchange()

    def _exchange(self):
        if self.exchange == "peer":
            for s in self.sims:
                s.h.mg_republish()  # (no-op unless a migration or a host-buffer step left the publication stale)
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
        out = np.full((self.n_total, 3), np.nan, np.float32)
        for s in self.sims:
            out[s.owned_original_ids()] = s.h.mg_get_owned(mode)
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


class LineGuard:
    """The ONE JSON line of a multi-rank bench run, printed exactly once by rank 0: either by finish(), with every variant
    that completed, or — when the variants have not finished within `budget_s` — by a watchdog thread, with the variants
    finished so far and an error note, after which every rank leaves the process (`leave`, default os._exit(0)): the main
    thread may be stuck inside a collective or a CUDA call and cannot be interrupted.  The headline is never lost to a
    hanging variant."""

    def __init__(self, rank, out, variants, emit, budget_s, leave=None):
        import threading
        self.rank, self.out, self.variants, self.emit, self.budget = rank, out, variants, emit, budget_s
        self.leave = leave or (lambda: os._exit(0))
        self._lock, self._emitted = threading.Lock(), False
        self._dog = threading.Timer(budget_s if rank == 0 else budget_s + 10.0, self._give_up)  # rank 0 prints first
        self._dog.daemon = True
        self._dog.start()

    def _emit_once(self, note=None):
        with self._lock:
            if self._emitted:
                return
            self._emitted = True
            if self.rank == 0:
                v = dict(self.variants)
                if note:
                    v["error"] = note
                self.out["variants"] = v
                self.emit(json.dumps(self.out))

    def _give_up(self):
        self._emit_once(f"variants stopped after {self.budget:.0f} s (NB200_VARIANT_BUDGET_S)")
        self.leave()

    def finish(self):
        """variants done: print the line (unless the watchdog already has) and arm a last timer for the closing barrier —
        a rank whose peers are gone still leaves"""
        import threading
        self._dog.cancel()
        self._emit_once()
        bye = threading.Timer(60.0, self.leave)
        bye.daemon = True
        bye.start()
        return bye


# ----------------------------------------------------------------------------------------------------
# bench entry (called by bench.py under torchrun)
# ----------------------------------------------------------------------------------------------------
def parity_digest(sim, w, dist, graft):
    """In-run MULTI-PROCESS parity proof (outside any timed region): every rank exports its list entries; a pair is
    counted by the rank that owns its lower-numbered atom (a cross-slab pair sits in both owners' lists), the per-rank
    digests (count, xor and sum of a 64-bit hash of (min id, max id, bits(d))) are combined, and rank 0 compares them with
    the oracle's independent O(N) cell-grid search over the gathered positions of the same step."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    O = graft.load_oracle()
    sim.search(None)  # synchronous search at the current positions: exact ghost segment, list readable
    a, b, d = sim.entries_global()
    mine = sim.owned_original_ids()            # (ownership follows the atoms when migration is on)
    owned = np.zeros(sim.n_total, bool)
    owned[mine] = True
    lo = np.minimum(a, b)
    # a pair with ONE owned atom sits in two ranks' lists: the owner of the lower id counts it; with BOTH owned it is only in mine
    keep = owned[lo]
    dg = O.digest_pairs(a[keep] + 1, b[keep] + 1, d[keep])
    t = torch.tensor([dg["count"], dg["xor"] & 0x7fffffffffffffff, dg["xor"] >> 63, dg["sum"] & 0x7fffffffffffffff, dg["sum"] >> 63],
                     dtype=torch.int64, device=sim.dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, sim.h.mg_get_owned(0)))
    out = None
    if rank == 0:
        x = np.empty((sim.n_total, 3), np.float32)
        for ids_g, pos_g in gathered:
            x[ids_g] = pos_g
        t0 = time.perf_counter()
        ref = O.cellgrid_digest(x, w["cutoff"])
        cnt, xo, su = 0, 0, 0
        for tt in allt:
            v = [int(z) for z in tt.cpu().tolist()]
            cnt += v[0]
            xo ^= v[1] | (v[2] << 63)
            su = (su + (v[3] | (v[4] << 63))) & 0xffffffffffffffff
        out = {"count": cnt, "count_match": cnt == ref["count"], "xor_match": xo == ref["xor"], "sum_match": su == ref["sum"],
               "oracle_pairs": ref["count"], "oracle_s": round(time.perf_counter() - t0, 2),
               "note": "union of the ranks' list entries (each pair counted by the owner of its lower atom id) vs the oracle's "
                       "cell-grid search on the gathered positions, same step; digests of (min id, max id, bits(d))"}
    return out


def bench_multi(args, make_workload, METRIC, UNIT, ClockSampler, measured_peak_hbm, emit=print):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as graft
    pkg = graft.load_package()
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    exchange = os.environ.get("NB200_EXCHANGE", "peer")
    migrate_every = int(os.environ.get("NB200_MIGRATE_EVERY", "20"))  # ownership follows the atoms (0: atoms never change rank)
    dev = torch.device("cuda", local_rank)

    def red(vals, op):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return [float(v) for v in t.tolist()]

    live = []

    def run_case(*a, **kw):
        try:
            return run_case_(*a, **kw)
        finally:
            while live:
                try:
                    live.pop().close()
                except Exception:
                    pass

    def run_case_(w, steps, warm, melt, parity=False, e2e_steps=0, split=True, clk=None, headroom=1.6):
        """One timed slab run of workload w: K steps between CUDA events on the library's stream, max over ranks."""
        n = w["n"]
        def together(fn, what):
            # a failure on one rank (capacity, a peer that stopped publishing) must take every rank out of the case together:
            # a rank that skipped ahead to the next collective alone would leave the others waiting in theirs for ever
            err, out = None, None
            try:
                out = fn()
            except Exception as exc:
                err = exc
            if red([1.0 if err is not None else 0.0], dist.ReduceOp.MAX)[0] > 0:
                raise RuntimeError(str(err) if err is not None else f"another rank failed in {what}")
            return out
        sim = SlabSimulation(pkg, w, rank, world, local_rank, dist, exchange=exchange, migrate_every=migrate_every, headroom=headroom, defer=True)
        live.append(sim)  # (a case that fails must not leave its handle — and its share of HBM — to the next case)
        together(sim.exchange_and_search, "the first search")
        def run(k):
            def go():
                if exchange == "peer":
                    sim.step_async(k)
                    sim.sync()
                else:
                    sim.step(k)
            together(go, "a step phase")
        run(warm)
        g_first = red([float(sim.n_ghost)], dist.ReduceOp.MAX)[0]
        if melt:
            run(melt)
        l0 = sim.h.get_stats()["kernel_launches"]
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if exchange == "peer":
            def timed():
                sim.h.timer_start()          # CUDA events on the library's own stream
                sim.step_async(steps)
                t = sim.h.timer_stop()
                sim.sync()
                return t
            ms = together(timed, "the timed region")
        else:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(torch.cuda.current_stream())
            sim.step(steps)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dist.barrier()
        l1 = sim.h.get_stats()["kernel_launches"]
        ms_max, wall_max, g_last = red([ms, wall * 1e3, float(sim.n_ghost)], dist.ReduceOp.MAX)
        ent_sum, launches = red([float(sim.n_entries), float(l1 - l0)], dist.ReduceOp.SUM)
        res = {"n_atoms": n, "atoms_per_gpu": n // world, "steps": steps, "ms_per_step": ms_max / steps, "value": n * steps / (ms_max * 1e-3),
               "ghosts_per_gpu_max": {"after_warmup": int(g_first), "at_end": int(g_last), "steps_between": melt + steps},
               "migrate_every": sim.migrate_every,
               "list_entries_sum_over_ranks": int(ent_sum), "gpu_launches": int(launches), "wall_ms_per_step": wall_max / steps}
        st = sim.h.get_stats()
        if split:  # stage split of rank 0: a short extra run with every stage bracketed by events on the main stream
            sim.h.set_profiling(True)
            run(20)
            stages = sim.h.get_stage_times()
            sim.h.set_profiling(False)
            res["stage_ms_per_step_rank0"] = {s_: round(stages[s_][0] / 20, 4) for s_ in stages if stages[s_][1] > 0}
            res["rank0_list"] = {"tile_words": st["n_slots"], "groups": st["n_segments"], "entries": st["n_entries"], "n_own": sim.n_own}
        if parity:
            try:
                res["parity"] = parity_digest(sim, w, dist, graft)
            except Exception as exc:
                res["parity"] = {"error": str(exc)[:300]}
        ke, pe = sim.h.mg_get_energies()
        e = red([ke, pe], dist.ReduceOp.SUM)
        res["energy"] = {"ke": e[0], "pe": e[1]}
        sim.close()
        dist.barrier()
        if e2e_steps:
            # HOST-buffer form of the slab step: every rank uploads its owned x(t) from pinned memory and downloads x(t+dt), every step.
            # (Own short run without migration: the caller's array keeps the hand-over order of its rows.)
            sim = SlabSimulation(pkg, w, rank, world, local_rank, dist, exchange=exchange, migrate_every=0)
            live.append(sim)
            sim.step_async(warm)
            sim.sync()
            xh = torch.from_numpy(sim.h.mg_get_owned(0)).pin_memory()
            def e2e(k):
                for _ in range(k):
                    sim.h.mg_leapfrog_host_async(xh.data_ptr(), 3, sim.dt)
                    sim.h.mg_sync()
            e2e(3)
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e(e2e_steps)
            torch.cuda.synchronize()
            wall_e = time.perf_counter() - t0
            dist.barrier()
            wall_e = red([wall_e], dist.ReduceOp.MAX)[0]
            res["e2e"] = {"value": n * e2e_steps / wall_e, "unit": UNIT, "h2d_bytes_per_step": 12 * n, "d2h_bytes_per_step": 12 * n,
                          "steps": e2e_steps, "ms_per_step": wall_e / e2e_steps * 1e3,
                          "api": "nb200_mg_leapfrog_host_async + nb200_mg_sync per step on every rank (C ABI, pinned host buffers, positions-only)",
                          "timer": "host wall clock, max over ranks, barrier on both sides",
                          "note": "every rank uploads its owned x(t), publishes it, pulls the halo over NVLink, rebuilds list and forces, "
                                  "integrates and downloads x(t+dt); bytes are the sum over ranks; separate short run without migration"}
            sim.close()
            dist.barrier()
        return res

    # ---- headline: weak scaling, ~1M atoms per GPU (BASELINE config 4 = 8M atoms on 8 GPUs), melted before timing ----
    if args.workload == "c5":  # BASELINE config 5: 64M-atom clustered gas on 8 GPUs (8M per GPU), traversal-imbalance stress test
        per_gpu = args.n or 8_000_000
        w = make_workload("c5", per_gpu * world)
        melt = 0
    else:
        per_gpu = args.n or 1_000_000
        m = int(round((per_gpu * world) ** (1 / 3)))
        while (m ** 3) % world:
            m += 1
        w = make_workload("c4", m ** 3)
        melt = args.melt
    n = w["n"]
    with ClockSampler(local_rank) as clk:
        head = run_case(w, args.steps, args.warmup, melt, parity=(n <= 9_000_000), e2e_steps=max(6, min(args.steps, 40)))
    out = None
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        ms_step = head["ms_per_step"]
        npairs = float(head["list_entries_sum_over_ranks"])  # half lists: cross-slab pairs are counted by both owners
        step_bytes = 440.0 * n + 16.0 * npairs
        stages = head.get("stage_ms_per_step_rank0", {})
        r0 = head.get("rank0_list", {})
        # dominant kernel on every rank: the owned pass of the traversal with fused forces; algorithmic bytes as for one GPU
        # (positions, tree, tile list written, forces), from rank 0's counts; its time from rank 0's stage split
        dom_ms = stages.get("traverse", 0.0)
        dom_bytes = 48.0 * r0.get("n_own", 0) + 96.0 * r0.get("n_own", 0) / 32 + 4.0 * r0.get("tile_words", 0) + 16.0 * r0.get("groups", 0)
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        out = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": w["desc"] + (f"; timed after {melt} steps of melting" if melt else ""),
                          "name": "c5" if args.workload == "c5" else "c4-weak", "n_atoms": n, "atoms_per_gpu": n // world,
                          "ghosts_per_gpu_max": head["ghosts_per_gpu_max"],
                          "parallelism": (f"morton-slab x{world}: owned atoms resident in curve order, halo pulled from peer memory over NVLink by "
                                          f"mg_pull_kernel (no collective), ghost tree + ghost pass on a second stream beside the owned pass"
                                          if exchange == "peer" else f"morton-slab x{world}, all_gather of float4 positions per step (NCCL)"),
                          "exchange": exchange, "list_entries_sum_over_ranks": int(npairs),
                          "l2_policy": "per-GPU working set exceeds the 126 MB L2"},
               "roofline": {"bound": "hbm", "kernel": "traverse_kernel<fused forces> (owned pass, rank 0)", "peak": peak, "unit": "GB/s",
                            "peak_source": peak_src, "achieved": round(achieved, 1), "frac": round(achieved / peak, 4), "traffic": None,
                            "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms_per_launch": dom_ms,
                            "kernel_share_of_step": round(dom_ms / ms_step, 3) if ms_step else None,
                            "limiter": "instruction issue, not HBM",
                            "whole_step_per_gpu": {"algorithmic_bytes": step_bytes / world, "achieved": round(step_bytes / world / (ms_step * 1e-3) / 1e9, 1),
                                                   "frac": round(step_bytes / world / (ms_step * 1e-3) / 1e9 / peak, 4)},
                            "stage_ms_per_step_rank0": stages,
                            "stage_split_note": "separate run of 20 steps with every main-stream stage bracketed by events (the ghost stream's "
                                                "work overlaps and is not in the split)"},
               "e2e": head.get("e2e"), "parity": head.get("parity"), "variants": {},
               "gpu_launches": head["gpu_launches"], "clocks": clk.summary(), "energy": head["energy"]}
    # ---- variants (after the headline is complete), guarded by a watchdog: a variant that hangs (a rank stuck in a
    # collective or waiting for a peer) must not cost the headline line
    variants = {}
    guard = LineGuard(rank, out, variants, emit, float(os.environ.get("NB200_VARIANT_BUDGET_S", "300")))
    if args.workload != "c5" and not args.n and not os.environ.get("NB200_NO_VARIANTS"):
        try:
            if world in (2, 4, 8):  # BASELINE config 4 as written: 8M atoms at 2 / 4 / 8 GPUs (strong scaling beside the weak headline)
                w8 = make_workload("c4", 200 ** 3)
                r = run_case(w8, max(20, min(args.steps, 100)), args.warmup, 100, parity=True, split=False)
                variants["c4_strong_8M"] = {k: r[k] for k in ("n_atoms", "atoms_per_gpu", "value", "ms_per_step", "steps", "ghosts_per_gpu_max", "parity")}
                variants["c4_strong_8M"]["scaling"] = "strong"
                del w8
        except Exception as exc:  # a variant must never cost the headline line
            variants["c4_strong_8M"] = {"error": str(exc)[:300]}
        try:
            if world == 8:          # BASELINE config 5: 64M-atom dilute/clustered gas, search + Coulomb force every step
                w5 = make_workload("c5", 64_000_000)
                r = run_case(w5, max(10, min(args.steps, 40)), args.warmup, 0, parity=False, split=False, headroom=2.0)
                variants["c5_64M"] = {k: r[k] for k in ("n_atoms", "atoms_per_gpu", "value", "ms_per_step", "steps", "ghosts_per_gpu_max")}
                del w5
        except Exception as exc:
            variants["c5_64M"] = {"error": str(exc)[:300]}
    guard.finish()
    dist.barrier()
    dist.destroy_process_group()
