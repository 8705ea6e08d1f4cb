"""Multi-GPU path: host-side partition logic on CPU (gloo, world_size 2) and the device path on one GPU
through VirtualCluster (all slabs on one device, all_gather replaced by a concatenation)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, uniform_positions


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n, cutoff, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    mg = __import__("importlib").import_module(graft.load_package().__name__ + ".multigpu")
    O = graft.load_oracle()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos = uniform_positions(n, 99)
    order, bounds = mg.morton_slab_partition(pos, world)
    mine = order[bounds[rank]:bounds[rank + 1]]
    send = torch.from_numpy(np.concatenate([pos[mine], np.zeros((len(mine), 1), np.float32)], 1))
    allp = torch.empty((n, 4), dtype=torch.float32)
    dist.all_gather_into_tensor(allp, send)  # the per-step exchange, on gloo
    allp = allp.numpy()
    own_begin, n_own = int(bounds[rank]), len(mine)
    ghosts = mg.select_ghosts_reference(allp, own_begin, n_own, cutoff)
    local = np.concatenate([np.arange(own_begin, own_begin + n_own), ghosts])
    a, b, d = O.brute_force(allp[local, :3].copy(), cutoff, "d2", nthreads=2)
    a, b = local[a - 1], local[b - 1]  # gathered indices
    # directed rows of OWNED atoms only
    own = lambda x: (x >= own_begin) & (x < own_begin + n_own)
    ra = np.concatenate([a[own(a)], b[own(b)]])
    rb = np.concatenate([b[own(a)], a[own(b)]])
    rows = np.stack([order[ra], order[rb]], 1)
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    if rank == 0:
        q.put((np.concatenate(gathered), len(ghosts)))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_partition_and_ghosts_gloo_world2(oracle):
    """world_size-2 gloo run of the host-side logic: Morton-slab partition, gathered positions, ghost
    selection; the union of the ranks' owned rows must be exactly the directed form of the global list."""
    import torch.multiprocessing as mp
    n, cutoff, world = 6000, 0.06, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, cutoff, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows, nghost = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pos = uniform_positions(n, 99)
    a, b, d = oracle.brute_force(pos, cutoff, "d2")
    ref = np.concatenate([np.stack([a - 1, b - 1], 1), np.stack([b - 1, a - 1], 1)])
    key = lambda r: np.sort(r[:, 0].astype(np.int64) * n + r[:, 1])
    assert len(rows) == len(ref) and np.array_equal(key(rows), key(ref))
    assert 0 < nghost < n // 2


def test_partition_properties(pkg):
    mg = __import__("importlib").import_module(pkg.__name__ + ".multigpu")
    pos = uniform_positions(4096, 5)
    order, bounds = mg.morton_slab_partition(pos, 8)
    assert sorted(order.tolist()) == list(range(4096)) and bounds.tolist() == [512 * i for i in range(9)]
    keys = mg.morton30(pos)[order]
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)
    with pytest.raises(ValueError):
        mg.morton_slab_partition(pos[:4095], 8)
    order, bounds = mg.morton_slab_partition(pos[:4095], 8, equal=False, snap=0.0)  # peer exchange: counts may differ by one
    assert bounds[0] == 0 and bounds[-1] == 4095 and set(np.diff(bounds)) <= {511, 512} and len(np.unique(order)) == 4095
    # default: every cut snapped to the coarsest octree boundary within 1 % of a slab (compact slabs, see the docstring)
    order, bounds, split = mg.morton_slab_partition(pos[:4095], 8, equal=False, with_splitters=True)
    assert bounds[0] == 0 and bounds[-1] == 4095 and np.all(np.abs(np.diff(bounds) - 4095 / 8) <= 0.02 * 512 + 1)
    keys = mg.morton30(pos[:4095])[order]
    for g in range(8):
        assert np.all(keys[bounds[g]:bounds[g + 1]] >= split[g]) and np.all(keys[bounds[g]:bounds[g + 1]].astype(np.int64) < int(split[g + 1]) + (g == 7))
    assert all(int(split[g]) % (1 << 20) == 0 for g in range(1, 8))  # uniform points: the cuts sit on coarse cell boundaries


def test_partition_order_is_the_stable_sort_of_the_keys(pkg):
    """the two-pass 16-bit argsort inside morton_slab_partition is the stable argsort of the 30-bit keys, duplicates included"""
    mg = __import__("importlib").import_module(pkg.__name__ + ".multigpu")
    rng = np.random.default_rng(11)
    pos = np.concatenate([uniform_positions(20000, 6), np.repeat(rng.random((50, 3)).astype(np.float32), 40, 0),
                          np.array([[0, 0, 0], [1 - 2 ** -24] * 3, [0.5, 0.5, 0.5]], np.float32)])
    order, _ = mg.morton_slab_partition(pos, 3, equal=False)
    assert np.array_equal(order, np.argsort(mg.morton30(pos), kind="stable"))


def test_line_guard_prints_the_bench_line_exactly_once(pkg):
    """bench_multi's LineGuard: the normal path prints the line with all variants; a hanging variant makes the watchdog print
    the headline with what is there plus an error note, and every rank leaves; never two lines, and only rank 0 prints."""
    import json
    import time
    mg = __import__("importlib").import_module(pkg.__name__ + ".multigpu")
    lines, left = [], []
    variants = {}
    g = mg.LineGuard(0, {"value": 1.0, "variants": {}}, variants, lines.append, 30.0, leave=lambda: left.append("main"))
    variants["c4_strong_8M"] = {"value": 2.0}
    bye = g.finish()
    bye.cancel()
    g._give_up()  # (a late watchdog must not print again)
    assert len(lines) == 1 and json.loads(lines[0]) == {"value": 1.0, "variants": {"c4_strong_8M": {"value": 2.0}}}
    # the watchdog path: rank 0 prints after the budget, with the finished variants and the note
    lines, left, variants = [], [], {"a": 1}
    g = mg.LineGuard(0, {"value": 3.0, "variants": {}}, variants, lines.append, 0.2, leave=lambda: left.append("dog"))
    t0 = time.time()
    while not left and time.time() - t0 < 10:
        time.sleep(0.02)
    assert left == ["dog"] and len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] == 3.0 and d["variants"]["a"] == 1 and "stopped after" in d["variants"]["error"]
    g.finish().cancel()
    assert len(lines) == 1
    # other ranks never print
    lines = []
    g = mg.LineGuard(1, None, {}, lines.append, 30.0, leave=lambda: None)
    g.finish().cancel()
    assert lines == []


@pytest.mark.gpu
def test_morton30_host_twin_matches_device(pkg, big_handle):
    mg = __import__("importlib").import_module(pkg.__name__ + ".multigpu")
    x = uniform_positions(50_000, 17)
    assert np.array_equal(mg.morton30(x), big_handle.morton30(x))


def _workload(n_side, seed=3):
    sys.path.insert(0, ROOT)
    from bench import make_workload
    return make_workload("c4", n_side ** 3, seed=seed)


@pytest.mark.gpu
@pytest.mark.parametrize("exchange,list_mode", [("peer", 1), ("peer", 0), ("nccl", 1), ("nccl", 0)])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_virtual_cluster_pair_set_and_forces(pkg, oracle, world, exchange, list_mode):
    """Half list: every rank holds each pair with >= 1 owned atom exactly once and no ghost-ghost pair; the union
    over ranks is the exact pair set.  Directed list: union over slabs of the owned rows == directed form of the
    exact pair set.  d bit-exact.  Forces of the slab run == fp64 oracle to Float32 summation accuracy.
    exchange: ghosts pulled from the peers' published memory by mg_pull_kernel, or selected from a gathered array."""
    mg = __import__("importlib").import_module(pkg.__name__ + ".multigpu")
    w = _workload(24)  # 13824 atoms
    n = w["n"]
    vc = mg.VirtualCluster(pkg, w, world, exchange=exchange, list_mode=list_mode)
    ra, rb, rd = oracle.brute_force(w["pos"], w["cutoff"], "d2")
    ra, rb = (ra - 1).astype(np.int64), (rb - 1).astype(np.int64)
    owner = np.empty(n, np.int64)
    for g, s in enumerate(vc.sims):
        owner[s.owned_ids] = g
    parts = vc.entries()
    if list_mode == 1:
        ref_key = np.minimum(ra, rb) * n + np.maximum(ra, rb)
        o = np.argsort(ref_key)
        ref_key, ref_d = ref_key[o], rd[o]
        for g, (a, b, d) in enumerate(parts):
            a, b = a.astype(np.int64), b.astype(np.int64)
            assert np.all((owner[a] == g) | (owner[b] == g)), "a ghost-ghost pair was emitted"
            key = np.minimum(a, b) * n + np.maximum(a, b)
            o = np.argsort(key)
            key, d = key[o], d[o]
            assert len(np.unique(key)) == len(key), "a pair was emitted twice on one rank"
            want = (owner[ra] == g) | (owner[rb] == g)
            wk = np.minimum(ra[want], rb[want]) * n + np.maximum(ra[want], rb[want])
            o2 = np.argsort(wk)
            assert np.array_equal(key, wk[o2]), "rank %d: pairs touching its owned atoms differ from the oracle" % g
            assert np.array_equal(d.view(np.uint32), rd[want][o2].view(np.uint32))
        allk = np.unique(np.concatenate([np.minimum(a.astype(np.int64), b) * n + np.maximum(a.astype(np.int64), b) for a, b, _ in parts]))
        assert np.array_equal(allk, ref_key)
    else:
        a = np.concatenate([p[0] for p in parts]).astype(np.int64)
        b = np.concatenate([p[1] for p in parts]).astype(np.int64)
        d = np.concatenate([p[2] for p in parts])
        assert len(a) == 2 * len(ra)
        key = a * n + b
        rkey = np.concatenate([ra * n + rb, rb * n + ra])
        assert np.array_equal(np.sort(key), np.sort(rkey))
        dd = d[np.argsort(key, kind="stable")]
        rdd = np.concatenate([rd, rd])[np.argsort(rkey, kind="stable")]
        assert np.array_equal(dd.view(np.uint32), rdd.view(np.uint32))
    f = vc.gather(2)
    f64, pe64, scale = oracle.forces_physical_f64(w["pos"], w["charge"], (ra + 1).astype(np.int32), (rb + 1).astype(np.int32), w["eps"],
                                                  w["sigma"], w["kcoul"], w["cutoff"], True)
    assert (np.abs(f - f64).max(axis=1) / scale).max() < 1e-5
    ke, pe = vc.energies()
    assert abs(pe - pe64.sum()) < 1e-5 * np.abs(pe64).sum()
    vc.close()


@pytest.mark.gpu
def test_virtual_cluster_trajectory_matches_single_gpu(pkg):
    mg = __import__("importlib").import_module(pkg.__name__ + ".multigpu")
    w = _workload(20)
    h = pkg.Handle(w["n"])
    h.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
    h.set_system(w["pos"], w["vel"], w["mass"], w["charge"])
    h.step(10, w["dt"])
    p1, v1 = h.get_positions(), h.get_velocities()
    ke1, pe1 = h.get_energies()
    vc = mg.VirtualCluster(pkg, w, 4)  # peer exchange, half list
    vc.step(10)
    p4, v4 = vc.gather(0), vc.gather(1)
    ke4, pe4 = vc.energies()
    # same physics; force sums differ only in Float32 summation order
    assert np.abs(p4 - p1).max() < 1e-5 * w["sigma"]
    assert np.abs(v4 - v1).max() < 1e-4 * np.abs(v1).max()
    assert abs(ke4 - ke1) < 1e-5 * abs(ke1) and abs(pe4 - pe1) < 1e-5 * abs(pe1)
    h.close()
    vc.close()


@pytest.mark.gpu
def test_virtual_cluster_async_steps_match_the_synchronous_ones(pkg):
    """nb200_mg_search_force_async: launches sized for n_own + ghost capacity, unused ghost slots hold NaN placeholders,
    no host round trip.  Same trajectory as the synchronous step; nb200_mg_sync reports the real ghost count."""
    mg = __import__("importlib").import_module(pkg.__name__ + ".multigpu")
    w = _workload(20)
    a = mg.VirtualCluster(pkg, w, 4)
    b = mg.VirtualCluster(pkg, w, 4)
    a.step(12)
    b.step_async(12)
    assert np.abs(a.gather(0) - b.gather(0)).max() < 1e-5 * w["sigma"]
    assert np.abs(a.gather(1) - b.gather(1)).max() < 1e-4 * np.abs(a.gather(1)).max()
    ea, eb = a.energies(), b.energies()
    assert abs(ea[0] - eb[0]) < 1e-5 * abs(ea[0]) and abs(ea[1] - eb[1]) < 1e-5 * abs(ea[1])
    for sa, sb in zip(a.sims, b.sims):
        assert sb.n_ghost == sa.n_ghost and sb.n_entries == sa.n_entries  # NaN placeholders contribute nothing
    # a synchronous search after asynchronous steps still works (exact n again)
    b.step(1)
    a.step(1)
    assert np.abs(a.gather(0) - b.gather(0)).max() < 1e-5 * w["sigma"]
    a.close(); b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,world", [("uniform", 5), ("uniform", 3), ("uniform", 7), ("c5", 8)])
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_virtual_cluster_with_ragged_slabs(pkg, oracle, exchange, kind, world):
    """Morton slabs of equal atom count are boxes only for 8^k ranks on uniform data.  With 3 or 5 ranks, or on the
    clustered gas of BASELINE config 5 (scaled down), a slab's AABB spans most of the domain, so the ghost selection
    also uses the slab's 64^3 occupancy grid: pair set per rank exact, and far fewer ghosts than 'every foreign atom
    inside the dilated slab AABB'."""
    mg = __import__("importlib").import_module(pkg.__name__ + ".multigpu")
    sys.path.insert(0, ROOT)
    from bench import make_workload
    if kind == "c5":
        w = make_workload("c5", 64_000)
    else:
        w = make_workload("random", 60_000)
        w["cutoff"] = 0.012
        w["eps"], w["kcoul"], w["charge"] = 0.0, 1e-6, np.random.default_rng(1).uniform(-1, 1, w["n"]).astype(np.float32)
    n = w["n"]
    if exchange == "nccl" and n % world:
        pytest.skip("the all-gather exchange needs equal slab sizes; the peer exchange does not")
    vc = mg.VirtualCluster(pkg, w, world, exchange=exchange, headroom=3.5)
    ra, rb, rd = oracle.brute_force(w["pos"], np.float32(w["cutoff"]), "d2")
    ra, rb = (ra - 1).astype(np.int64), (rb - 1).astype(np.int64)
    owner = np.empty(n, np.int64)
    for g, s in enumerate(vc.sims):
        owner[s.owned_ids] = g
    ghosts_aabb = 0
    for g, (a, b, d) in enumerate(vc.entries()):
        a, b = a.astype(np.int64), b.astype(np.int64)
        key = np.sort(np.minimum(a, b) * n + np.maximum(a, b))
        want = (owner[ra] == g) | (owner[rb] == g)
        wk = np.sort(np.minimum(ra[want], rb[want]) * n + np.maximum(ra[want], rb[want]))
        assert np.array_equal(key, wk), g
        own = w["pos"][vc.sims[g].owned_ids]
        lo, hi = own.min(0) - w["cutoff"], own.max(0) + w["cutoff"]
        inside = np.all((w["pos"] >= lo) & (w["pos"] <= hi), axis=1)
        ghosts_aabb += int(inside.sum()) - len(own)
    ghosts = sum(s.n_ghost for s in vc.sims)
    if kind == "uniform":
        assert ghosts < 0.6 * ghosts_aabb, (ghosts, ghosts_aabb)
    else:
        assert ghosts <= ghosts_aabb
    vc.close()


@pytest.mark.parametrize("kind,world,cutoff", [("uniform", 5, 0.02), ("uniform", 8, 0.05), ("clustered", 3, 0.004), ("clustered", 8, 0.03),
                                               ("outside", 4, 0.03)])
def test_occupancy_grid_never_drops_a_needed_ghost(pkg, kind, world, cutoff):
    """The slab's 64^3 occupancy grid (numpy twin of the device kernels) is conservative: every foreign atom within the
    cutoff of ANY owned atom falls into a set cell — also for ragged slabs, clustered data and atoms outside the box
    (cell indices are clamped) — while it prunes most of a ragged slab's AABB."""
    from scipy.spatial import cKDTree
    mg = __import__("importlib").import_module(pkg.__name__ + ".multigpu")
    rng = np.random.default_rng(world * 1000 + int(cutoff * 1e4))
    n = 24_000 - 24_000 % world
    if kind == "uniform":
        pos = rng.random((n, 3)).astype(np.float32)
    elif kind == "clustered":
        centres = rng.random((40, 3))
        pos = np.concatenate([rng.random((n // 2, 3)), centres[rng.integers(0, 40, n - n // 2)] + 0.01 * rng.standard_normal((n - n // 2, 3))])
        pos = np.clip(pos, 0, 0.999999).astype(np.float32)
    else:
        pos = (rng.random((n, 3)) * 1.3 - 0.15).astype(np.float32)  # 15 % of the atoms sit outside the box on every side
    order, bounds = mg.morton_slab_partition(pos, world)
    pruned = 0
    for g in range(world):
        own = pos[order[bounds[g]:bounds[g + 1]]]
        foreign = np.delete(pos[order], np.arange(bounds[g], bounds[g + 1]), axis=0)
        grid = mg.occupancy_grid_reference(own, cutoff)
        inside = mg.grid_lookup_reference(grid, foreign)
        d, _ = cKDTree(own.astype(np.float64)).query(foreign.astype(np.float64), k=1)
        needed = d <= cutoff * (1 + 1e-6)
        assert not np.any(needed & ~inside), (g, int((needed & ~inside).sum()))
        lo, hi = own.min(0) - cutoff, own.max(0) + cutoff
        in_aabb = np.all((foreign >= lo) & (foreign <= hi), axis=1)
        pruned += int((in_aabb & ~inside).sum())
    if kind != "uniform" or world != 8:
        assert pruned > 0  # ragged slabs: the grid removes atoms the AABB test alone would keep


@pytest.mark.gpu
@pytest.mark.parametrize("world,async_steps", [(2, False), (4, True), (8, True)])
def test_virtual_cluster_migration_keeps_the_pair_set_and_the_trajectory(pkg, oracle, world, async_steps):
    """nb200_mg_set_migration: ownership follows the atoms.  A hot liquid is run with a migration every 5th step: atoms must
    be conserved (every global id owned by exactly one rank), owners must match the key ranges right after a migration, the
    union of the ranks' lists must be the exact pair set of the current positions (d bit-exact, no ghost-ghost pair, no
    duplicate on a rank), and the trajectory and energies must be those of the single-GPU run."""
    mg = __import__("importlib").import_module(pkg.__name__ + ".multigpu")
    w = _workload(20)  # 8000 atoms
    w = dict(w)
    w["vel"] = (w["vel"] * 4.0).astype(np.float32)  # hot: many atoms cross the slab borders within a few steps
    n = w["n"]
    h = pkg.Handle(n)
    h.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
    h.set_system(w["pos"], w["vel"], w["mass"], w["charge"])
    vc = mg.VirtualCluster(pkg, w, world, migrate_every=5, headroom=3.5)
    first_owner = np.empty(n, np.int64)
    for g, s in enumerate(vc.sims):
        first_owner[s.owned_ids] = g
    steps = 40
    if async_steps:
        vc.step_async(steps)
    else:
        vc.step(steps)
    h.step(steps, w["dt"])
    # conservation and ownership
    owner = np.full(n, -1, np.int64)
    counts = []
    for g, s in enumerate(vc.sims):
        ids = s.owned_original_ids()
        assert len(np.unique(ids)) == len(ids)
        assert np.all(owner[ids] == -1), "an atom is owned by two ranks"
        owner[ids] = g
        counts.append(len(ids))
    assert np.all(owner >= 0) and sum(counts) == n, "atoms were lost"
    assert np.count_nonzero(owner != first_owner) > 20, "the test did not exercise migration"
    x = vc.gather(0)
    assert not np.isnan(x).any()
    keys = mg.morton30(x)
    split = vc.sims[0].split
    # 40 steps with a migration every 5th: the last step was a migration step, so owners match the key ranges exactly
    want_owner = np.searchsorted(split[1:-1].astype(np.int64), keys.astype(np.int64), side="right")
    assert np.array_equal(owner, want_owner)
    # same physics as one GPU
    p1, v1 = h.get_positions(), h.get_velocities()
    assert np.abs(x - p1).max() < 2e-5 * w["sigma"] * 10
    assert np.abs(vc.gather(1) - v1).max() < 1e-3 * np.abs(v1).max()
    ke1, pe1 = h.get_energies()
    ke, pe = vc.energies()
    assert abs(ke - ke1) < 1e-4 * abs(ke1) and abs(pe - pe1) < 1e-4 * abs(pe1)
    # exact pair set at the current positions, rank by rank
    vc._exchange()
    ra, rb, rd = oracle.brute_force(x, w["cutoff"], "d2")
    ra, rb = (ra - 1).astype(np.int64), (rb - 1).astype(np.int64)
    for g, (a, b, d) in enumerate(vc.entries()):
        a, b = a.astype(np.int64), b.astype(np.int64)
        assert np.all((owner[a] == g) | (owner[b] == g)), "a ghost-ghost pair was emitted"
        key = np.minimum(a, b) * n + np.maximum(a, b)
        o = np.argsort(key)
        key, d = key[o], d[o]
        assert len(np.unique(key)) == len(key)
        want = (owner[ra] == g) | (owner[rb] == g)
        wk = np.minimum(ra[want], rb[want]) * n + np.maximum(ra[want], rb[want])
        o2 = np.argsort(wk)
        assert np.array_equal(key, wk[o2]), "rank %d: pairs touching its owned atoms differ from the oracle" % g
        assert np.array_equal(d.view(np.uint32), rd[want][o2].view(np.uint32))
    h.close()
    vc.close()


@pytest.mark.gpu
def test_two_processes_two_gpus():
    """The real multi-process path (skipped on a one-GPU box): two ranks under torch.distributed.run, one GPU each — CUDA IPC
    mapping, flags and pulls across GPUs, migration — must give the oracle's exact pair set (digest of the union of the ranks'
    lists), keep every atom owned exactly once, and reproduce the single-GPU energies."""
    import json
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mg_two_rank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("TWO_RANK_RESULT ")]
    assert r.returncode == 0 and line, (r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(line[0][len("TWO_RANK_RESULT "):])
    one = res["single_gpu"]
    for key in ("migrate_every_0", "migrate_every_5"):
        got = res[key]
        assert got["parity"]["count_match"] and got["parity"]["xor_match"] and got["parity"]["sum_match"], (key, got)
        assert got["every_atom_owned_once"] and got["atoms"] == 64000, (key, got)
        assert abs(got["ke"] - one["ke"]) < 1e-4 * abs(one["ke"]) and abs(got["pe"] - one["pe"]) < 1e-4 * abs(one["pe"]), (key, got, one)
    assert res["migrate_every_0"]["moved_rank"] == 0 and res["migrate_every_5"]["moved_rank"] > 0
