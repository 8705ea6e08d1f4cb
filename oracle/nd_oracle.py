"""ctypes wrapper around the CPU oracle (oracle/nd_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py — never by the product package.

Each function names the reference routine it restates (file:line in nd_oracle.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnd_oracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only; no GPU, no reference needed)."""
    src = os.path.join(_HERE, "nd_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libnd_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.nd_hardware_threads.restype = C.c_int32
    L.nd_mortoncodes.argtypes = [_f32p, C.c_int32, _i32p]
    L.nd_sortperm.argtypes = [_i32p, C.c_int32, _i32p]
    L.nd_spec.argtypes = [C.c_float, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.nd_spec.restype = C.c_int32
    L.nd_tree.argtypes = [_f32p, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32, _f32p, _f32p, _i32p, _i32p,
                          _i32p, _i32p, _f32p]
    L.nd_tree.restype = C.c_int32
    L.nd_build_traverse.argtypes = [_f32p, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.POINTER(C.c_int32)]
    L.nd_build_traverse.restype = C.c_void_p
    L.nd_brute_force.argtypes = [_f32p, C.c_int32, C.c_float, C.c_int32, C.c_int32]
    L.nd_brute_force.restype = C.c_void_p
    L.nd_pairs_count.argtypes = [C.c_void_p]
    L.nd_pairs_count.restype = C.c_int64
    L.nd_pairs_seconds_build.argtypes = [C.c_void_p]
    L.nd_pairs_seconds_build.restype = C.c_double
    L.nd_pairs_seconds_traverse.argtypes = [C.c_void_p]
    L.nd_pairs_seconds_traverse.restype = C.c_double
    L.nd_pairs_copy.argtypes = [C.c_void_p, _i32p, _i32p, _f32p]
    L.nd_pairs_free.argtypes = [C.c_void_p]
    L.nd_digest_pairs.argtypes = [_i32p, _i32p, _f32p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    L.nd_cellgrid_digest.argtypes = [_f32p, C.c_int32, C.c_float, C.c_int32, C.POINTER(C.c_int64),
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.c_void_p]
    L.nd_force_lennardjones.argtypes = [_f32p, C.c_int32, _i32p, _i32p, _f32p, C.c_int64]
    L.nd_force_coulomb.argtypes = [_f32p, C.c_int32, _i32p, _i32p, _f32p, C.c_int64, _f32p]
    L.nd_sum_forces.argtypes = [_f32p, _f32p, _f32p, C.c_int64]
    L.nd_forces_physical_f64.argtypes = [_f32p, C.c_void_p, C.c_int32, _i32p, _i32p, C.c_int64, C.c_double, C.c_double,
                                         C.c_double, C.c_double, C.c_int32, _f64p, _f64p, _f64p]
    L.nd_verlet.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_float]
    L.nd_boundary_reflect.argtypes = [_f32p, _f32p, C.c_int32, _f32p, _f32p]
    L.nd_md_steps_f64.argtypes = [_f64p, _f64p, _f64p, _f32p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_float,
                                  C.c_double, C.c_double, C.c_double, C.c_int32, _f32p, _f32p, C.c_int32, _f64p]
    L.nd_cpu_step.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_int32,
                              C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, _f32p, _f32p, _f64p]
    L.nd_cpu_step.restype = C.c_int64
    _lib = L
    return L


def hardware_threads() -> int:
    return int(lib().nd_hardware_threads())


def _xyz(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


def mortoncodes(xyz) -> np.ndarray:
    """mortoncodes! (BVHTraverse.jl:237-288) — the 10-bit masked key."""
    xyz = _xyz(xyz)
    out = np.empty(len(xyz), np.int32)
    lib().nd_mortoncodes(xyz, len(xyz), out)
    return out


def sortperm(codes) -> np.ndarray:
    """sortperm (BVHTraverse.jl:570), stable, 1-based."""
    codes = np.ascontiguousarray(codes, np.int32)
    out = np.empty(len(codes), np.int32)
    lib().nd_sortperm(codes, len(codes), out)
    return out


class SpecError(ValueError):
    pass


_SPEC_ERRORS = {
    1: "Please use an 'atomsperleaf' that evenly divides into 'atom_count' in BVH Specification",
    2: "Please use more than one leaf in BVH Specification",
    3: "leafTreeData allocates atom_count-1 nodes (BVHTraverse.jl:378): atomsperleaf must be >= 2",
}


def spec(r: float, n: int, apl: int):
    """SpheresBVHSpecs (BVHTraverse.jl:71-94) -> (leaves_count, branches_count)."""
    a, b = C.c_int32(), C.c_int32()
    rc = lib().nd_spec(np.float32(r), n, apl, C.byref(a), C.byref(b))
    if rc:
        raise SpecError(_SPEC_ERRORS[rc])
    return a.value, b.value


def tree(xyz, r: float, apl: int, leaf_variant: bool = True, nthreads: int = 1):
    """leafTreeData (BVHTraverse.jl:544-597) or TreeData (:500-543): dump of the tree."""
    xyz = _xyz(xyz)
    n = len(xyz)
    cap = max(n, 2 * n)
    nmin = np.zeros((cap, 3), np.float32)
    nmax = np.zeros((cap, 3), np.float32)
    left = np.zeros(cap, np.int32)
    skip = np.zeros(cap, np.int32)
    sidx = np.zeros(n, np.int32)
    scode = np.zeros(n, np.int32)
    spos = np.zeros((n, 3), np.float32)
    nn = lib().nd_tree(xyz, n, np.float32(r), apl, int(leaf_variant), nthreads, nmin, nmax, left, skip, sidx, scode, spos)
    if nn < 0:
        raise SpecError(_SPEC_ERRORS[-nn])
    return dict(min=nmin[:nn], max=nmax[:nn], left=left[:nn], skip=skip[:nn], index=sidx, morton=scode, position=spos)


def _take_pairs(ptr):
    L = lib()
    m = L.nd_pairs_count(ptr)
    a = np.empty(m, np.int32)
    b = np.empty(m, np.int32)
    d = np.empty(m, np.float32)
    L.nd_pairs_copy(ptr, a, b, d)
    tb, tt = L.nd_pairs_seconds_build(ptr), L.nd_pairs_seconds_traverse(ptr)
    L.nd_pairs_free(ptr)
    return a, b, d, (tb, tt)


def leafbuild_traverse_bvh(xyz, r: float, apl: int = 4, nthreads: int = 1, qstride: int = 1, timings=False):
    """leafbuild_traverse_bvh (BVHTraverse.jl:1423-1428): (a, b, d), 1-based original ids."""
    xyz = _xyz(xyz)
    rc = C.c_int32()
    p = lib().nd_build_traverse(xyz, len(xyz), np.float32(r), apl, 1, nthreads, qstride, C.byref(rc))
    if rc.value:
        lib().nd_pairs_free(p)
        raise SpecError(_SPEC_ERRORS[rc.value])
    a, b, d, t = _take_pairs(p)
    return (a, b, d, t) if timings else (a, b, d)


def build_traverse_bvh(xyz, r: float, apl: int = 1, nthreads: int = 1):
    """build_traverse_bvh (BVHTraverse.jl:1416-1421), atom-query variant restated as intended."""
    xyz = _xyz(xyz)
    rc = C.c_int32()
    p = lib().nd_build_traverse(xyz, len(xyz), np.float32(r), apl, 0, nthreads, 1, C.byref(rc))
    if rc.value:
        lib().nd_pairs_free(p)
        raise SpecError(_SPEC_ERRORS[rc.value])
    return _take_pairs(p)[:3]


def brute_force(xyz, r: float, predicate: str = "d2", nthreads: int = 0):
    """All-pairs list.  predicate 'd2': d2 < fl(r*r) (BVHTraverse.jl:1027,1248);
    'sqrt': sqrt(d2) < r = threshold_pairs(unique_pairs(p), r) (AllToAll.jl:65-88)."""
    xyz = _xyz(xyz)
    nt = nthreads or hardware_threads()
    p = lib().nd_brute_force(xyz, len(xyz), np.float32(r), 1 if predicate == "sqrt" else 0, nt)
    return _take_pairs(p)[:3]


def canonical(a, b, d):
    """Sort a pair list into canonical (min,max) order -> (lo, hi, d) arrays."""
    a = np.asarray(a, np.int64)
    b = np.asarray(b, np.int64)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    order = np.lexsort((hi, lo))
    return lo[order].astype(np.int32), hi[order].astype(np.int32), np.asarray(d, np.float32)[order]


def digest_pairs(a, b, d):
    a = np.ascontiguousarray(a, np.int32)
    b = np.ascontiguousarray(b, np.int32)
    d = np.ascontiguousarray(d, np.float32)
    cnt, xh, sh, sd = C.c_int64(), C.c_uint64(), C.c_uint64(), C.c_double()
    lib().nd_digest_pairs(a, b, d, len(a), C.byref(cnt), C.byref(xh), C.byref(sh), C.byref(sd))
    return dict(count=cnt.value, xor=xh.value, sum=sh.value, sum_d=sd.value)


def cellgrid_digest(xyz, r: float, nthreads: int = 0, per_atom: bool = False):
    """Independent O(N) exact search (cell grid, N5 predicate) -> digest of the pair set."""
    xyz = _xyz(xyz)
    nt = nthreads or hardware_threads()
    cnt, xh, sh, sd = C.c_int64(), C.c_uint64(), C.c_uint64(), C.c_double()
    pac = np.zeros(len(xyz), np.int32) if per_atom else None
    lib().nd_cellgrid_digest(xyz, len(xyz), np.float32(r), nt, C.byref(cnt), C.byref(xh), C.byref(sh), C.byref(sd),
                             pac.ctypes.data if per_atom else None)
    out = dict(count=cnt.value, xor=xh.value, sum=sh.value, sum_d=sd.value)
    if per_atom:
        out["per_atom"] = pac
    return out


def force_lennardjones(n: int, a, b, d) -> np.ndarray:
    """force_lennardjones! (Forces.jl:15-45), literal."""
    f = np.zeros((n, 3), np.float32)
    a = np.ascontiguousarray(a, np.int32)
    lib().nd_force_lennardjones(f, n, a, np.ascontiguousarray(b, np.int32), np.ascontiguousarray(d, np.float32), len(a))
    return f


def force_coulomb(n: int, a, b, d, charge) -> np.ndarray:
    """force_coulomb! (Forces.jl:56-66), literal and order dependent."""
    f = np.zeros((n, 3), np.float32)
    a = np.ascontiguousarray(a, np.int32)
    lib().nd_force_coulomb(f, n, a, np.ascontiguousarray(b, np.int32), np.ascontiguousarray(d, np.float32), len(a),
                           np.ascontiguousarray(charge, np.float32))
    return f


def sum_forces(f1, f2) -> np.ndarray:
    """sum_forces! (Forces.jl:68-75)."""
    f1 = np.ascontiguousarray(f1, np.float32)
    f2 = np.ascontiguousarray(f2, np.float32)
    out = np.empty_like(f1)
    lib().nd_sum_forces(out, f1, f2, f1.size)
    return out


def forces_physical_f64(xyz, charge, a, b, eps, sigma, kc, rc, shift=True):
    """fp64 LJ 12-6 + Coulomb over a half pair list -> (force[n,3], pe_atom[n], scale[n])."""
    xyz = _xyz(xyz)
    n = len(xyz)
    f = np.zeros((n, 3), np.float64)
    pe = np.zeros(n, np.float64)
    sc = np.zeros(n, np.float64)
    q = None if charge is None else np.ascontiguousarray(charge, np.float32)
    a = np.ascontiguousarray(a, np.int32)
    lib().nd_forces_physical_f64(xyz, None if q is None else q.ctypes.data, n, a, np.ascontiguousarray(b, np.int32),
                                 len(a), eps, sigma, kc, rc, int(shift), f, pe, sc)
    return f, pe, sc


def verlet(pos, vel, force, force_next, mass, dt):
    """Velocity-Verlet body (Simulator.jl:198-222), literal Float32; returns (pos, vel)."""
    pos = _xyz(pos).copy()
    vel = _xyz(vel).copy()
    lib().nd_verlet(pos, vel, _xyz(force), _xyz(force_next), np.ascontiguousarray(mass, np.float32), len(pos),
                    np.float32(dt))
    return pos, vel


def boundary_reflect(pos, vel, bmin, bmax):
    """boundary_reflect! (Simulator.jl:81-111); returns (pos, vel)."""
    pos = _xyz(pos).copy()
    vel = _xyz(vel).copy()
    lib().nd_boundary_reflect(pos, vel, len(pos), np.asarray(bmin, np.float32), np.asarray(bmax, np.float32))
    return pos, vel


def rescale_velocity(vel, tf, gamma, mass, objectcount):
    """rescale_velocity!(velocity, Tf, gamma, mass, objectcount) (Simulator.jl:119-144), Float32 like Julia:
    Ti accumulates (2/(3*objects*kb)) * |v| * m/2 atom by atom (kb = 1; the SPEED, not its square), then
    every velocity is multiplied by beta = (1 + gamma*(Tf/Ti - 1))^0.5.  Returns (scaled velocities, Ti, beta).
    PARITY UNPINNED: the reference has no test or expected value for this function."""
    f = np.float32
    v = np.ascontiguousarray(vel, f)
    m = np.ascontiguousarray(mass, f)
    sq = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]  # Float32, left fold
    # `x ^ 0.5` has a Float64 exponent: Julia promotes the Float32 base, the power is Float64, and the assignment to v::T
    # rounds it back to Float32 (:133)
    speed = np.sqrt(sq.astype(np.float64)).astype(f)
    coef = f(2.0) / (f(3.0) * f(objectcount) * f(1.0))
    term = ((coef * speed) * m) / f(2.0)
    ti = np.add.accumulate(term, dtype=f)[-1] if len(term) else f(0)  # sequential Float32 accumulation
    # beta = (Float32 expression) ^ 0.5 is a Float64 (:140), and `velocity[each] .*= beta` multiplies in Float64 before the
    # store rounds to Float32 (:145)
    beta = np.sqrt(np.float64(f(1.0) + f(gamma) * (f(tf) / ti - f(1.0))))
    return (v.astype(np.float64) * beta).astype(f), float(ti), float(beta)


# ---- system setup (MDInput.jl) with the counter-based draws of nb200_collect_objects ------------------------------------
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123
    library's philox4x32_R(10, ...)).  counter: (..., 4) uint32, key: (k0, k1) -> (..., 4) uint32.  Pinned by the
    library's published known-answer vectors in tests/test_oracle_golden.py."""
    c = np.asarray(counter, np.uint32).astype(np.uint64)
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> s32, p0 & mask
        hi1, lo1 = p1 >> s32, p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], -1).astype(np.uint32)


def _u53(hi, lo):
    """Float64 in [0,1) on Julia's rand(Float64) grid: the top 53 of 64 random bits times 2^-53."""
    w = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return (w >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def _uniform64(a, b, u):
    """rand(Uniform(a, b)) with Float32 bounds (Distributions.jl: a + (b - a) * rand()): width in Float32, rest Float64."""
    f = np.float32
    return np.float64(f(a)) + np.float64(f(b) - f(a)) * u


def _setup_block(idx, stream, rnd, seed):
    ctr = np.zeros((len(idx), 4), np.uint32)
    ctr[:, 0] = idx
    ctr[:, 1] = stream
    ctr[:, 2] = rnd
    return philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


def setup_positions(idx, rnd, seed, bmin, bmax):
    """generate_positions / generate_onePosition (MDInput.jl:175-190, 208-226): Float64 uniform draws per axis, stored
    as Float32.  Draw number `rnd` of atoms `idx` (0-based)."""
    idx = np.asarray(idx, np.uint32)
    a = _setup_block(idx, 0, rnd, seed)
    b = _setup_block(idx, 1, rnd, seed)
    out = np.empty((len(idx), 3), np.float32)
    out[:, 0] = _uniform64(bmin[0], bmax[0], _u53(a[:, 0], a[:, 1]))
    out[:, 1] = _uniform64(bmin[1], bmax[1], _u53(a[:, 2], a[:, 3]))
    out[:, 2] = _uniform64(bmin[2], bmax[2], _u53(b[:, 0], b[:, 1]))
    return out


def collect_objects(n, seed, bmin, bmax, minmass, maxmass, mincharge, maxcharge, temperature, randomvelocity,
                    minimumdistance, max_rounds=0):
    """collect_objects(Collector::GenericRandomCollector) (MDInput.jl:305-369) on the draws of nb200_collect_objects.
    mass/charge (:307-312), velocity (:319-336: Float32 left fold, Float64 mass draw divides last), positions (:175-190),
    and generate_pruned_positions! (:260-283) as it is meant to work: while unique_pairs_prune (:228-258) finds pairs
    closer than minimumdistance, the lower-numbered atom of each gets a new position.  The too-close test is the BVH
    pair predicate d2 < fl(r*r) (brute_force 'd2'); the reference's prune writes sqrt(dx^2+dy^2+dz^2) < threshold.
    PARITY UNPINNED against the reference (unseeded global RNG, no test, loop never entered at HEAD: `tooClose < 0`).
    Returns dict(position, velocity, mass, charge, rounds, redrawn)."""
    f = np.float32
    idx = np.arange(n, dtype=np.uint32)
    seed = int(seed)
    m = _setup_block(idx, 2, 0, seed)
    mass64 = _uniform64(minmass, maxmass, _u53(m[:, 0], m[:, 1]))
    charge64 = _uniform64(mincharge, maxcharge, _u53(m[:, 2], m[:, 3]))
    v = _setup_block(idx, 3, 0, seed)
    veldist = ((v[:, :3] >> np.uint32(8)).astype(f) * f(2.0 ** -24)).astype(f)   # rand(Float32, n), one column per axis
    vel = np.empty((n, 3), f)
    nf = f(n)
    for d in range(3):
        if randomvelocity:
            total = f(np.sum(veldist[:, d].astype(np.float64)))       # exact sum, rounded once
            share = (veldist[:, d] / total).astype(f)
            t = (((f(temperature) * share).astype(f) * f(3)).astype(f) * nf).astype(f)
        else:
            t = np.full(n, ((f(temperature) / nf) * f(3)) * nf, f)
        vel[:, d] = (t.astype(np.float64) / mass64).astype(f)
    pos = setup_positions(idx, 0, seed, bmin, bmax)
    rounds, redrawn = 0, 0
    limit = max_rounds if max_rounds > 0 else 10 * n
    if minimumdistance > 0 and n > 1:
        while True:
            a, b, _ = brute_force(pos, minimumdistance, "d2")
            if len(a) == 0:
                break
            if rounds >= limit:
                raise RuntimeError("Objects could not be placed, increase box size, reduce object count, or decrease "
                                   "minimum spawning distance")
            rounds += 1
            marked = np.unique(np.minimum(a, b) - 1)                  # brute_force ids are 1-based
            pos[marked] = setup_positions(marked, rounds, seed, bmin, bmax)
            redrawn += len(marked)
    return dict(position=pos, velocity=vel, mass=mass64.astype(f), charge=charge64.astype(f), rounds=rounds, redrawn=redrawn)


def md_steps_f64(pos, vel, mass, charge, nsteps, dt, cutoff, eps, sigma, kc, shift, bmin, bmax, force=None):
    """fp64 kick-drift-kick velocity Verlet with reflective walls (integrator oracle)."""
    pos = np.ascontiguousarray(pos, np.float64).copy()
    vel = np.ascontiguousarray(vel, np.float64).copy()
    n = len(pos)
    init = force is None
    f = np.zeros((n, 3), np.float64) if init else np.ascontiguousarray(force, np.float64).copy()
    q = None if charge is None else np.ascontiguousarray(charge, np.float32)
    en = np.zeros(2, np.float64)
    lib().nd_md_steps_f64(pos, vel, f, np.ascontiguousarray(mass, np.float32), None if q is None else q.ctypes.data, n,
                          nsteps, dt, np.float32(cutoff), eps, sigma, kc, int(shift), np.asarray(bmin, np.float32),
                          np.asarray(bmax, np.float32), int(init), en)
    return pos, vel, f, dict(ke=en[0], pe=en[1])


def cpu_step(pos, vel, force, mass, charge, dt, cutoff, apl, nthreads, qstride, eps, sigma, kc, bmin, bmax):
    """One reference-shaped CPU MD step (timed baseline).  Mutates pos/vel/force in place.
    Returns (pair_count, timings[tree, traverse, force, verlet, total])."""
    t = np.zeros(5, np.float64)
    q = None if charge is None else np.ascontiguousarray(charge, np.float32)
    np_ = lib().nd_cpu_step(pos, vel, force, mass, None if q is None else q.ctypes.data, len(pos), np.float32(dt),
                            np.float32(cutoff), apl, nthreads, qstride, np.float32(eps), np.float32(sigma),
                            np.float32(kc), np.asarray(bmin, np.float32), np.asarray(bmax, np.float32), t)
    if np_ < 0:
        raise SpecError(_SPEC_ERRORS[-np_])
    return int(np_), t
