"""Host-side mirror of the NaiveDynamics.jl API for the hot path, on top of the C ABI.

The reference's host language is Julia; no Julia exists in this image, so this module mirrors the
reference's interface for the path in Python — same names (Julia's trailing `!` becomes `_`), same
argument meaning, same error behaviour — so the parity tests read like the reference's own tests
(test/BVHTraverse.jl).  The Julia package extension that binds the identical C ABI with `ccall` is
julia/NaiveB200.jl.

Everything here routes to libnaiveb200.so.  There is no CPU path.
"""
from __future__ import annotations

import os
from collections import namedtuple
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib
from ._lib import Handle, NB200Error

Float32 = np.float32
Float64 = np.float64


class B200Backend:
    """Dispatch tag, the analogue of the KernelAbstractions `backend` argument of
    gpubvh_neighborlist(backend, position, spec) (ext/NaiveKA.jl:470)."""

    def __init__(self, device: int = 0):
        self.device = device


# ----------------------------------------------------------------------------------------------
# SpheresBVHSpecs (src/Neighbors/BVHTraverse.jl:63-94)
# ----------------------------------------------------------------------------------------------
class SpheresBVHSpecs:
    """SpheresBVHSpecs(; neighbor_distance, atom_count, floattype, atomsperleaf).

    Same validation and messages as BVHTraverse.jl:74-80.  `atomsperleaf` is accepted for
    compatibility: the GPU tree always uses 32-atom leaves and the pair set does not depend on it
    (the reference's own tests assert that, test/BVHTraverse.jl:246-265)."""

    def __init__(self, *, neighbor_distance, atom_count, floattype=Float32, atomsperleaf):
        if atom_count % atomsperleaf != 0:
            raise ValueError("Please use an 'atomsperleaf' that evenly divides into 'atom_count' in BVH Specification")
        leaves_count = atom_count / atomsperleaf
        if leaves_count < 2:
            raise ValueError("Please use more than one leaf in BVH Specification")
        if floattype not in (Float32, np.dtype("float32")):
            # Float64 specs construct in the reference but mortoncodes! then errors (:285-287)
            raise ValueError("K-type integer is not implemented")
        self.neighbor_distance = Float32(neighbor_distance)
        self.atom_count = int(atom_count)
        self.leaves_count = int(leaves_count)
        self.branches_count = int(leaves_count) - 1
        self.atomsperleaf = int(atomsperleaf)


# ----------------------------------------------------------------------------------------------
# handle cache: one handle per (device, capacity class)
# ----------------------------------------------------------------------------------------------
_handles: dict = {}


def get_handle(n: int, device: int = 0) -> Handle:
    cap = 1 << max(10, int(np.ceil(np.log2(max(n, 2)))))
    key = (device, cap)
    h = _handles.get(key)
    if h is None:
        h = Handle(cap, device=device)
        _handles[key] = h
    return h


def release_handles():
    for h in _handles.values():
        h.close()
    _handles.clear()


def _positions(position) -> np.ndarray:
    """Vec3D{Float32} (Vector{MVector{3,Float32}}, src/MDInput.jl:29) -> contiguous (n,3) float32."""
    a = np.ascontiguousarray(position, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("position must be n x 3")
    return a


class PairList:
    """Vector{Tuple{Int32,Int32,Float32}} (BVHTraverse.jl:1328) held as three arrays."""

    def __init__(self, a, b, d):
        self.a, self.b, self.d = a, b, d

    def __len__(self):
        return len(self.a)

    def __getitem__(self, k):
        return (int(self.a[k]), int(self.b[k]), Float32(self.d[k]))

    def __iter__(self):
        return iter(zip(self.a.tolist(), self.b.tolist(), self.d))

    def sorted(self):
        """sort!(list, by=x->x[2]); sort!(list, by=x->x[1]) after (min,max) canonicalisation."""
        lo = np.minimum(self.a, self.b)
        hi = np.maximum(self.a, self.b)
        order = np.lexsort((hi, lo))
        return PairList(lo[order], hi[order], self.d[order])


# ----------------------------------------------------------------------------------------------
# neighbour search entry points (BVHTraverse.jl:1416-1428, PkgExtensions.jl:66)
# ----------------------------------------------------------------------------------------------
def _search(position, spec: SpheresBVHSpecs, device: int = 0) -> PairList:
    pos = _positions(position)
    if len(pos) != spec.atom_count:
        raise ValueError("spec.atom_count does not match the number of positions")
    h = get_handle(len(pos), device)
    h.neighbors(pos, spec.neighbor_distance)
    a, b, d = h.get_pairs(index_base=1)
    return PairList(a, b, d)


def leafbuild_traverse_bvh(position, spec: SpheresBVHSpecs) -> PairList:
    """leafbuild_traverse_bvh(position, spec) (BVHTraverse.jl:1423-1428)."""
    return _search(position, spec)


def build_traverse_bvh(position, spec: SpheresBVHSpecs) -> PairList:
    """build_traverse_bvh(position, spec) (BVHTraverse.jl:1416-1421)."""
    return _search(position, spec)


NeighborResult = namedtuple("NeighborResult", ["pairlist", "treedata"])


def gpubvh_neighborlist(backend: B200Backend, position, spec: SpheresBVHSpecs) -> NeighborResult:
    """gpubvh_neighborlist(backend, position, spec) -> (pairlist, treedata) (ext/NaiveKA.jl:470-558)."""
    pl = _search(position, spec, backend.device)
    h = get_handle(spec.atom_count, backend.device)
    tree = h.get_tree()
    tree["sorted_ids"] = h.get_sorted_ids() + 1
    return NeighborResult(pl, tree)


def TreeData(position, spec: SpheresBVHSpecs):
    """TreeData(position, spec) (BVHTraverse.jl:500-543): builds the tree (and, on the GPU, the
    list in the same pass).  Returns the dump of the GPU tree."""
    return gpubvh_neighborlist(B200Backend(), position, spec).treedata


# ----------------------------------------------------------------------------------------------
# forces (src/Forces.jl)
# ----------------------------------------------------------------------------------------------
def _pairs(pairslist):
    if isinstance(pairslist, PairList):
        return pairslist.a, pairslist.b, pairslist.d
    a, b, d = pairslist
    return a, b, d


def force_lennardjones_(force: np.ndarray, pairslist, position=None, device: int = 0):
    """force_lennardjones!(force, pairslist, position) (Forces.jl:15-45), literal. In place."""
    a, b, d = _pairs(pairslist)
    n = len(force)
    force[...] = get_handle(n, device).force_lennardjones(n, a, b, d, index_base=1)
    return None


def force_coulomb_(force: np.ndarray, pairslist, charge, device: int = 0):
    """force_coulomb!(force, pairslist, charge) (Forces.jl:56-66), literal (order dependent). In place."""
    a, b, d = _pairs(pairslist)
    n = len(force)
    force[...] = get_handle(n, device).force_coulomb(n, a, b, d, charge, index_base=1)
    return None


def sum_forces_(force: np.ndarray, force1, force2, device: int = 0):
    """sum_forces!(force, force1, force2) (Forces.jl:68-75). In place; returns force."""
    force[...] = get_handle(len(force), device).sum_forces(force1, force2)
    return force


# ----------------------------------------------------------------------------------------------
# system setup (src/MDInput.jl) — host-side, as in the reference
# ----------------------------------------------------------------------------------------------
@dataclass
class GenericRandomCollector:
    """GenericRandomCollector(; ...) (MDInput.jl:67-102)."""
    objectnumber: int
    minDim: tuple
    maxDim: tuple
    temperature: float
    randomvelocity: bool
    minmass: float
    maxmass: float
    minimumdistance: float
    mincharge: float
    maxcharge: float
    floattype: type = Float32
    pregeneratedposition: bool = False
    seed: Optional[int] = None  # the reference draws from the global RNG; a seed makes tests repeatable


@dataclass
class GenericObjectCollection:
    """GenericObjectCollection{T} (MDInput.jl:33-46) with flat arrays."""
    currentstep: np.ndarray
    name: list
    mass: np.ndarray
    charge: np.ndarray
    radius: np.ndarray
    index: np.ndarray
    position: np.ndarray
    velocity: np.ndarray
    force: np.ndarray


def generate_positions(collector: GenericRandomCollector, rng=None) -> np.ndarray:
    """generate_positions(Collector) (MDInput.jl:175-190): Float64 uniform draws -> Float32."""
    rng = rng or np.random.default_rng(collector.seed)
    n = collector.objectnumber
    lo = np.asarray(collector.minDim, np.float64)
    hi = np.asarray(collector.maxDim, np.float64)
    xyz = np.empty((n, 3), np.float64)
    for d in range(3):  # x, then y, then z, each one rand(Uniform, n) call (:183-185)
        xyz[:, d] = rng.uniform(lo[d], hi[d], n)
    return xyz.astype(np.float32)


def collect_objects(collector: GenericRandomCollector, position=None, backend=None, cutoff: Optional[float] = None,
                    model: Optional["ForceModel"] = None, max_rounds: int = 0) -> GenericObjectCollection:
    """collect_objects(Collector; position) (MDInput.jl:305-369).  Velocity rule as :319-336.

    Without `backend` the draws are made on the host with numpy (test inputs).  With a B200Backend the system is drawn
    ON THE DEVICE (nb200_collect_objects): masses, charges, velocities, positions, and the re-draw of atoms closer than
    collector.minimumdistance (generate_pruned_positions!, :260-283) through the BVH search instead of the O(N^2) loop.
    The system stays resident in the handle simulate_bvh_ uses (same atom count), with the pair model `model` and the
    neighbour cutoff `cutoff` (default: minimumdistance, or 0.03 if that is 0) set for its first forces."""
    if backend is not None:
        if collector.pregeneratedposition:
            raise ValueError("pregeneratedposition=true: upload the positions with set_system instead")
        n = collector.objectnumber
        h = get_handle(n, backend.device)
        model = model or ForceModel()
        r = float(cutoff if cutoff is not None else (collector.minimumdistance or 0.03))
        h.set_box(collector.minDim, collector.maxDim)
        h.set_forcefield(model.eps, model.sigma, model.kcoul, r, model.shift)
        seed = collector.seed if collector.seed is not None else int.from_bytes(os.urandom(8), "little")
        mass, charge, _, _ = h.collect_objects(n, seed, collector.minmass, collector.maxmass, collector.mincharge,
                                               collector.maxcharge, collector.temperature, collector.randomvelocity,
                                               collector.minimumdistance, max_rounds)
        T = collector.floattype
        return GenericObjectCollection(
            currentstep=np.full(n, 1, np.int64), name=["duck"] * n, mass=mass.astype(T), charge=charge.astype(T),
            radius=np.full(n, 0.01, T), index=np.arange(1, n + 1, dtype=np.int64), position=h.get_positions(),
            velocity=h.get_velocities(), force=np.zeros((n, 3), T))
    rng = np.random.default_rng(collector.seed)
    n = collector.objectnumber
    T = collector.floattype
    mass = rng.uniform(collector.minmass, collector.maxmass, n)
    charge = rng.uniform(collector.mincharge, collector.maxcharge, n)
    velocity = np.zeros((n, 3), T)
    kb = 1
    if collector.randomvelocity:
        for i in range(3):
            veldist = rng.random(n).astype(T)
            veldist /= veldist.sum()
            velocity[:, i] = collector.temperature * veldist * 3 * n * kb / mass
    else:
        for i in range(3):
            velocity[:, i] = collector.temperature / n * 3 * n * kb / mass
    if collector.pregeneratedposition:
        if position is None:
            raise ValueError("pregeneratedposition=true needs position")
        pos = _positions(position).copy()
    else:
        pos = generate_positions(collector, rng)
    return GenericObjectCollection(
        currentstep=np.full(n, 1, np.int64), name=["duck"] * n, mass=mass.astype(T), charge=charge.astype(T),
        radius=np.full(n, 0.01, T), index=np.arange(1, n + 1, dtype=np.int64), position=pos,
        velocity=velocity.astype(T), force=np.zeros((n, 3), T))


# ----------------------------------------------------------------------------------------------
# simulation loops (src/Simulator.jl)
# ----------------------------------------------------------------------------------------------
@dataclass
class SimSpec:
    """SimSpec(; inttype, floattype, duration, stepwidth, currentstep, logLength, vDamp, threshold)
    (Simulator.jl:30-49)."""
    duration: int
    stepwidth: float
    currentstep: int = 1
    logLength: int = 10
    vDamp: float = 1.0
    threshold: float = 0.03
    inttype: type = np.int64
    floattype: type = Float32


@dataclass
class ForceModel:
    """Physical pair model of the device step loop (see DESIGN.md "Forces").  eps = kcoul = 0 gives
    the reference's force-free simulate_bvh! loop."""
    eps: float = 0.0
    sigma: float = 1.0
    kcoul: float = 0.0
    shift: bool = True


def boundary_reflect_(position: np.ndarray, velocity: np.ndarray, collector: GenericRandomCollector, device: int = 0):
    """boundary_reflect!(position, velocity, collector) (Simulator.jl:81-111). In place."""
    n = len(position)
    zeros = np.zeros((n, 3), np.float32)
    ones = np.ones(n, np.float32)
    # dt = 0 turns the Verlet body into the identity, leaving only the reflection
    p, v = get_handle(n, device).verlet_update(position, velocity, zeros, zeros, ones, 0.0, collector.minDim, collector.maxDim)
    position[...] = p
    velocity[...] = v


def rescale_velocity_(velocity: np.ndarray, Tf: float, gamma: float, mass: np.ndarray, objectcount: int, device: int = 0):
    """rescale_velocity!(velocity, Tf, γ, mass, objectcount) (Simulator.jl:119-144). In place."""
    n = len(velocity)
    # A handle of its own: the cached handle of get_handle may hold a system a previous call left resident (collect_objects,
    # simulate), and this stand-alone form needs a dummy system and force field that must not replace it.
    h = Handle(max(n, 2), device=device)
    try:
        h.set_forcefield(0.0, 1.0, 0.0, 1e-6, True)  # velocities only: no pair model needed
        h.set_system(np.zeros((n, 3), np.float32) + (np.arange(n, dtype=np.float32)[:, None] + 0.5) / n, velocity, mass, None)
        h.rescale_velocity(Tf, gamma, physical=False)
        velocity[...] = h.get_velocities()
    finally:
        h.close()


def simulate_bvh_(sys: GenericObjectCollection, spec: SimSpec, bvhspec: SpheresBVHSpecs, clct: GenericRandomCollector,
                  model: Optional[ForceModel] = None, log_every: int = 1, rescale_every: int = 0, device: int = 0):
    """simulate_bvh!(sys, spec, bvhspec, clct) (Simulator.jl:327-379): velocity Verlet with the
    neighbour list rebuilt from a fresh BVH every step.  The reference never computes forces in this
    loop (all zero) — that is the default here too (model=None); pass a ForceModel for LJ/Coulomb.
    Returns poslog: list of (n,3) arrays, entry 0 = initial positions (Simulator.jl:340).
    The whole loop is ONE library call (nb200_simulate): the frames are copied out asynchronously while the
    following steps run."""
    model = model or ForceModel()
    n = len(sys.position)
    h = get_handle(n, device)
    h.set_box(clct.minDim, clct.maxDim)
    h.set_forcefield(model.eps, model.sigma, model.kcoul, float(bvhspec.neighbor_distance), model.shift)
    h.set_system(sys.position, sys.velocity, sys.mass, sys.charge)
    frames = h.simulate(int(spec.duration), float(spec.stepwidth), log_every=log_every, rescale_every=rescale_every,
                        target_temperature=float(clct.temperature), gamma=float(spec.vDamp))
    poslog = [np.array(sys.position, copy=True)] + [f for f in frames]
    rest = int(spec.duration) % log_every if log_every > 0 else 0
    sys.position[...] = h.get_positions() if (rest or not len(frames)) else poslog[-1]
    sys.velocity[...] = h.get_velocities()
    sys.force[...] = h.get_forces()
    return poslog


def simulate_(sys: GenericObjectCollection, spec: SimSpec, clct: GenericRandomCollector,
              model: Optional[ForceModel] = None, cutoff: Optional[float] = None, device: int = 0):
    """simulate!(sys, spec, clct) (Simulator.jl:154-256).  The reference drives this loop from an
    O(N^2) pair list; here the list comes from the BVH search at `cutoff` (default spec.threshold),
    with the physical LJ+Coulomb model (DESIGN.md "Forces" documents the divergence from
    Forces.jl's literal formulas), and rescale_velocity! every 10th step as in the reference (:241-243).
    Returns poslog of length duration+1 like the reference (:167,245)."""
    model = model or ForceModel(eps=1.0, sigma=float(spec.threshold) / 2.5, kcoul=1.0)
    r = float(cutoff if cutoff is not None else spec.threshold)
    bvhspec = SpheresBVHSpecs(neighbor_distance=r, atom_count=len(sys.position), floattype=Float32, atomsperleaf=1)
    return simulate_bvh_(sys, spec, bvhspec, clct, model=model, rescale_every=10, device=device)
