"""ctypes binding of libnaiveb200.so (include/naiveb200.h).

This is the same ABI the Julia shim (julia/NaiveB200.jl) calls with `ccall`; Python is only the
host language of this repo's tests and bench because no Julia toolchain exists in the image.
There is NO fallback: if the shared library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NAIVEB200_LIB", os.path.join(_HERE, "libnaiveb200.so"))

NB200_OK = 0
NB200_ERR_BAD_ARG = 1
NB200_ERR_CUDA = 2
NB200_ERR_PAIR_OVERFLOW = 3
NB200_ERR_STATE = 4
NB200_ERR_CAPACITY = 5

NB200_LIST_DIRECTED = 0
NB200_LIST_HALF = 1

STAGES = ("integrate", "morton", "sort", "reorder", "build", "traverse", "force", "export")
LEAF_SIZE = 32

_f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_H = C.c_void_p
_vp = C.c_void_p  # nullable pointer arguments


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_atoms", "n_leaves", "n_entries", "n_segments", "entry_capacity",
                                         "kernel_launches", "steps_done", "regrows", "n_pairs", "list_half", "n_slots")]


# name -> (restype, argtypes): every symbol include/naiveb200.h declares
SIGNATURES = {
    "nb200_version": (C.c_int32, []),
    "nb200_device_count": (C.c_int32, []),
    "nb200_create": (C.c_int32, [C.c_int32, C.c_int64, C.c_int64, C.POINTER(_H)]),
    "nb200_destroy": (C.c_int32, [_H]),
    "nb200_last_error": (C.c_char_p, [_H]),
    "nb200_set_box": (C.c_int32, [_H, _f32, _f32]),
    "nb200_neighbors": (C.c_int32, [_H, _vp, C.c_int32, C.c_int32, C.c_float, C.POINTER(C.c_int64)]),
    "nb200_get_pairs": (C.c_int32, [_H, _vp, _vp, _vp, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]),
    "nb200_force_lennardjones": (C.c_int32, [_H, _f32, C.c_int32, _vp, _vp, _vp, C.c_int64, C.c_int32]),
    "nb200_force_coulomb": (C.c_int32, [_H, _f32, C.c_int32, _vp, _vp, _vp, C.c_int64, _f32, C.c_int32]),
    "nb200_sum_forces": (C.c_int32, [_H, _f32, _f32, _f32, C.c_int64]),
    "nb200_verlet_update": (C.c_int32, [_H, _f32, _f32, _f32, _f32, _f32, C.c_int32, C.c_float, _vp, _vp]),
    "nb200_set_forcefield": (C.c_int32, [_H, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32]),
    "nb200_set_system": (C.c_int32, [_H, _vp, _vp, C.c_int32, _vp, _vp, C.c_int32]),
    "nb200_step": (C.c_int32, [_H, C.c_int32, C.c_float]),
    "nb200_step_async": (C.c_int32, [_H, C.c_int32, C.c_float]),
    "nb200_sync": (C.c_int32, [_H]),
    "nb200_step_host": (C.c_int32, [_H, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_float]),
    "nb200_leapfrog_host_async": (C.c_int32, [_H, _vp, _vp, C.c_int32, C.c_int32, C.c_float, C.c_int32]),
    "nb200_rescale_velocity": (C.c_int32, [_H, C.c_float, C.c_float, C.c_int32]),
    "nb200_simulate": (C.c_int32, [_H, C.c_int32, C.c_float, C.c_int32, _vp, C.c_int32, C.c_int64, C.c_int32, C.c_float, C.c_float,
                                   C.POINTER(C.c_int64)]),
    "nb200_collect_objects": (C.c_int32, [_H, C.c_int32, C.c_uint64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                          C.c_float, C.c_int32, _vp, _vp, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "nb200_get_positions": (C.c_int32, [_H, _vp, C.c_int32]),
    "nb200_get_velocities": (C.c_int32, [_H, _vp, C.c_int32]),
    "nb200_get_forces": (C.c_int32, [_H, _vp, C.c_int32]),
    "nb200_get_energies": (C.c_int32, [_H, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nb200_pair_count": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "nb200_set_stream": (C.c_int32, [_H, _vp]),
    "nb200_mg_set_owned": (C.c_int32, [_H, _vp, _vp, C.c_int32, _vp, _vp, C.c_int32]),
    "nb200_mg_owned_pos_device": (C.c_int32, [_H, C.POINTER(C.c_void_p)]),
    "nb200_mg_integrate": (C.c_int32, [_H, C.c_float]),
    "nb200_mg_search_force": (C.c_int32, [_H, _vp, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nb200_mg_search_force_async": (C.c_int32, [_H]),
    "nb200_mg_sync": (C.c_int32, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nb200_mg_step_async": (C.c_int32, [_H, C.c_int32, C.c_float]),
    "nb200_mg_leapfrog_host_async": (C.c_int32, [_H, _vp, C.c_int32, C.c_float]),
    "nb200_mg_get_owned": (C.c_int32, [_H, _vp, C.c_int32, C.c_int32]),
    "nb200_mg_owned_count": (C.c_int32, [_H, C.POINTER(C.c_int32)]),
    "nb200_mg_republish": (C.c_int32, [_H]),
    "nb200_mg_get_owned_ids": (C.c_int32, [_H, _i32]),
    "nb200_mg_set_migration": (C.c_int32, [_H, _vp, C.c_int32, C.c_int32]),
    "nb200_mg_get_energies": (C.c_int32, [_H, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nb200_mg_get_entries": (C.c_int32, [_H, _vp, _vp, _vp, C.c_int64, C.POINTER(C.c_int64)]),
    "nb200_mg_publication": (C.c_int32, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), _vp]),
    "nb200_mg_connect": (C.c_int32, [_H, C.c_int32, C.c_int32, _i64, _i32, _vp, _vp]),
    "nb200_morton30": (C.c_int32, [_H, _vp, C.c_int32, C.c_int32, _u32]),
    "nb200_set_curve": (C.c_int32, [_H, C.c_int32]),
    "nb200_set_list_mode": (C.c_int32, [_H, C.c_int32]),
    "nb200_set_resort_interval": (C.c_int32, [_H, C.c_int32]),
    "nb200_set_list_reuse": (C.c_int32, [_H, C.c_float, C.c_int32]),
    "nb200_set_fused_force": (C.c_int32, [_H, C.c_int32]),
    "nb200_sort_keys": (C.c_int32, [_H, _vp, C.c_int32, C.c_int32, _u32]),
    "nb200_sort_pairs": (C.c_int32, [_H, _u32, _u32, C.c_int64]),
    "nb200_get_sorted_ids": (C.c_int32, [_H, _i32]),
    "nb200_get_tree": (C.c_int32, [_H, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _vp, _vp, _vp]),
    "nb200_get_neighbor_counts": (C.c_int32, [_H, _i32]),
    "nb200_debug_traverse_profile": (C.c_int32, [_H, _i64]),
    "nb200_set_profiling": (C.c_int32, [_H, C.c_int32]),
    "nb200_get_stage_times": (C.c_int32, [_H, _f64, _i64]),
    "nb200_timer_start": (C.c_int32, [_H]),
    "nb200_timer_stop": (C.c_int32, [_H, C.POINTER(C.c_double)]),
    "nb200_get_stats": (C.c_int32, [_H, C.POINTER(Stats)]),
}

_lib = None


class NB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libnaiveb200 error {code}: {message}")
        self.code = code
        self.message = message


def load():
    """Load the shared library (no GPU needed just to load and resolve symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


def _as_f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None and (a.ndim != 2 or a.shape[1] not in cols):
        raise ValueError(f"expected an (n, {cols}) float32 array, got shape {a.shape}")
    return a


class Handle:
    """Owns one nb200_handle (device memory, stream).  Thin, 1:1 with the C ABI."""

    def __init__(self, n_max: int, device: int = 0, pair_capacity_hint: int = 0):
        self._L = load()
        self._h = _H()
        rc = self._L.nb200_create(int(device), int(n_max), int(pair_capacity_hint), C.byref(self._h))
        if rc != NB200_OK:
            msg = self._L.nb200_last_error(None).decode()
            self._h = None
            raise NB200Error(rc, msg)
        self.n_max = int(n_max)
        self.device = int(device)
        self.n = 0

    # -- plumbing --
    def _check(self, rc: int):
        if rc != NB200_OK:
            raise NB200Error(rc, self._L.nb200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.nb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration --
    def set_box(self, box_min, box_max):
        self._check(self._L.nb200_set_box(self._h, np.asarray(box_min, np.float32), np.asarray(box_max, np.float32)))

    def set_forcefield(self, eps=1.0, sigma=1.0, kcoul=0.0, cutoff=2.5, shift=True):
        self._check(self._L.nb200_set_forcefield(self._h, eps, sigma, kcoul, cutoff, int(bool(shift))))

    # -- neighbour search --
    def neighbors(self, xyz, cutoff: float) -> int:
        xyz = _as_f32(xyz, (3, 4))
        cnt = C.c_int64()
        self._check(self._L.nb200_neighbors(self._h, _ptr(xyz), xyz.shape[1], xyz.shape[0], np.float32(cutoff), C.byref(cnt)))
        self.n = xyz.shape[0]
        return cnt.value

    def neighbors_ptr(self, host_ptr: int, stride: int, n: int, cutoff: float) -> int:
        """Same as neighbors() for a raw host pointer (e.g. pinned torch tensor)."""
        cnt = C.c_int64()
        self._check(self._L.nb200_neighbors(self._h, host_ptr, stride, n, np.float32(cutoff), C.byref(cnt)))
        self.n = n
        return cnt.value

    def pair_count(self) -> int:
        cnt = C.c_int64()
        self._check(self._L.nb200_pair_count(self._h, C.byref(cnt)))
        return cnt.value

    def get_pairs(self, index_base: int = 1, out=None):
        npairs = self.pair_count()
        if out is None:
            a = np.empty(npairs, np.int32)
            b = np.empty(npairs, np.int32)
            d = np.empty(npairs, np.float32)
        else:
            a, b, d = out
        w = C.c_int64()
        self._check(self._L.nb200_get_pairs(self._h, _ptr(a), _ptr(b), _ptr(d), len(a), index_base, C.byref(w)))
        return a[: w.value], b[: w.value], d[: w.value]

    # -- literal reference entry points --
    def force_lennardjones(self, n, a, b, d, index_base=1):
        a = np.ascontiguousarray(a, np.int32)
        b = np.ascontiguousarray(b, np.int32)
        d = np.ascontiguousarray(d, np.float32)
        f = np.zeros((n, 3), np.float32)
        self._check(self._L.nb200_force_lennardjones(self._h, f, n, _ptr(a), _ptr(b), _ptr(d), len(a), index_base))
        return f

    def force_coulomb(self, n, a, b, d, charge, index_base=1):
        a = np.ascontiguousarray(a, np.int32)
        b = np.ascontiguousarray(b, np.int32)
        d = np.ascontiguousarray(d, np.float32)
        f = np.zeros((n, 3), np.float32)
        self._check(self._L.nb200_force_coulomb(self._h, f, n, _ptr(a), _ptr(b), _ptr(d), len(a),
                                                np.ascontiguousarray(charge, np.float32), index_base))
        return f

    def sum_forces(self, f1, f2):
        f1 = _as_f32(f1)
        f2 = _as_f32(f2)
        out = np.empty_like(f1)
        self._check(self._L.nb200_sum_forces(self._h, out.reshape(-1), f1.reshape(-1), f2.reshape(-1), f1.size))
        return out

    def verlet_update(self, pos, vel, force, force_next, mass, dt, box_min=None, box_max=None):
        pos = _as_f32(pos, (3,)).copy()
        vel = _as_f32(vel, (3,)).copy()
        bmin = None if box_min is None else np.asarray(box_min, np.float32)
        bmax = None if box_max is None else np.asarray(box_max, np.float32)
        self._check(self._L.nb200_verlet_update(self._h, pos.reshape(-1), vel.reshape(-1), _as_f32(force).reshape(-1),
                                                _as_f32(force_next).reshape(-1), np.ascontiguousarray(mass, np.float32),
                                                len(pos), np.float32(dt), _ptr(bmin), _ptr(bmax)))
        return pos, vel

    # -- MD system --
    def set_system(self, xyz, vel=None, mass=None, charge=None):
        xyz = _as_f32(xyz, (3, 4))
        n, stride = xyz.shape
        vel = None if vel is None else _as_f32(vel, (stride,))
        mass = None if mass is None else np.ascontiguousarray(mass, np.float32)
        charge = None if charge is None else np.ascontiguousarray(charge, np.float32)
        self._check(self._L.nb200_set_system(self._h, _ptr(xyz), _ptr(vel), stride, _ptr(mass), _ptr(charge), n))
        self.n = n

    def collect_objects(self, n: int, seed: int, minmass: float, maxmass: float, mincharge: float, maxcharge: float,
                        temperature: float, randomvelocity: bool, minimumdistance: float, max_rounds: int = 0):
        """collect_objects drawn on the device (box = set_box); the system is resident as after set_system.
        Returns (mass, charge, rounds, redrawn)."""
        mass = np.empty(n, np.float32)
        charge = np.empty(n, np.float32)
        rounds, redrawn = C.c_int32(), C.c_int64()
        self._check(self._L.nb200_collect_objects(self._h, n, int(seed) & 0xFFFFFFFFFFFFFFFF, np.float32(minmass), np.float32(maxmass),
                                                  np.float32(mincharge), np.float32(maxcharge), np.float32(temperature),
                                                  int(bool(randomvelocity)), np.float32(minimumdistance), int(max_rounds),
                                                  _ptr(mass), _ptr(charge), C.byref(rounds), C.byref(redrawn)))
        self.n = n
        return mass, charge, rounds.value, redrawn.value

    def step(self, nsteps: int, dt: float):
        self._check(self._L.nb200_step(self._h, nsteps, np.float32(dt)))

    def step_async(self, nsteps: int, dt: float):
        self._check(self._L.nb200_step_async(self._h, nsteps, np.float32(dt)))

    def sync(self):
        self._check(self._L.nb200_sync(self._h))

    def step_host(self, xyz, vel, nsteps: int, dt: float):
        """xyz / vel: C-contiguous float32 (n, 3|4) arrays, updated in place."""
        assert xyz.dtype == np.float32 and xyz.flags.c_contiguous
        n, stride = xyz.shape
        self._check(self._L.nb200_step_host(self._h, _ptr(xyz), _ptr(vel), stride, n, nsteps, np.float32(dt)))

    def step_host_ptr(self, xyz_ptr: int, vel_ptr, stride: int, n: int, nsteps: int, dt: float):
        self._check(self._L.nb200_step_host(self._h, xyz_ptr, vel_ptr, stride, n, nsteps, np.float32(dt)))

    def leapfrog_host_async(self, xyz_ptr: int, vel_ptr: int, stride: int, n: int, dt: float, vel_is_half_step: bool):
        """Raw host pointers (pinned memory for true asynchrony); buffers are rewritten when sync() returns."""
        self._check(self._L.nb200_leapfrog_host_async(self._h, xyz_ptr, vel_ptr, stride, n, np.float32(dt),
                                                      int(bool(vel_is_half_step))))

    def rescale_velocity(self, target_temperature: float, gamma: float, physical: bool = False):
        self._check(self._L.nb200_rescale_velocity(self._h, np.float32(target_temperature), np.float32(gamma), int(bool(physical))))

    def simulate(self, nsteps: int, dt: float, log_every: int = 1, rescale_every: int = 0, target_temperature: float = 0.0,
                 gamma: float = 0.0, out=None, out_ptr: int = 0, stride: int = 3):
        """simulate!/simulate_bvh! loop in one call; returns the (frames, n, stride) position log.  `out_ptr`: raw host
        pointer (e.g. a pinned torch tensor) of at least nsteps // log_every frames instead of a numpy array."""
        frames = nsteps // log_every if log_every > 0 else 0
        w = C.c_int64()
        if out_ptr:
            self._check(self._L.nb200_simulate(self._h, nsteps, np.float32(dt), log_every, out_ptr, stride, frames, rescale_every,
                                               np.float32(target_temperature), np.float32(gamma), C.byref(w)))
            return w.value
        if out is None:
            out = np.empty((frames, self.n, stride), np.float32)
        self._check(self._L.nb200_simulate(self._h, nsteps, np.float32(dt), log_every, _ptr(out) if frames else None, stride,
                                           len(out), rescale_every, np.float32(target_temperature), np.float32(gamma), C.byref(w)))
        return out[: w.value]

    def _get_vec(self, fn, stride=3):
        out = np.empty((self.n, stride), np.float32)
        self._check(fn(self._h, _ptr(out), stride))
        return out

    def get_positions(self, stride=3):
        return self._get_vec(self._L.nb200_get_positions, stride)

    def get_velocities(self, stride=3):
        return self._get_vec(self._L.nb200_get_velocities, stride)

    def get_forces(self, stride=3):
        return self._get_vec(self._L.nb200_get_forces, stride)

    def get_energies(self):
        ke, pe = C.c_double(), C.c_double()
        self._check(self._L.nb200_get_energies(self._h, C.byref(ke), C.byref(pe)))
        return ke.value, pe.value

    # -- multi-GPU --
    def set_stream(self, cuda_stream: int):
        self._check(self._L.nb200_set_stream(self._h, cuda_stream))

    def mg_set_owned(self, xyz, vel=None, mass=None, charge=None):
        xyz = _as_f32(xyz, (3, 4))
        n, stride = xyz.shape
        vel = None if vel is None else _as_f32(vel, (stride,))
        mass = None if mass is None else np.ascontiguousarray(mass, np.float32)
        charge = None if charge is None else np.ascontiguousarray(charge, np.float32)
        self._check(self._L.nb200_mg_set_owned(self._h, _ptr(xyz), _ptr(vel), stride, _ptr(mass), _ptr(charge), n))
        self.n_own = n

    def mg_owned_pos_device(self) -> int:
        p = C.c_void_p()
        self._check(self._L.nb200_mg_owned_pos_device(self._h, C.byref(p)))
        return p.value

    def mg_integrate(self, dt: float):
        self._check(self._L.nb200_mg_integrate(self._h, np.float32(dt)))

    def mg_search_force(self, all_pos_device=None, n_all: int = 0, own_begin: int = 0):
        """all_pos_device None: peer exchange (NVLink loads from the peers' publications)."""
        ng, nd = C.c_int64(), C.c_int64()
        self._check(self._L.nb200_mg_search_force(self._h, all_pos_device, n_all, own_begin, C.byref(ng), C.byref(nd)))
        return ng.value, nd.value

    def mg_search_force_async(self):
        self._check(self._L.nb200_mg_search_force_async(self._h))

    def mg_step_async(self, nsteps: int, dt: float):
        self._check(self._L.nb200_mg_step_async(self._h, int(nsteps), np.float32(dt)))

    def mg_leapfrog_host_async(self, xyz_ptr: int, stride: int, dt: float):
        self._check(self._L.nb200_mg_leapfrog_host_async(self._h, xyz_ptr, stride, np.float32(dt)))

    def mg_sync(self):
        ng, ne = C.c_int64(), C.c_int64()
        self._check(self._L.nb200_mg_sync(self._h, C.byref(ng), C.byref(ne)))
        return ng.value, ne.value

    def mg_publication(self):
        """(device base pointer, bytes, 64-byte CUDA IPC handle) of this rank's published region."""
        base, nbytes = C.c_void_p(), C.c_int64()
        handle = C.create_string_buffer(64)
        self._check(self._L.nb200_mg_publication(self._h, C.byref(base), C.byref(nbytes), handle))
        return base.value, nbytes.value, handle.raw

    def mg_connect(self, world: int, rank: int, own_begin, n_own, direct_base=None, ipc_handles=None):
        ob = np.ascontiguousarray(own_begin, np.int64)
        no = np.ascontiguousarray(n_own, np.int32)
        db = None
        if direct_base is not None:
            db = (C.c_void_p * world)(*[C.c_void_p(p) if p else C.c_void_p(None) for p in direct_base])
        ih = None
        if ipc_handles is not None:
            ih = C.create_string_buffer(b"".join(ipc_handles), 64 * world)
        self._check(self._L.nb200_mg_connect(self._h, world, rank, ob, no, db, ih))

    def mg_republish(self):
        self._check(self._L.nb200_mg_republish(self._h))

    def mg_owned_count(self) -> int:
        n = C.c_int32()
        self._check(self._L.nb200_mg_owned_count(self._h, C.byref(n)))
        self.n_own = n.value
        return n.value

    def mg_get_owned_ids(self):
        """global id of every row of mg_get_owned"""
        ids = np.empty(self.mg_owned_count(), np.int32)
        self._check(self._L.nb200_mg_get_owned_ids(self._h, ids))
        return ids

    def mg_set_migration(self, split, every: int):
        sp = None if split is None else np.ascontiguousarray(split, np.uint32)
        self._check(self._L.nb200_mg_set_migration(self._h, _ptr(sp), 0 if sp is None else len(sp), int(every)))

    def mg_get_owned(self, mode: int, stride: int = 3):
        out = np.empty((self.mg_owned_count(), stride), np.float32)
        self._check(self._L.nb200_mg_get_owned(self._h, _ptr(out), stride, mode))
        return out

    def mg_get_energies(self):
        ke, pe = C.c_double(), C.c_double()
        self._check(self._L.nb200_mg_get_energies(self._h, C.byref(ke), C.byref(pe)))
        return ke.value, pe.value

    def mg_get_entries(self, n_entries: int):
        a = np.empty(n_entries, np.int32)
        b = np.empty(n_entries, np.int32)
        d = np.empty(n_entries, np.float32)
        w = C.c_int64()
        self._check(self._L.nb200_mg_get_entries(self._h, _ptr(a), _ptr(b), _ptr(d), n_entries, C.byref(w)))
        return a[: w.value], b[: w.value], d[: w.value]

    # -- stage level --
    def morton30(self, xyz):
        xyz = _as_f32(xyz, (3, 4))
        keys = np.empty(xyz.shape[0], np.uint32)
        self._check(self._L.nb200_morton30(self._h, _ptr(xyz), xyz.shape[1], xyz.shape[0], keys))
        self.n = xyz.shape[0]
        return keys

    def set_curve(self, curve: int):
        self._check(self._L.nb200_set_curve(self._h, curve))

    def set_list_mode(self, mode: int):
        """1 = half list (default), 0 = directed list (include/naiveb200.h)."""
        self._check(self._L.nb200_set_list_mode(self._h, int(mode)))

    def set_fused_force(self, enable: bool):
        self._check(self._L.nb200_set_fused_force(self._h, 1 if enable else 0))

    def set_list_reuse(self, skin: float, every: int):
        self._check(self._L.nb200_set_list_reuse(self._h, np.float32(skin), int(every)))

    def set_resort_interval(self, every: int):
        self._check(self._L.nb200_set_resort_interval(self._h, int(every)))

    def sort_keys(self, xyz):
        xyz = _as_f32(xyz, (3, 4))
        keys = np.empty(xyz.shape[0], np.uint32)
        self._check(self._L.nb200_sort_keys(self._h, _ptr(xyz), xyz.shape[1], xyz.shape[0], keys))
        self.n = xyz.shape[0]
        return keys

    def sort_pairs(self, keys, vals):
        keys = np.ascontiguousarray(keys, np.uint32).copy()
        vals = np.ascontiguousarray(vals, np.uint32).copy()
        self._check(self._L.nb200_sort_pairs(self._h, keys, vals, len(keys)))
        return keys, vals

    def get_sorted_ids(self):
        ids = np.empty(self.n, np.int32)
        self._check(self._L.nb200_get_sorted_ids(self._h, ids))
        return ids

    def get_tree(self):
        nl, root = C.c_int32(), C.c_int32()
        self._check(self._L.nb200_get_tree(self._h, C.byref(nl), C.byref(root), None, None, None))
        nL = nl.value
        child = np.zeros((max(nL - 1, 0), 4), np.int32)
        nbox = np.zeros((max(nL - 1, 0), 6), np.float32)
        lbox = np.zeros((nL, 6), np.float32)
        self._check(self._L.nb200_get_tree(self._h, C.byref(nl), C.byref(root), _ptr(child) if nL > 1 else None,
                                           _ptr(nbox) if nL > 1 else None, _ptr(lbox)))
        return dict(n_leaves=nL, root=root.value, node_child=child, node_box=nbox, leaf_box=lbox)

    def get_neighbor_counts(self):
        out = np.empty(self.n, np.int32)
        self._check(self._L.nb200_get_neighbor_counts(self._h, out))
        return out

    def debug_traverse_profile(self):
        nl = (self.n + LEAF_SIZE - 1) // LEAF_SIZE
        out = np.zeros((nl, 4), np.int64)
        self._check(self._L.nb200_debug_traverse_profile(self._h, out.reshape(-1)))
        return out

    def set_profiling(self, enable, only_stage=None):
        """only_stage: name from STAGES — bracket just that stage with events (the loop is barely perturbed)."""
        level = 0 if not enable else (1 if only_stage is None else 2 + STAGES.index(only_stage))
        self._check(self._L.nb200_set_profiling(self._h, level))

    def get_stage_times(self):
        ms = np.zeros(len(STAGES), np.float64)
        launches = np.zeros(len(STAGES), np.int64)
        self._check(self._L.nb200_get_stage_times(self._h, ms, launches))
        return {s: (float(ms[i]), int(launches[i])) for i, s in enumerate(STAGES)}

    def timer_start(self):
        self._check(self._L.nb200_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self._L.nb200_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def get_stats(self):
        st = Stats()
        self._check(self._L.nb200_get_stats(self._h, C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in Stats._fields_}
