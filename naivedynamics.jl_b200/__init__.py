"""naivedynamics.jl_b200 — B200 (sm_100a) implementation of the NaiveDynamics.jl MD hot path.

Holds only what the path needs: csrc/ (CUDA kernels + the C ABI of include/naiveb200.h, built into
libnaiveb200.so), _lib.py (ctypes binding of that ABI), api.py (host-side mirror of the reference's
interface for the path) and julia/NaiveB200.jl (the ccall shim for the reference itself).

The directory name contains a dot, so it is imported through __graft_entry__.load_package()
under the module name `naivedynamics_jl_b200`.
"""
from . import _lib  # noqa: F401
from ._lib import Handle, NB200Error, STAGES, LEAF_SIZE  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (B200Backend, SpheresBVHSpecs, PairList, leafbuild_traverse_bvh, build_traverse_bvh,  # noqa: F401
                  gpubvh_neighborlist, TreeData, force_lennardjones_, force_coulomb_, sum_forces_,
                  GenericRandomCollector, GenericObjectCollection, generate_positions, collect_objects, SimSpec,
                  ForceModel, boundary_reflect_, rescale_velocity_, simulate_bvh_, simulate_, get_handle, release_handles)
