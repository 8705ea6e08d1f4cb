# NaiveB200.jl — the reference-side binding of libnaiveb200.so (include/naiveb200.h).
#
# Drop this file into the reference as `ext/NaiveB200.jl` (see INTEGRATION.md for the two Project.toml lines).
# It adds METHODS ON THE REFERENCE'S OWN FUNCTION NAMES that dispatch on a `B200Backend` first argument —
#   gpubvh_neighborlist / gpubvh_neighborlist! / gpubuild_traverse_bvh /
#   gpubounding_volume_hierarchy! / gpuneighbor_traverse!                (stubs of src/PkgExtensions.jl:55-67)
#   leafbuild_traverse_bvh, build_traverse_bvh                           (BVHTraverse.jl:1416-1428)
#   force_lennardjones!, force_coulomb!, sum_forces!                     (Forces.jl:15-75)
#   boundary_reflect!, rescale_velocity!, simulate!, simulate_bvh!       (Simulator.jl:81-379)
#   collect_objects                                                      (MDInput.jl:305-369)
# — exactly how ext/NaiveKA.jl:470 hooks `gpubvh_neighborlist(backend, position, spec)`.  A caller switches to the GPU by
# passing `B200Backend()` and nothing else changes.  Nothing here computes: every function packs Julia arrays into flat
# Float32 buffers and forwards with `ccall`.  `NaiveB200.MG` binds the multi-GPU entry points (nb200_mg_*).
#
# NOT EXERCISED IN THIS REPO'S CI: the build image has no Julia toolchain.  The identical ABI is
# exercised through ctypes by tests/ (naivedynamics.jl_b200/_lib.py is a line-for-line twin).
module NaiveB200

using NaiveDynamics
using NaiveDynamics: Vec3D, SpheresBVHSpecs, GenericObjectCollection, GenericRandomCollector, SimSpec
using StaticArrays

export B200Backend

const LIB = get(ENV, "NAIVEB200_LIB", "libnaiveb200.so")

"Dispatch tag, used like the KernelAbstractions backend in gpubvh_neighborlist(backend, position, spec)."
struct B200Backend
    device::Int32
end
B200Backend() = B200Backend(Int32(0))

mutable struct Handle
    ptr::Ptr{Cvoid}
    n_max::Int64
end

function check(h::Handle, rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:nb200_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr))
    error("libnaiveb200 error $rc: $msg")   # same style as the reference's error("...") (BVHTraverse.jl:75)
end

function Handle(n_max::Integer; device::Integer=0, pair_hint::Integer=0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:nb200_create, LIB), Int32, (Int32, Int64, Int64, Ref{Ptr{Cvoid}}), device, n_max, pair_hint, out)
    if rc != 0
        msg = unsafe_string(ccall((:nb200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))
        error("libnaiveb200 error $rc: $msg")
    end
    h = Handle(out[], n_max)
    finalizer(x -> ccall((:nb200_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
    return h
end

const HANDLES = Dict{Tuple{Int32,Int64},Handle}()
function handle_for(n::Integer, device::Integer=0)
    cap = Int64(nextpow(2, max(n, 1024)))
    get!(() -> Handle(cap; device=device), HANDLES, (Int32(device), cap))
end

"Vec3D{Float32} is a Vector of pointers to heap MVectors (MDInput.jl:29): pack it to a 3xN matrix."
function pack(v::Vec3D{Float32})
    m = Matrix{Float32}(undef, 3, length(v))
    @inbounds for i in eachindex(v)
        m[1, i] = v[i][1]; m[2, i] = v[i][2]; m[3, i] = v[i][3]
    end
    return m
end
function unpack!(v::Vec3D{Float32}, m::Matrix{Float32})
    @inbounds for i in eachindex(v)
        v[i][1] = m[1, i]; v[i][2] = m[2, i]; v[i][3] = m[3, i]
    end
    return v
end

"""
    b200_neighborlist(position, spec) -> Vector{Tuple{Int32,Int32,Float32}}

Same contract as `leafbuild_traverse_bvh(position, spec)` / `build_traverse_bvh(position, spec)`
(BVHTraverse.jl:1416-1428): 1-based original atom ids, `a` = the atom that comes first in the
reference's own sort order, d = sqrt(d2), list order unspecified.
"""
function b200_neighborlist(position::Vec3D{Float32}, spec::SpheresBVHSpecs{Float32,Int32}; device=0)
    n = length(position)
    h = handle_for(n, device)
    xyz = pack(position)
    cnt = Ref{Int64}(0)
    GC.@preserve xyz check(h, ccall((:nb200_neighbors, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Int32, Int32, Float32, Ref{Int64}), h.ptr, xyz, 3, n, spec.neighbor_distance, cnt))
    np = cnt[]
    a = Vector{Int32}(undef, np); b = Vector{Int32}(undef, np); d = Vector{Float32}(undef, np)
    written = Ref{Int64}(0)
    check(h, ccall((:nb200_get_pairs, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Int64, Int32, Ref{Int64}), h.ptr, a, b, d, np, 1, written))
    return [(a[k], b[k], d[k]) for k in 1:written[]]
end

# ---- methods for the reference's own extension stubs (src/PkgExtensions.jl:66-67) ----------------------
function NaiveDynamics.gpubvh_neighborlist(backend::B200Backend, position::Vec3D{Float32}, spec::SpheresBVHSpecs{Float32,Int32})
    pairlist = b200_neighborlist(position, spec; device=backend.device)
    return (pairlist=pairlist, treedata=nothing)   # ext/NaiveKA.jl:557 returns (pairlist, treedata)
end
# (`gpubuild_traverse_bvh` is in the export list of src/PkgExtensions.jl:31 but has no `function ... end` stub at v0.0.4:
#  the method is only added when the name exists)
if isdefined(NaiveDynamics, :gpubuild_traverse_bvh)
    @eval NaiveDynamics.gpubuild_traverse_bvh(backend::B200Backend, position::Vec3D{Float32}, spec::SpheresBVHSpecs{Float32,Int32}) =
        b200_neighborlist(position, spec; device=backend.device)
end

# ---- Forces.jl entry points (literal semantics, see include/naiveb200.h) ---------------------------------
function soa(pairslist)
    a = Int32[p[1] for p in pairslist]; b = Int32[p[2] for p in pairslist]; d = Float32[p[3] for p in pairslist]
    return a, b, d
end

"force_lennardjones!(force, pairslist, position) (Forces.jl:15-45) on the GPU."
function b200_force_lennardjones!(force::Vec3D{Float32}, pairslist, position=nothing)
    n = length(force); h = handle_for(n)
    a, b, d = soa(pairslist)
    f = Matrix{Float32}(undef, 3, n)
    check(h, ccall((:nb200_force_lennardjones, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Int64, Int32), h.ptr, f, n, a, b, d, length(a), 1))
    unpack!(force, f)
    return nothing
end

"force_coulomb!(force, pairslist, charge) (Forces.jl:56-66) on the GPU (sequential list order, as in the reference)."
function b200_force_coulomb!(force::Vec3D{Float32}, pairslist, charge::Vector{Float32})
    n = length(force); h = handle_for(n)
    a, b, d = soa(pairslist)
    f = Matrix{Float32}(undef, 3, n)
    check(h, ccall((:nb200_force_coulomb, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Int64, Ptr{Float32}, Int32),
        h.ptr, f, n, a, b, d, length(a), charge, 1))
    unpack!(force, f)
    return nothing
end

# ---- simulate_bvh!(sys, spec, bvhspec, clct) (Simulator.jl:327-379) ---------------------------------------
"""
    b200_simulate_bvh!(sys, spec, bvhspec, clct; eps=0, sigma=1, kcoul=0) -> poslog

Velocity Verlet + boundary_reflect! with the neighbour list rebuilt from a fresh BVH every step, all on
the device.  With the default eps = kcoul = 0 it is the reference's loop exactly (it never computes
forces, Simulator.jl:351-376) and reproduces its trajectory bit for bit; LJ/Coulomb parameters switch
the physical pair model on.  Returns `poslog::Vector{Vec3D{Float32}}` of length duration+1 (:340,378).
"""
function b200_simulate_bvh!(sys::GenericObjectCollection{Float32}, spec::SimSpec, bvhspec::SpheresBVHSpecs{Float32,Int32},
                            clct::GenericRandomCollector{Float32}; eps=0f0, sigma=1f0, kcoul=0f0, rescale_every=0, device=0)
    n = length(sys.position); h = handle_for(n, device)
    lo = Float32[clct.minDim...]; hi = Float32[clct.maxDim...]
    check(h, ccall((:nb200_set_box, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}), h.ptr, lo, hi))
    check(h, ccall((:nb200_set_forcefield, LIB), Int32, (Ptr{Cvoid}, Float32, Float32, Float32, Float32, Int32),
                   h.ptr, eps, sigma, kcoul, bvhspec.neighbor_distance, 1))
    xyz = pack(sys.position); vel = pack(sys.velocity)
    check(h, ccall((:nb200_set_system, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int32, Ptr{Float32}, Ptr{Float32}, Int32), h.ptr, xyz, vel, 3, sys.mass, sys.charge, n))
    # the whole loop is one library call: frames are copied out while the following steps run
    nsteps = Int32(spec.duration)
    frames = Array{Float32}(undef, 3, n, Int(nsteps))
    written = Ref{Int64}(0)
    check(h, ccall((:nb200_simulate, LIB), Int32,
        (Ptr{Cvoid}, Int32, Float32, Int32, Ptr{Float32}, Int32, Int64, Int32, Float32, Float32, Ref{Int64}),
        h.ptr, nsteps, Float32(spec.stepwidth), 1, frames, 3, nsteps, Int32(rescale_every), Float32(clct.temperature),
        Float32(spec.velocityDampening), written))
    poslog = [deepcopy(sys.position)]
    for k in 1:written[]
        unpack!(sys.position, frames[:, :, k])
        push!(poslog, deepcopy(sys.position))   # Simulator.jl:245
    end
    check(h, ccall((:nb200_get_velocities, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int32), h.ptr, vel, 3))
    unpack!(sys.velocity, vel)
    return poslog
end

# ---- collect_objects(Collector) (MDInput.jl:305-369) with the draws and the minimum-distance re-draw on the device ------
"""
    b200_collect_objects(Collector; seed, cutoff, eps=0, sigma=1, kcoul=0) -> GenericObjectCollection{Float32}

`collect_objects(Collector::GenericRandomCollector)` drawn on the GPU: masses, charges, velocities (MDInput.jl:319-336),
positions (:175-190) and the re-draw of atoms closer than `Collector.minimumdistance` (:228-283) through the BVH search
instead of the O(N^2) loop.  The system stays resident in the handle `b200_simulate_bvh!` uses for this atom count.
"""
function b200_collect_objects(Collector::GenericRandomCollector{Float32}; seed::Integer=rand(UInt64),
                              cutoff=max(Collector.minimumdistance, 0.03f0), eps=0f0, sigma=1f0, kcoul=0f0, max_rounds=0, device=0)
    n = Int(Collector.objectnumber); h = handle_for(n, device)
    lo = Float32[Collector.minDim...]; hi = Float32[Collector.maxDim...]
    check(h, ccall((:nb200_set_box, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}), h.ptr, lo, hi))
    check(h, ccall((:nb200_set_forcefield, LIB), Int32, (Ptr{Cvoid}, Float32, Float32, Float32, Float32, Int32),
                   h.ptr, eps, sigma, kcoul, cutoff, 1))
    mass = Vector{Float32}(undef, n); charge = Vector{Float32}(undef, n)
    rounds = Ref{Int32}(0); redrawn = Ref{Int64}(0)
    check(h, ccall((:nb200_collect_objects, LIB), Int32,
        (Ptr{Cvoid}, Int32, UInt64, Float32, Float32, Float32, Float32, Float32, Int32, Float32, Int32,
         Ptr{Float32}, Ptr{Float32}, Ref{Int32}, Ref{Int64}),
        h.ptr, n, UInt64(seed), Collector.minmass, Collector.maxmass, Collector.mincharge, Collector.maxcharge,
        Collector.temperature, Collector.randomvelocity ? 1 : 0, Collector.minimumdistance, max_rounds, mass, charge, rounds, redrawn))
    xyz = Matrix{Float32}(undef, 3, n); vel = Matrix{Float32}(undef, 3, n)
    check(h, ccall((:nb200_get_positions, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int32), h.ptr, xyz, 3))
    check(h, ccall((:nb200_get_velocities, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int32), h.ptr, vel, 3))
    vec3(m) = [MVector{3,Float32}(m[1, i], m[2, i], m[3, i]) for i in 1:n]
    return GenericObjectCollection{Float32}(fill(1, n), fill("duck", n), mass, charge, fill(0.01f0, n), [1:n;],   # MDInput.jl:342-352
                                            vec3(xyz), vec3(vel), [MVector{3,Float32}(0, 0, 0) for _ in 1:n])
end

# ---- rescale_velocity!(velocity, Tf, γ, mass, objectcount) (Simulator.jl:119-144) on the resident system ----------
"Apply the reference's velocity rescaling to the system held by the handle of `n` atoms (after b200_simulate_bvh!)."
function b200_rescale_velocity!(n::Integer, Tf::Float32, γ::Float32; physical::Bool=false, device=0)
    h = handle_for(n, device)
    check(h, ccall((:nb200_rescale_velocity, LIB), Int32, (Ptr{Cvoid}, Float32, Float32, Int32), h.ptr, Tf, γ, physical ? 1 : 0))
    return nothing
end

# ---- tuning of the step loop (defaults reproduce simulate_bvh!: everything rebuilt every step) ------------------------
"Rebuild the neighbour list only every `every`-th step, with a Verlet skin (include/naiveb200.h: nb200_set_list_reuse)."
b200_set_list_reuse!(n::Integer, skin::Float32, every::Integer; device=0) =
    check(handle_for(n, device), ccall((:nb200_set_list_reuse, LIB), Int32, (Ptr{Cvoid}, Float32, Int32), handle_for(n, device).ptr, skin, every))
"Re-sort the atoms along the curve only every `every`-th step, leaf boxes refreshed in between (TreeData!, BVHTraverse.jl:601-655)."
b200_set_resort_interval!(n::Integer, every::Integer; device=0) =
    check(handle_for(n, device), ccall((:nb200_set_resort_interval, LIB), Int32, (Ptr{Cvoid}, Int32), handle_for(n, device).ptr, every))

# =========================================================================================================================
# Methods on the reference's own names, dispatching on B200Backend (the b200_* functions above are their implementation)
# =========================================================================================================================
NaiveDynamics.leafbuild_traverse_bvh(backend::B200Backend, position::Vec3D{Float32}, spec::SpheresBVHSpecs{Float32,Int32}) =
    b200_neighborlist(position, spec; device=backend.device)                       # BVHTraverse.jl:1423-1428
NaiveDynamics.build_traverse_bvh(backend::B200Backend, position::Vec3D{Float32}, spec::SpheresBVHSpecs{Float32,Int32}) =
    b200_neighborlist(position, spec; device=backend.device)                       # BVHTraverse.jl:1416-1421

"gpubvh_neighborlist!(neighborlist, treedata, spec, backend) (ext/NaiveKA.jl:597): rebuild the list in place from the
positions the treedata carries; returns (neighborlist, treedata) like the KA method."
function NaiveDynamics.gpubvh_neighborlist!(neighborlist, treedata, spec::SpheresBVHSpecs{Float32,Int32}, backend::B200Backend)
    position = treedata isa Vec3D{Float32} ? treedata : treedata.position
    fresh = b200_neighborlist(position, spec; device=backend.device)
    resize!(neighborlist, length(fresh)); copyto!(neighborlist, fresh)
    return (neighborlist, treedata)
end

# The two KernelAbstractions kernel stubs (src/PkgExtensions.jl:59-60; ext/NaiveKA.jl:135,352 define them as
# `kernel = gpubounding_volume_hierarchy!(backend); kernel(keys, store, spec, pos; ndrange)`).  The library builds tree and list
# in ONE pass over device-resident data, so the "build" closure runs the search for the primitives it is given (ids taken
# from `pos[i].index`) and keeps the result; the "traverse" closure copies it into the caller's pre-allocated list and
# returns the number of pairs written (the KA kernel leaves the unused tail zeroed for prune_neighbors!, NaiveKA.jl:455).
const LAST_LIST = Ref{Vector{Tuple{Int32,Int32,Float32}}}(Tuple{Int32,Int32,Float32}[])
function NaiveDynamics.gpubounding_volume_hierarchy!(backend::B200Backend)
    return function (keys, store, spec::SpheresBVHSpecs{Float32,Int32}, pos; ndrange=nothing)
        position = [MVector{3,Float32}(p.position[1], p.position[2], p.position[3]) for p in pos]
        raw = b200_neighborlist(position, spec; device=backend.device)
        LAST_LIST[] = [(pos[a].index, pos[b].index, d) for (a, b, d) in raw]     # back to the caller's atom numbering
        return nothing
    end
end
function NaiveDynamics.gpuneighbor_traverse!(backend::B200Backend)
    return function (list, keys, positions, spec; ndrange=nothing)
        fresh = LAST_LIST[]
        length(fresh) <= length(list) || error("neighbor list buffer holds $(length(list)) tuples, the search found $(length(fresh))")
        copyto!(list, 1, fresh, 1, length(fresh))
        return length(fresh)
    end
end

NaiveDynamics.force_lennardjones!(::B200Backend, force::Vec3D{Float32}, pairslist, position) =
    b200_force_lennardjones!(force, pairslist, position)                           # Forces.jl:15-45
NaiveDynamics.force_coulomb!(::B200Backend, force::Vec3D{Float32}, pairslist, charge::Vector{Float32}) =
    b200_force_coulomb!(force, pairslist, charge)                                  # Forces.jl:56-66

"sum_forces!(force, force1, force2) (Forces.jl:68-75)"
function NaiveDynamics.sum_forces!(::B200Backend, force::Vec3D{Float32}, force1::Vec3D{Float32}, force2::Vec3D{Float32})
    n = length(force); h = handle_for(n)
    f = Matrix{Float32}(undef, 3, n); f1 = pack(force1); f2 = pack(force2)
    check(h, ccall((:nb200_sum_forces, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Int64), h.ptr, f, f1, f2, 3n))
    unpack!(force, f)
    return force
end

"boundary_reflect!(position, velocity, collector) (Simulator.jl:81-111): the Verlet entry point with dt = 0 is the reflection alone."
function NaiveDynamics.boundary_reflect!(backend::B200Backend, position::Vec3D{Float32}, velocity::Vec3D{Float32}, collector::GenericRandomCollector{Float32})
    n = length(position); h = handle_for(n, backend.device)
    x = pack(position); v = pack(velocity); z = zeros(Float32, 3, n); m = ones(Float32, n)
    lo = Float32[collector.minDim...]; hi = Float32[collector.maxDim...]
    check(h, ccall((:nb200_verlet_update, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Int32, Float32, Ptr{Float32}, Ptr{Float32}),
        h.ptr, x, v, z, z, m, n, 0f0, lo, hi))
    unpack!(position, x); unpack!(velocity, v)
    return nothing
end

"rescale_velocity!(velocity, Tf, γ, mass, objectcount) (Simulator.jl:119-144) for caller-supplied arrays: a private handle, so a
system another call left resident is not disturbed."
function NaiveDynamics.rescale_velocity!(backend::B200Backend, velocity::Vec3D{Float32}, Tf::Float32, γ::Float32, mass::Vector{Float32}, objectcount::Int64)
    n = length(velocity); h = Handle(max(n, 2); device=backend.device)
    x = Matrix{Float32}(undef, 3, n); for i in 1:n; x[:, i] .= (i - 0.5f0) / n; end
    v = pack(velocity)
    check(h, ccall((:nb200_set_forcefield, LIB), Int32, (Ptr{Cvoid}, Float32, Float32, Float32, Float32, Int32), h.ptr, 0f0, 1f0, 0f0, 1f-6, 1))
    check(h, ccall((:nb200_set_system, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int32, Ptr{Float32}, Ptr{Float32}, Int32), h.ptr, x, v, 3, mass, C_NULL, n))
    check(h, ccall((:nb200_rescale_velocity, LIB), Int32, (Ptr{Cvoid}, Float32, Float32, Int32), h.ptr, Tf, γ, 0))
    check(h, ccall((:nb200_get_velocities, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int32), h.ptr, v, 3))
    unpack!(velocity, v)
    finalize(h)
    return nothing
end

"simulate_bvh!(sys, spec, bvhspec, clct) (Simulator.jl:327-379) on the GPU; keyword arguments switch the physical pair model on."
NaiveDynamics.simulate_bvh!(backend::B200Backend, sys::GenericObjectCollection{Float32}, spec::SimSpec, bvhspec::SpheresBVHSpecs{Float32,Int32},
                            clct::GenericRandomCollector{Float32}; kwargs...) =
    b200_simulate_bvh!(sys, spec, bvhspec, clct; device=backend.device, kwargs...)

"simulate!(sys, spec, clct) (Simulator.jl:154-256): the reference drives this loop from an O(N^2) pair list at `spec.threshold`;
here the list comes from the BVH search at the same distance, with LJ + Coulomb forces and rescale_velocity! every 10th step (:241-243)."
function NaiveDynamics.simulate!(backend::B200Backend, sys::GenericObjectCollection{Float32}, spec::SimSpec, clct::GenericRandomCollector{Float32};
                                 eps=1f0, sigma=Float32(spec.threshold) / 2.5f0, kcoul=1f0)
    bvhspec = SpheresBVHSpecs(; neighbor_distance=Float32(spec.threshold), atom_count=length(sys.position), floattype=Float32, atomsperleaf=1)
    return b200_simulate_bvh!(sys, spec, bvhspec, clct; eps=eps, sigma=sigma, kcoul=kcoul, rescale_every=10, device=backend.device)
end

"collect_objects(Collector) (MDInput.jl:305-369) drawn on the GPU"
NaiveDynamics.collect_objects(backend::B200Backend, Collector::GenericRandomCollector{Float32}; kwargs...) =
    b200_collect_objects(Collector; device=backend.device, kwargs...)

# =========================================================================================================================
# Multi-GPU (one process per GPU; include/naiveb200.h "nb200_mg_*", DESIGN.md section 7).  The host language only moves the
# 64-byte IPC handles once at setup (MPI.Allgather below is the caller's); the step loop has no collective.
# =========================================================================================================================
module MG
using ..NaiveB200: Handle, check, LIB

"Upload this rank's slab (3 x n_own matrices, mass, charge or nothing) and publish it."
set_owned!(h::Handle, xyz::Matrix{Float32}, vel::Matrix{Float32}, mass::Vector{Float32}, charge) =
    check(h, ccall((:nb200_mg_set_owned, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int32, Ptr{Float32}, Ptr{Float32}, Int32),
                   h.ptr, xyz, vel, 3, mass, charge === nothing ? C_NULL : charge, size(xyz, 2)))

"64-byte CUDA IPC handle of this rank's published region (all-gather these, e.g. MPI.Allgather)."
function publication(h::Handle)
    ipc = zeros(UInt8, 64); base = Ref{Ptr{Cvoid}}(C_NULL); bytes = Ref{Int64}(0)
    check(h, ccall((:nb200_mg_publication, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{Int64}, Ptr{UInt8}), h.ptr, base, bytes, ipc))
    return ipc
end

"Map the peers' publications: own_begin / n_own of every rank in the global order, ipc = world x 64 bytes."
connect!(h::Handle, world::Integer, rank::Integer, own_begin::Vector{Int64}, n_own::Vector{Int32}, ipc::Vector{UInt8}) =
    check(h, ccall((:nb200_mg_connect, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{UInt8}),
                   h.ptr, world, rank, own_begin, n_own, C_NULL, ipc))

"Ownership follows the atoms: split = world + 1 Morton-key splitters, hand-over every `every`-th step (0 = never)."
set_migration!(h::Handle, split::Vector{UInt32}, every::Integer) =
    check(h, ccall((:nb200_mg_set_migration, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt32}, Int32, Int32), h.ptr, split, length(split), every))

integrate!(h::Handle, dt::Float32) = check(h, ccall((:nb200_mg_integrate, LIB), Int32, (Ptr{Cvoid}, Float32), h.ptr, dt))

"Synchronous halo + search + forces (first call of a run: sizes the ghost region and the neighbour buffer). -> (n_ghost, n_entries)"
function search_force!(h::Handle)
    ng = Ref{Int64}(0); ne = Ref{Int64}(0)
    check(h, ccall((:nb200_mg_search_force, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ref{Int64}, Ref{Int64}), h.ptr, C_NULL, 0, 0, ng, ne))
    return ng[], ne[]
end
search_force_async!(h::Handle) = check(h, ccall((:nb200_mg_search_force_async, LIB), Int32, (Ptr{Cvoid},), h.ptr))

"nsteps x (integrate + asynchronous search) without a host round trip; finish with sync!."
step_async!(h::Handle, nsteps::Integer, dt::Float32) =
    check(h, ccall((:nb200_mg_step_async, LIB), Int32, (Ptr{Cvoid}, Int32, Float32), h.ptr, nsteps, dt))

"Wait for the enqueued steps; reports ghost-capacity / list overflow / silent peer of any of them. -> (n_ghost, n_entries)"
function sync!(h::Handle)
    ng = Ref{Int64}(0); ne = Ref{Int64}(0)
    check(h, ccall((:nb200_mg_sync, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), h.ptr, ng, ne))
    return ng[], ne[]
end

"The slab step with a host buffer: x(t) of the owned atoms in, x(t+dt) out (xyz must stay alive until sync!)."
leapfrog_host_async!(h::Handle, xyz::Matrix{Float32}, dt::Float32) =
    check(h, ccall((:nb200_mg_leapfrog_host_async, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int32, Float32), h.ptr, xyz, 3, dt))

republish!(h::Handle) = check(h, ccall((:nb200_mg_republish, LIB), Int32, (Ptr{Cvoid},), h.ptr))

function owned_count(h::Handle)
    n = Ref{Int32}(0)
    check(h, ccall((:nb200_mg_owned_count, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}), h.ptr, n))
    return Int(n[])
end

"Owned atoms back to the host: mode 0 positions, 1 velocities, 2 forces (3 x n_own); rows are identified by owned_ids."
function get_owned(h::Handle, mode::Integer)
    out = Matrix{Float32}(undef, 3, owned_count(h))
    check(h, ccall((:nb200_mg_get_owned, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int32, Int32), h.ptr, out, 3, mode))
    return out
end
function owned_ids(h::Handle)
    ids = Vector{Int32}(undef, owned_count(h))
    check(h, ccall((:nb200_mg_get_owned_ids, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), h.ptr, ids))
    return ids
end
function energies(h::Handle)
    ke = Ref{Float64}(0); pe = Ref{Float64}(0)
    check(h, ccall((:nb200_mg_get_energies, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), h.ptr, ke, pe))
    return ke[], pe[]
end
"This rank's list entries with global atom ids (0-based) and d."
function entries(h::Handle, n_entries::Integer)
    a = Vector{Int32}(undef, n_entries); b = similar(a); d = Vector{Float32}(undef, n_entries); w = Ref{Int64}(0)
    check(h, ccall((:nb200_mg_get_entries, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Int64, Ref{Int64}), h.ptr, a, b, d, n_entries, w))
    return a[1:w[]], b[1:w[]], d[1:w[]]
end
end # module MG

end # module
