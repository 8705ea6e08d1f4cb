# NaiveB200.jl — the reference-side binding of libnaiveb200.so (include/naiveb200.h).
#
# Drop this file into the reference as `ext/NaiveB200.jl` (see INTEGRATION.md for the two
# Project.toml lines); it gives methods to the extension stubs NaiveDynamics already declares in
# src/PkgExtensions.jl:55-67 (`gpubvh_neighborlist`, `gpubvh_neighborlist!`, ...) and mirrors the
# CPU entry points of the hot path for a `B200Backend` argument.  Nothing here computes: every
# function packs Julia arrays into flat Float32 buffers and forwards with `ccall`.
#
# NOT EXERCISED IN THIS REPO'S CI: the build image has no Julia toolchain.  The identical ABI is
# exercised through ctypes by tests/ (naivedynamics.jl_b200/_lib.py is a line-for-line twin).
module NaiveB200

using NaiveDynamics
using NaiveDynamics: Vec3D, SpheresBVHSpecs, GenericObjectCollection, GenericRandomCollector, SimSpec
using StaticArrays

export B200Backend, b200_neighborlist, b200_simulate_bvh!, b200_force_lennardjones!, b200_force_coulomb!, b200_collect_objects

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
NaiveDynamics.gpubuild_traverse_bvh(backend::B200Backend, position::Vec3D{Float32}, spec::SpheresBVHSpecs{Float32,Int32}) =
    b200_neighborlist(position, spec; device=backend.device)

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

end # module
