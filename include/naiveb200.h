/* naiveb200.h — C ABI of libnaiveb200.so, the B200 (sm_100a) implementation of the
 * NaiveDynamics.jl per-step MD hot path:  LBVH neighbour search -> pair forces -> velocity Verlet.
 *
 * This header is the drop-in boundary.  The reference is pure Julia and has no FFI today; the
 * entry points below are what a Julia package extension (`ext/NaiveB200.jl`, shown in
 * INTEGRATION.md and shipped in naivedynamics.jl_b200/julia/) binds with `ccall` to give methods
 * to the reference's own extension stubs (src/PkgExtensions.jl:55-67) and to mirror its CPU entry
 * points.  Each declaration cites the reference interface it replaces (paths relative to the
 * reference checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are HOST pointers unless the name says `_device`;
 *   - every call returns an nb200_status (0 == ok) and never throws / exits; the message for the
 *     last failure on a handle is nb200_last_error(handle);
 *   - the library copies inputs to the GPU during the call and never retains host pointers;
 *   - a handle is used from one host thread at a time; different handles are independent;
 *   - there is NO CPU fallback: without a CUDA device nb200_create fails with NB200_ERR_CUDA.
 */
#ifndef NAIVEB200_H
#define NAIVEB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nb200_handle nb200_handle;

typedef enum nb200_status {
    NB200_OK = 0,
    NB200_ERR_BAD_ARG = 1,       /* mirrors the reference's error("...") argument checks            */
    NB200_ERR_CUDA = 2,          /* CUDA runtime failure (message has the CUDA error string)       */
    NB200_ERR_PAIR_OVERFLOW = 3, /* neighbour buffer too small inside a step loop: regrow & retry  */
    NB200_ERR_STATE = 4,         /* call sequence error (e.g. get_pairs before neighbors)          */
    NB200_ERR_CAPACITY = 5       /* caller-provided output buffer too small                        */
} nb200_status;

/* ---- lifecycle --------------------------------------------------------------------------- */

int32_t nb200_version(void);
int32_t nb200_device_count(void);

/* n_max: largest atom count the handle will see.  pair_capacity_hint: expected number of unique
 * pairs within the cutoff (0 = let the library size it on first use; it regrows when needed). */
int32_t nb200_create(int32_t device, int64_t n_max, int64_t pair_capacity_hint, nb200_handle** out);
int32_t nb200_destroy(nb200_handle* h);
const char* nb200_last_error(const nb200_handle* h); /* h may be NULL: last create() failure */

/* Simulation box.  Replaces GenericRandomCollector.minDim/maxDim (src/MDInput.jl:72-73) as read by
 * boundary_reflect! (src/Simulator.jl:81-111).  Also the Morton quantisation domain (the reference
 * hard-codes [0,1]^3, src/Neighbors/BVHTraverse.jl:184,856-858).  Default [0,1]^3. */
int32_t nb200_set_box(nb200_handle* h, const float box_min[3], const float box_max[3]);

/* ---- neighbour search -------------------------------------------------------------------- */

/* Replaces leafbuild_traverse_bvh(position, spec) / build_traverse_bvh(position, spec)
 * (src/Neighbors/BVHTraverse.jl:1416-1428) and gpubvh_neighborlist(backend, position, spec)
 * (src/PkgExtensions.jl:66, ext/NaiveKA.jl:470-558).
 *   xyz: n points, `stride` floats apart (3 = packed SVector{3,Float32}, 4 = padded);
 *   cutoff = spec.neighbor_distance.  Pair predicate is the reference's, bit for bit:
 *   fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)) < fl(cutoff*cutoff)   (BVHTraverse.jl:1026-1027,1248).
 * Builds Morton keys -> radix sort -> LBVH -> traversal on the GPU and leaves the neighbour list
 * resident in the handle.  *pair_count receives the number of unique pairs. */
int32_t nb200_neighbors(nb200_handle* h, const float* xyz, int32_t stride, int32_t n, float cutoff, int64_t* pair_count);

/* Second half of the two-call size protocol: copies the list out as the reference's
 * Vector{Tuple{Int32,Int32,Float32}} in SoA form (a[k], b[k], d[k]) (BVHTraverse.jl:1328).
 * Ids are original atom numbers + index_base (1 for Julia).  Each unordered pair appears once,
 * oriented like the reference: `a` is the atom that comes first in the reference's own sort order
 * (10-bit key of mortoncodes!, then atom id — BVHTraverse.jl:259-284,570).  d = sqrt_rn(d2).
 * List order is unspecified (the reference's is thread-count dependent, :1255-1314). */
int32_t nb200_get_pairs(nb200_handle* h, int32_t* a, int32_t* b, float* d, int64_t capacity, int32_t index_base,
                        int64_t* written);

/* ---- reference force entry points (operate on caller-provided pair lists) ----------------- */

/* Replaces force_lennardjones!(force, pairslist, position) (src/Forces.jl:15-45) literally:
 * force zeroed; force[a] .+= (24eps/d)((2sigma/d)^12 - (sigma/d)^6) on all three components,
 * eps = -1f10, sigma = 1e-4 (Float64 arithmetic, as in Julia); nothing is added to b.
 * force: n*3 floats out. */
int32_t nb200_force_lennardjones(nb200_handle* h, float* force, int32_t n, const int32_t* a, const int32_t* b,
                                 const float* d, int64_t npairs, int32_t index_base);

/* Replaces force_coulomb!(force, pairslist, charge) (src/Forces.jl:56-66) literally, including its
 * sequential order dependence (force[b] .-= force[a] uses the running force[a]).  Evaluated by one
 * GPU thread in list order — a compatibility path for small lists, not a fast one. */
int32_t nb200_force_coulomb(nb200_handle* h, float* force, int32_t n, const int32_t* a, const int32_t* b, const float* d,
                            int64_t npairs, const float* charge, int32_t index_base);

/* Replaces sum_forces!(force, force1, force2) (src/Forces.jl:68-75). n3 = number of floats. */
int32_t nb200_sum_forces(nb200_handle* h, float* force, const float* force1, const float* force2, int64_t n3);

/* Replaces the velocity-Verlet body + boundary_reflect! of simulate!/simulate_bvh!
 * (src/Simulator.jl:198-223, 81-111) for caller-provided force arrays, same operation order,
 * Float32, no FMA contraction:  x += v*dt + (F/m*dt^2)/2 ; v += ((F/m + Fnext/m)*dt)/2 ; reflect.
 * pos, vel: n*3 in/out.  box_min/box_max may be NULL to skip the reflection. */
int32_t nb200_verlet_update(nb200_handle* h, float* pos, float* vel, const float* force, const float* force_next,
                            const float* mass, int32_t n, float dt, const float box_min[3], const float box_max[3]);

/* ---- device-resident MD system (the simulate!/simulate_bvh! loop, src/Simulator.jl:154,327) - */

/* Physical pair model used by the step loop: Lennard-Jones 12-6 (eps, sigma) + Coulomb
 * (kcoul*q_i*q_j/r), both cut at `cutoff` (= the neighbour distance); shift != 0 subtracts the
 * value at the cutoff from the pair energy.  See DESIGN.md "Forces" for why the step loop cannot
 * use Forces.jl's formulas as written. */
int32_t nb200_set_forcefield(nb200_handle* h, float eps, float sigma, float kcoul, float cutoff, int32_t shift);

/* Uploads a GenericObjectCollection (src/MDInput.jl:33-46): position/velocity (n*stride floats),
 * mass, charge (n floats; NULL = 1 / 0).  Computes the forces at the initial positions. */
int32_t nb200_set_system(nb200_handle* h, const float* xyz, const float* vel, int32_t stride, const float* mass,
                         const float* charge, int32_t n);

/* nsteps of: kick-drift (+wall reflection) -> Morton -> sort -> LBVH -> traverse -> force,
 * neighbour list rebuilt every step (as simulate_bvh! does, Simulator.jl:351-376).
 * nb200_step blocks until done; nb200_step_async only enqueues, nb200_sync waits and reports
 * NB200_ERR_PAIR_OVERFLOW if any step overflowed the neighbour buffer. */
int32_t nb200_step(nb200_handle* h, int32_t nsteps, float dt);
int32_t nb200_step_async(nb200_handle* h, int32_t nsteps, float dt);
int32_t nb200_sync(nb200_handle* h);

/* One call = upload state -> nsteps -> download positions (+velocities if vel != NULL).  This is
 * the host-buffer form the reference's simulate! has (positions in, poslog entry out). */
int32_t nb200_step_host(nb200_handle* h, float* xyz, float* vel, int32_t stride, int32_t n, int32_t nsteps, float dt);

/* The same host-buffer step in LEAPFROG order, fully asynchronous — the pipelined form of the loop body
 * of simulate_bvh! (src/Simulator.jl:351-376) for callers that keep the state on the host:
 *     H2D x(t), v  ->  Morton -> sort -> LBVH -> traverse -> force F(x(t))        (ONE search per call)
 *     v(t+dt/2) = v + (F/m)*k ,  k = dt if vel_is_half_step (v is v(t-dt/2)) else dt/2 (v is v(t))
 *     x(t+dt)   = x + v(t+dt/2)*dt , wall reflection          ->  D2H x(t+dt), v(t+dt/2) into xyz / vel.
 * Everything is enqueued on the handle's stream and the call returns at once; xyz / vel must stay valid
 * (and should be pinned host memory) until nb200_sync(h), which also reports a neighbour-buffer overflow.
 * Calls on different handles overlap: the copies of one handle run under the kernels of another.
 * Iterating it reproduces the trajectory of nb200_step (velocity Verlet with merged half kicks).
 * vel == NULL: positions-only exchange — what simulate!'s poslog contract moves per step (Simulator.jl:245): the
 * positions come from and go back to the caller, the velocities stay resident on the device (at their half step
 * after the first call; vel_is_half_step is ignored).  Half the PCIe bytes per step. */
int32_t nb200_leapfrog_host_async(nb200_handle* h, float* xyz, float* vel, int32_t stride, int32_t n, float dt,
                                  int32_t vel_is_half_step);

/* Replaces rescale_velocity!(velocity, Tf, gamma, mass, objectcount) (src/Simulator.jl:119-144) on the resident
 * system: v *= (1 + gamma*(Tf/Ti - 1))^0.5.  physical == 0: Ti as the reference computes it,
 * sum_i (2/(3 N kb)) |v_i| m_i / 2 with kb = 1 (it uses the speed, not its square); physical != 0: the kinetic
 * temperature sum_i m_i v_i^2 / (3 N).  A pending half kick of the step loop is closed first. */
int32_t nb200_rescale_velocity(nb200_handle* h, float target_temperature, float gamma, int32_t physical);

/* The loop of simulate! / simulate_bvh! (src/Simulator.jl:154-256, 327-379) in ONE call: nsteps velocity-Verlet
 * steps, neighbour list rebuilt every step; every `log_every`-th step the positions (original atom order,
 * n*stride floats) become the next frame of `poslog` — push!(poslog, deepcopy(sys.position)), :245 — copied out
 * on a second stream so the transfer overlaps the following steps (pinned host memory recommended); every
 * `rescale_every`-th step (0 = never; the reference uses 10, :241) rescale_velocity! is applied with
 * (target_temperature, gamma).  log_every == 0 logs nothing.  Blocks until steps and copies are done. */
int32_t nb200_simulate(nb200_handle* h, int32_t nsteps, float dt, int32_t log_every, float* poslog, int32_t stride,
                       int64_t frame_capacity, int32_t rescale_every, float target_temperature, float gamma,
                       int64_t* frames_written);

/* Replaces collect_objects(Collector::GenericRandomCollector) (src/MDInput.jl:305-369) with its helpers
 * generate_positions (:175-190) and unique_pairs_prune / generate_pruned_positions! (:228-283): the system is DRAWN ON
 * THE DEVICE and left resident exactly as after nb200_set_system (forces of the first step included), so a large
 * system never goes through the reference's O(N^2) host loop.  Box = nb200_set_box (Collector.minDim / maxDim).
 *   mass, charge ~ Uniform(min, max); positions ~ Uniform(box) per axis (Float64 draws stored as Float32);
 *   velocity (:319-336): randomvelocity != 0: per axis veldist = rand(Float32, n) / sum, v = temperature*veldist*3*n/mass;
 *                        otherwise v = temperature/n*3*n/mass  (kb = 1; the division by the Float64 mass draw is last).
 *   minimumdistance > 0: while any two atoms are closer (pair predicate of nb200_neighbors with that cutoff), the
 *   lower-numbered atom of every such pair gets a new position — what generate_pruned_positions! is written to do (at
 *   HEAD its loop condition `tooClose < 0` never holds and the marked atoms keep Inf coordinates).  More than
 *   max_rounds rounds (<= 0: 10*n, the reference's recursion_limit) fail with NB200_ERR_STATE and the reference's
 *   message "Objects could not be placed, ...".
 * Draws: Philox4x32-10 keyed by `seed`, counter (atom, stream, round) — a pure function of the arguments, restated in
 * numpy by the tests' oracle (the reference uses Julia's unseeded global RNG, so only distributions can be compared with it).
 * mass_out / charge_out (n floats, may be NULL) receive the Float32 masses and charges; positions and velocities are
 * read with nb200_get_positions / nb200_get_velocities.  rounds / redrawn (may be NULL): re-draw rounds, atoms re-drawn. */
int32_t nb200_collect_objects(nb200_handle* h, int32_t n, uint64_t seed, float minmass, float maxmass, float mincharge,
                              float maxcharge, float temperature, int32_t randomvelocity, float minimumdistance,
                              int32_t max_rounds, float* mass_out, float* charge_out, int32_t* rounds, int64_t* redrawn);

/* Downloads in ORIGINAL atom order.  Velocities are synchronised to the positions' time. */
int32_t nb200_get_positions(nb200_handle* h, float* xyz, int32_t stride);
int32_t nb200_get_velocities(nb200_handle* h, float* vel, int32_t stride);
int32_t nb200_get_forces(nb200_handle* h, float* force, int32_t stride);
int32_t nb200_get_energies(nb200_handle* h, double* kinetic, double* potential);
/* Unique pairs in the list the last step / search built. */
int32_t nb200_pair_count(nb200_handle* h, int64_t* pair_count);

/* ---- multi-GPU: Morton-slab partition, one process per GPU --------------
 * The reference has no multi-process code.  Each rank owns a contiguous slab of the global
 * Morton order and calls, once per MD step,
 *     nb200_mg_integrate  ->  nb200_mg_search_force
 * The halo exchange between the two is done by the library's own kernel with loads from the peers' GPU memory
 * over NVLink (peer exchange, the default: no collective, no host-side barrier), or — when the driver passes an
 * all-gathered position array — from that array (NCCL exchange, naivedynamics.jl_b200/multigpu.py).
 * Owned atoms keep the order they were handed over in; ghosts live only inside one search. */

/* Run all work of this handle on the caller's CUDA stream (e.g. torch's current stream) so the
 * library's kernels and the driver's collectives are ordered without host synchronisation. */
int32_t nb200_set_stream(nb200_handle* h, void* cuda_stream);
/* Upload this rank's slab of the GenericObjectCollection (n_own atoms) and publish it (step 0). */
int32_t nb200_mg_set_owned(nb200_handle* h, const float* xyz, const float* vel, int32_t stride, const float* mass,
                           const float* charge, int32_t n_own);
/* DEVICE pointer to the owned positions of the CURRENT step, float4{x,y,z,charge}[n_own]: the all-gather send
 * buffer of the NCCL exchange.  Changes every step (the publication is double buffered). */
int32_t nb200_mg_owned_pos_device(nb200_handle* h, void** ptr);
/* The region this rank publishes for its peers: device base pointer, size, and a 64-byte CUDA IPC handle
 * (cudaIpcMemHandle_t) to hand to the other processes.  Any output pointer may be NULL. */
int32_t nb200_mg_publication(nb200_handle* h, void** device_base, int64_t* bytes, void* ipc_handle64);
/* Map the peers' publications.  own_begin[p] / n_own[p]: slab of rank p in the global order (world entries).
 * For each peer either direct_base[p] (a device pointer valid in this process: ranks that share a process) or
 * entry p of ipc_handles (world x 64 bytes, from nb200_mg_publication on rank p) is used.  Call after every
 * nb200_mg_set_owned of any rank. */
int32_t nb200_mg_connect(nb200_handle* h, int32_t world, int32_t rank, const int64_t* own_begin, const int32_t* n_own,
                         const void* const* direct_base, const void* ipc_handles);
/* Kick-drift(+wall reflection) of the owned atoms (velocity Verlet, Simulator.jl:198-223,81-111); the same kernel
 * publishes positions, global ids and leaf boxes and releases the step flag.  Every `every`-th call (nb200_mg_set_migration)
 * also hands the atoms that left this rank's key range to their new owners. */
int32_t nb200_mg_integrate(nb200_handle* h, float dt);
/* Ghost selection -> LBVH of the owned atoms + LBVH of the ghosts -> owned pass and ghost pass of the traversal ->
 * forces on the owned atoms (complete: no reverse force exchange; a pair of two ghosts is never generated).
 * all_pos_device == NULL: peer exchange (waits on the peers' step flags, pulls only atoms within the cutoff of
 * this slab's box, and of its occupancy grid when the slab is ragged).  Otherwise DEVICE float4[n_all], the all-gathered
 * owned positions of all ranks, this rank's atoms at [own_begin, own_begin + n_own).  n_entries: unique pairs with >= 1
 * owned atom in this rank's list (half list; a cross-slab pair is in both owners' lists). */
int32_t nb200_mg_search_force(nb200_handle* h, const void* all_pos_device, int64_t n_all, int64_t own_begin, int64_t* n_ghost,
                              int64_t* n_entries);
/* The same search for the step loop, without the host round trip that sizes the launches from the ghost count:
 * peer exchange only; every launch is sized for n_own + a ghost capacity taken from the last synchronous
 * nb200_mg_search_force (+50 %, and raised on the fly from the ghost counts of the steps that have completed — the host runs at most 16 steps ahead of the GPU), unused ghost slots hold inert NaN placeholders, and the
 * call returns as soon as the work is enqueued.  nb200_mg_sync waits and reports, for all steps since the last
 * sync, a ghost count above the capacity, a neighbour-buffer overflow or a peer that never published. */
int32_t nb200_mg_search_force_async(nb200_handle* h);
int32_t nb200_mg_sync(nb200_handle* h, int64_t* n_ghost, int64_t* n_entries);
/* nsteps x (nb200_mg_integrate + nb200_mg_search_force_async) in one call.  Every rank calls it with the same nsteps (the
 * ranks move in lockstep through the flags); finish with nb200_mg_sync.  (Environment NB200_MG_GRAPH=1 replays two
 * consecutive steps — both streams, the waits for the peers' flags included — as a CUDA graph; off by default because
 * captured nodes lose the ghost stream's priority and the step gets slower) */
int32_t nb200_mg_step_async(nb200_handle* h, int32_t nsteps, float dt);
/* The slab step with HOST buffers (positions-only exchange, leapfrog order like nb200_leapfrog_host_async): x(t) of the
 * owned atoms comes from `xyz` (hand-over order; pinned memory), is published to the peers, halo, list and forces are
 * rebuilt at x(t), the owned atoms are kicked and drifted and x(t + dt) is written back into `xyz`.  Velocities stay
 * resident.  Asynchronous: `xyz` must stay valid until nb200_mg_sync.  Every rank calls it once per step. */
int32_t nb200_mg_leapfrog_host_async(nb200_handle* h, float* xyz, int32_t stride, float dt);
/* Owned atoms back to the host, one row per owned atom.  mode 0 positions, 1 velocities, 2 forces.  Row order: the
 * hand-over order of nb200_mg_set_owned as long as no atom has migrated, afterwards the rank's current curve order —
 * nb200_mg_get_owned_ids gives the global id of every row either way, nb200_mg_owned_count the number of rows. */
int32_t nb200_mg_get_owned(nb200_handle* h, float* out, int32_t stride, int32_t mode);
int32_t nb200_mg_owned_count(nb200_handle* h, int32_t* n_own);
/* Publishes the owned atoms where they are now if the rank's publication is stale (host-buffer steps or a migration since
 * it was made).  nb200_mg_search_force does this itself; a driver stepping several slabs from ONE host thread calls it for
 * all of them before the searches, because a search waits for every peer's publication. */
int32_t nb200_mg_republish(nb200_handle* h);
/* Global id of the atom in each row: own_begin (nb200_mg_connect) + its index in its first owner's hand-over order. */
int32_t nb200_mg_get_owned_ids(nb200_handle* h, int32_t* ids);
/* MIGRATION: ownership follows the atoms.  split[0..world] are splitters of the 30-bit Morton key space of the box
 * (nb200_morton30; split[0] = 0, split[world] = 2^30): rank g owns the atoms whose key lies in [split[g], split[g+1]).
 * Every `every`-th nb200_mg_integrate each rank hands the atoms that left its range to their new owner through an outbox
 * in its published region (position, velocity, global id); the new owner appends them before that step's sort.  One host
 * round trip per migration step (the owned count changes).  every = 0 (default): atoms never change rank — the slabs
 * interpenetrate as atoms diffuse and the ghost count grows.  Same arguments on every rank, after nb200_mg_connect;
 * peer exchange only. */
int32_t nb200_mg_set_migration(nb200_handle* h, const uint32_t* split, int32_t n_split, int32_t every);
int32_t nb200_mg_get_energies(nb200_handle* h, double* kinetic, double* potential);
/* This rank's list entries as indices into the global (gathered) order, with d: a = row atom, b = partner.
 * Half list: every pair with at least one owned atom, once.  Directed list: the complete rows of the owned atoms. */
int32_t nb200_mg_get_entries(nb200_handle* h, int32_t* a, int32_t* b, float* d, int64_t capacity, int64_t* written);

/* ---- stage-level entry points (parity tests, profiling) ----------------------------------- */

/* 30-bit Morton keys of n points in the handle's box (10 bits per axis, x in bit 0). */
int32_t nb200_morton30(nb200_handle* h, const float* xyz, int32_t stride, int32_t n, uint32_t* keys);
/* Space-filling curve the atoms are sorted along: 0 = Morton (the key above), 1 = Hilbert (default): the
 * same 10-bit quantisation, coordinates passed through Skilling's axes-to-transpose before the interleave.
 * Octree cells still share key prefixes (so the LBVH is built the same way), but consecutive atoms are
 * always spatial neighbours, which halves the candidate leaves per query.  The pair set does not depend on it. */
int32_t nb200_set_curve(nb200_handle* h, int32_t curve);
/* Form of the neighbour list the traversal emits and the force evaluation consumes.  The list is stored as
 * cluster-pair TILES: 32 query atoms (one leaf) x 32 gathered target atoms, the target slots plus one 32-bit hit mask
 * per query atom (256 bytes per tile), never expanded into one entry per pair.
 *   NB200_LIST_HALF (default): each unordered pair once, in the row of its Morton-earlier atom — the
 *     reference's own rule (a query leaf walks only the Morton-later part of the tree,
 *     BVHTraverse.jl:1267-1309); the reaction on the partner is accumulated by the partner's own lane from the
 *     transposed masks and added with one vector reduction per target and tile, so per-atom force sums are
 *     accumulated in a run-dependent order (differences at the 1e-7 level).
 *   NB200_LIST_DIRECTED: each pair in the row of either atom; owner-computes forces, bit-reproducible
 *     sums, twice the traversal work.  (The multi-GPU path uses the half list by default as well: a rank emits
 *     every pair with at least one owned atom once and discards what would land on ghosts.)
 * nb200_get_pairs returns the same unique pairs in either mode. */
enum { NB200_LIST_DIRECTED = 0, NB200_LIST_HALF = 1 };
int32_t nb200_set_list_mode(nb200_handle* h, int32_t mode);
/* Step loop only.  enable != 0 (default): the pair forces of every tile are evaluated inside the traversal kernel,
 * while the tile's targets are still in shared memory; the tile list is written all the same (energies on demand,
 * nb200_get_pairs and list reuse read it) but the step never reads it back.  enable == 0: the traversal only writes
 * the list and a separate force kernel reads it (the two paths are checked against each other). */
int32_t nb200_set_fused_force(nb200_handle* h, int32_t enable);
/* Step loop only: rebuild the neighbour list every `every`-th step (default 1 = every step, as simulate_bvh! does) and
 * reuse it in between — the list is then built with cutoff + skin (Verlet skin) and the force kernel re-applies the
 * reference's exact pair predicate at the force cutoff, so every step still evaluates exactly the pairs a fresh
 * search would find, provided no atom moved more than skin/2 since the list was built.  That condition is checked
 * on the device every step; nb200_sync / nb200_step fail with NB200_ERR_STATE when it was violated.  Adjacent
 * component of the hot path (list reuse across steps); the bench's headline keeps every = 1. */
int32_t nb200_set_list_reuse(nb200_handle* h, float skin, int32_t every);
/* Step loop only: re-sort the atoms along the curve every `every`-th step (default 1 = every step, the reference's
 * simulate_bvh! shape).  On the steps in between the atoms keep their order, the leaf boxes are recomputed from the
 * current positions and the tree is rebuilt over them — the update path the reference sketches with TreeData!
 * (src/Neighbors/BVHTraverse.jl:601-655).  The neighbour list is rebuilt from scratch on EVERY step either way and
 * the pair set stays exact; only the tree is a few steps "older". */
int32_t nb200_set_resort_interval(nb200_handle* h, int32_t every);
/* The 30-bit keys the pipeline actually sorts by (current curve). */
int32_t nb200_sort_keys(nb200_handle* h, const float* xyz, int32_t stride, int32_t n, uint32_t* keys);
/* Stable LSD radix sort of (key, value) pairs, in place on host arrays. */
int32_t nb200_sort_pairs(nb200_handle* h, uint32_t* keys, uint32_t* vals, int64_t n);
/* After a search/step: original atom id (0-based) held by each sorted slot. */
int32_t nb200_get_sorted_ids(nb200_handle* h, int32_t* ids);
/* After a search/step: the LBVH over leaves of NB200_LEAF_SIZE consecutive sorted atoms.
 * node_child: n_internal*4 ints  {left, right, range_first, range_last}; child >= 0 internal,
 * child < 0 leaf ~child.  boxes are [min xyz, max xyz] = 6 floats.  Any pointer may be NULL. */
int32_t nb200_get_tree(nb200_handle* h, int32_t* n_leaves, int32_t* root, int32_t* node_child, float* node_box,
                       float* leaf_box);
/* Full neighbour count of every atom (original order) in the current list. */
int32_t nb200_get_neighbor_counts(nb200_handle* h, int32_t* counts);

/* Re-runs the traversal with per-warp instrumentation.  per_leaf4: n_leaves*4 int64
 * {SM cycles, candidate leaves, tree-walk rounds, surviving targets}.  Tuning aid. */
int32_t nb200_debug_traverse_profile(nb200_handle* h, int64_t* per_leaf4);

#define NB200_LEAF_SIZE 32

enum { NB200_STAGE_INTEGRATE = 0, NB200_STAGE_MORTON, NB200_STAGE_SORT, NB200_STAGE_REORDER, NB200_STAGE_BUILD,
       NB200_STAGE_TRAVERSE, NB200_STAGE_FORCE, NB200_STAGE_EXPORT, NB200_STAGE_COUNT };

/* Per-stage CUDA-event timing on the handle's stream.  enable != 0 starts recording (and resets
 * the accumulators); enable == 1 brackets every stage, enable == 2 + s only stage s (an event pair per stage
 * costs ~5 us of GPU idle per step: six stages are 5 % of a 0.55 ms step); stage_ms / stage_launches receive NB200_STAGE_COUNT entries: summed device
 * milliseconds and kernel launches since enabling. */
int32_t nb200_set_profiling(nb200_handle* h, int32_t enable);
int32_t nb200_get_stage_times(nb200_handle* h, double* stage_ms, int64_t* stage_launches);

/* CUDA-event stopwatch on the handle's stream: start records an event, stop records a second one,
 * waits for it and returns the device time between the two (what bench.py brackets the loop with). */
int32_t nb200_timer_start(nb200_handle* h);
int32_t nb200_timer_stop(nb200_handle* h, double* elapsed_ms);

typedef struct nb200_stats {
    int64_t n_atoms;
    int64_t n_leaves;
    int64_t n_entries;        /* list entries (set mask bits): unique pairs (half list) or 2 * unique pairs (directed list) */
    int64_t n_segments;       /* tile groups (one per drain pass of a query leaf) */
    int64_t entry_capacity;
    int64_t kernel_launches;  /* kernels this handle has launched since creation */
    int64_t steps_done;
    int64_t regrows;
    int64_t n_pairs;          /* unique pairs in the current list */
    int64_t list_half;        /* 1: the current list is a half list, 0: directed */
    int64_t n_slots;          /* 4-byte words of the tile list in use = 64 x tiles (what entry_capacity bounds) */
} nb200_stats;
int32_t nb200_get_stats(nb200_handle* h, nb200_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* NAIVEB200_H */
