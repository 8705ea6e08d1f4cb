// atoms.cu — per-atom streaming kernels: pack/unpack, 30-bit Morton encode, the fused
// kick-drift-reflect-encode integrator, the Morton-order gather (+ leaf boxes), energies, and the
// literal reference Verlet / sum_forces entry points.  All HBM-bound: one float4 (16 B) access per
// array per atom, fully coalesced, no shared memory needed (no reuse).
#include "nb200_internal.cuh"
#include "curve.cuh"
#include "slab_grid.cuh"

namespace nb200 {

namespace {

constexpr int TPB = 256;
inline int blocks_for(int64_t n, int tpb = TPB) { return (int)((n + tpb - 1) / tpb); }

struct Box3 {
    float lo[3];
    float hi[3];
};

// ---- pack: host layout (stride 3/4 AoS) -> float4 state ----------------------------------------
__global__ void pack_kernel(const float* __restrict__ xyz, int stride, const float* __restrict__ vel,
                            const float* __restrict__ mass, const float* __restrict__ charge, int n, float4* __restrict__ pos,
                            float4* __restrict__ velo, int32_t* __restrict__ id) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = xyz + (int64_t)i * stride;
    pos[i] = make_float4(p[0], p[1], p[2], charge ? charge[i] : 0.f);
    if (velo) {
        float im = mass ? 1.0f / mass[i] : 1.0f;
        if (vel) {
            const float* v = vel + (int64_t)i * stride;
            velo[i] = make_float4(v[0], v[1], v[2], im);
        } else {
            velo[i] = make_float4(0.f, 0.f, 0.f, im);
        }
    }
    id[i] = i;
}

// refresh the xyz lanes of the sorted state from caller arrays in original order (nb200_step_host); with `keys` the
// curve keys of the refreshed positions are written in the same pass (nb200_leapfrog_host_async)
__global__ void refresh_kernel(const float* __restrict__ xyz, const float* __restrict__ vel, int stride,
                               const int32_t* __restrict__ id, int n, float4* __restrict__ pos, float4* __restrict__ velo, BoxQ q,
                               uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int id_off) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int64_t o = (int64_t)(id[s] - id_off) * stride;
    float4 p = pos[s];
    p.x = xyz[o]; p.y = xyz[o + 1]; p.z = xyz[o + 2];
    pos[s] = p;
    if (vel) {
        float4 v = velo[s];
        v.x = vel[o]; v.y = vel[o + 1]; v.z = vel[o + 2];
        velo[s] = v;
    }
    if (keys) {
        keys[s] = morton30(p.x, p.y, p.z, q);
        vals[s] = (uint32_t)s;
    }
}

__global__ void morton_kernel(const float4* __restrict__ pos, int n, BoxQ q, uint32_t* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos[i];
    keys[i] = morton30(p.x, p.y, p.z, q);
    vals[i] = (uint32_t)i;
}

// ---- fused kick + drift + wall reflection + Morton encode ---------------------------------------
// Velocity Verlet (Simulator.jl:198-222) in kick-drift-kick form.  `kick_dt` is dt/2 when the stored
// velocity is synchronised with the positions and dt when the closing half kick of the previous
// step is still pending (the two half kicks with the same force are merged into one).
// Wall handling follows boundary_reflect! (Simulator.jl:81-111): clamp to the wall, flip the
// velocity component.  48 B read + 32 B written + 8 B key/value per atom.
// order-preserving float <-> int map so atomicMin/atomicMax work on floats of either sign
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// PUBLISH (multi-GPU slab, peer_exchange.cu): the kernel also writes the box of every 32-atom publication leaf
// (pub_box[leaf][2]), accumulates the slab box (slab_box6, ordered-int min/max, initialised by the previous
// step) and resets the other parity's slab box for the next step — three launches and one 16-MB read fewer.
// PUBLISH (multi-GPU slab, peer_exchange.cu): the owned atoms are integrated in place in their resident curve order, and
// the same pass writes the rank's PUBLICATION of the step — positions and hand-over ids of the owned atoms in that
// order, the box of every 32-atom publication leaf, the slab box (ordered-int min/max; the other parity's is reset for
// the next step) — fills the ghost pre-sort slots with inert NaN placeholders (see mg_pull_kernel), and the LAST block to
// finish releases the publication flag at system scope: no separate box, fill or flag launches.
struct PublishArgs {
    float4* pub_pos;         // published positions (peer visible), this step's parity
    int32_t* pub_id;         // published hand-over ids
    const int32_t* id_in;    // hand-over id of the atom in each sorted slot
    float4* pub_box;         // boxes of the publication leaves [leaf][2]
    unsigned int* n_pub;     // header word: owned atoms in this publication
    int* slab_box6;          // slab box of this parity (initialised by the previous step)
    int* slab_box6_next;     // the other parity's slab box, reset here
    float4* g_pos;           // ghost pre-sort arrays: slots [0, g_fill) get NaN placeholders + keys
    uint32_t* g_keys;
    uint32_t* g_vals;
    int g_fill;
    unsigned int* flag;      // publication flag (peer visible): this rank's publication count, incremented here
    unsigned int* done;      // block counter for the last-block-done release (zero before and after the launch)
    // Migration on: an atom that has left this rank's Morton key range but has not been handed over yet (a STRAY — the
    // hand-over happens every k-th step) keeps the sort key it had when it was last inside.  With its new key it would
    // sort to wherever the curve leaves this rank's region, far from its neighbours: a few hundred isolated strays made
    // as many leaves with box-sized AABBs, every query leaf got them all as candidates, and the step took 40 % longer.
    const uint32_t* prev_keys;  // sorted keys of the last step, slot by slot (null: off)
    const uint32_t* split;
    int world, rank;
    BoxQ mq;                 // plain Morton quantisation (ownership), q is the sort curve's
};

template <bool PUBLISH>
__global__ void integrate_kernel(const float4* pos, float4* pos_out, float4* __restrict__ vel, const float4* __restrict__ force, int n,
                                 float kick_dt, float dt, Box3 box, BoxQ q, uint32_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals, PublishArgs pa) {
    float4* __restrict__ pub_box = pa.pub_box;
    int* __restrict__ slab_box6 = pa.slab_box6;
    int* __restrict__ slab_box6_next = pa.slab_box6_next;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float inf = __int_as_float(0x7f800000);
    float4 p = make_float4(inf, inf, inf, 0.f);
    if (i < n) {
        p = pos[i];
        float4 v = vel[i];
        float4 f = force[i];
        float k = v.w * kick_dt;  // dt/m
        v.x = fmaf(f.x, k, v.x);
        v.y = fmaf(f.y, k, v.y);
        v.z = fmaf(f.z, k, v.z);
        // drift without contraction: x + fl(v*dt) is what Simulator.jl:203 evaluates when F == 0, so the
        // force-free simulate_bvh! trajectory is reproduced bit for bit
        p.x = __fadd_rn(p.x, __fmul_rn(v.x, dt));
        p.y = __fadd_rn(p.y, __fmul_rn(v.y, dt));
        p.z = __fadd_rn(p.z, __fmul_rn(v.z, dt));
        if (p.x < box.lo[0]) { v.x = -v.x; p.x = box.lo[0]; }
        if (p.x > box.hi[0]) { v.x = -v.x; p.x = box.hi[0]; }
        if (p.y < box.lo[1]) { v.y = -v.y; p.y = box.lo[1]; }
        if (p.y > box.hi[1]) { v.y = -v.y; p.y = box.hi[1]; }
        if (p.z < box.lo[2]) { v.z = -v.z; p.z = box.lo[2]; }
        if (p.z > box.hi[2]) { v.z = -v.z; p.z = box.hi[2]; }
        pos_out[i] = p;
        vel[i] = v;
        uint32_t key = morton30(p.x, p.y, p.z, q);
        if (PUBLISH && pa.prev_keys && mg_owner(morton30(p.x, p.y, p.z, pa.mq), pa.split, pa.world) != pa.rank) key = pa.prev_keys[i];
        keys[i] = key;
        vals[i] = (uint32_t)i;
        if (PUBLISH) {
            pa.pub_pos[i] = p;
            pa.pub_id[i] = pa.id_in[i];
        }
    } else if (PUBLISH && i - n < pa.g_fill) {
        const float nan = __int_as_float(0x7fc00000);
        const int j = i - n;
        pa.g_pos[j] = make_float4(nan, nan, nan, 0.f);
        pa.g_keys[j] = morton30(nan, nan, nan, q);  // NaN quantises to cell 0
        pa.g_vals[j] = (uint32_t)j;
    }
    if (PUBLISH) {
        const unsigned full = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        float3 lo = make_float3(p.x, p.y, p.z);
        float3 hi = i < n ? lo : make_float3(-inf, -inf, -inf);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo.x = fminf(lo.x, __shfl_xor_sync(full, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(full, lo.y, o));
            lo.z = fminf(lo.z, __shfl_xor_sync(full, lo.z, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(full, hi.x, o));
            hi.y = fmaxf(hi.y, __shfl_xor_sync(full, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(full, hi.z, o));
        }
        __shared__ float s_lo[8][3], s_hi[8][3];
        if (lane == 0) {
            if ((i - lane) < n) {
                pub_box[2 * (size_t)(i >> 5)] = make_float4(lo.x, lo.y, lo.z, 0.f);
                pub_box[2 * (size_t)(i >> 5) + 1] = make_float4(hi.x, hi.y, hi.z, 0.f);
            }
            const int w = threadIdx.x >> 5;
            s_lo[w][0] = lo.x; s_lo[w][1] = lo.y; s_lo[w][2] = lo.z;
            s_hi[w][0] = hi.x; s_hi[w][1] = hi.y; s_hi[w][2] = hi.z;
        }
        __syncthreads();
        if (threadIdx.x < 6) {  // one ordered-int atomic per component per block
            const int d = threadIdx.x % 3;
            const bool is_hi = threadIdx.x >= 3;
            float v = is_hi ? -inf : inf;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v = is_hi ? fmaxf(v, s_hi[w][d]) : fminf(v, s_lo[w][d]);
            if (is_hi) atomicMax(&slab_box6[3 + d], f2ord(v));
            else atomicMin(&slab_box6[d], f2ord(v));
            if (blockIdx.x == 0) slab_box6_next[threadIdx.x] = is_hi ? (int)0x80000000 : 0x7fffffff;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) *pa.n_pub = (unsigned int)n;
        if (pa.flag) {
            // last block done: everything every block wrote (positions, boxes, slab box) is visible to the peers
            // before the counter moves
            __shared__ bool last;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) last = atomicAdd(pa.done, 1u) == gridDim.x - 1;
            __syncthreads();
            if (last && threadIdx.x == 0) {
                *pa.done = 0u;
                __threadfence_system();
                const unsigned int v = *(volatile unsigned int*)pa.flag + 1u;  // only this rank ever writes its flag
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pa.flag), "r"(v) : "memory");
            }
        }
    }
}

// ---- gather into Morton order + leaf boxes -------------------------------------------------------
// One warp == one leaf (32 consecutive sorted slots).  The permutation is near identity from the
// second step on (state is kept sorted), so the gathers are almost coalesced.
// Besides the leaf AABB the warp emits, for WIDE leaves only (AABB wider than wide_leaf_limit(cutoff)), 4 SUB-BOXES: a run of 32 Morton-consecutive atoms that crosses a
// coarse cell boundary has a huge AABB (measured: up to 42 cutoffs wide at 1M atoms, 3000+ candidate
// tiles for one query leaf), so the run is cut at its 3 largest key jumps (key[i] xor key[i+1]) and each
// piece gets its own tight box.  The traversal uses the union of the 4 boxes as the query region.
__global__ void reorder_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ keys_sorted,
                               const float4* __restrict__ pos_in, const float4* __restrict__ vel_in,
                               const int32_t* __restrict__ id_in, float4* __restrict__ pos_out, float4* __restrict__ vel_out,
                               int32_t* __restrict__ id_out, float4* __restrict__ force_zero, float4* __restrict__ leaf_lo,
                               float4* __restrict__ leaf_hi, float4* __restrict__ leaf_sub, int n, float wide_limit,
                               Housekeeping hk) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    {   // housekeeping for the kernels that follow (nb200_internal.cuh): a handful of stores per thread
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long i = s; i < hk.n_flag; i += stride) hk.node_flag[i] = -1;
        for (long long i = s; i < hk.n_counter_words; i += stride) hk.counters[i] = 0u;
        for (long long i = s; i < hk.n_hist; i += stride) hk.sort_hist[i] = 0u;
        for (long long i = s; i < hk.n_ticket; i += stride) hk.sort_ticket[i] = 0u;
        for (long long i = s; i < hk.n_status; i += stride) hk.sort_status[i] = 0u;
    }
    int lane = threadIdx.x & 31;
    bool valid = s < n;
    const float inf = __int_as_float(0x7f800000);
    float3 p3 = make_float3(0.f, 0.f, 0.f);
    uint32_t key = 0;
    if (valid) {
        // perm == nullptr: REFRESH — the atoms stay where they are (order of the last sort), only the leaf boxes
        // are recomputed from the current positions; `keys_sorted` then holds the fresh keys in that order
        uint32_t src = perm ? perm[s] : (uint32_t)s;
        float4 p = pos_in[src];
        if (perm) {
            pos_out[s] = p;
            if (vel_in) vel_out[s] = vel_in[src];
            id_out[s] = id_in ? id_in[src] : (int32_t)src;  // (no id table: the pre-sort index itself — the ghost segment)
        }
        if (force_zero) force_zero[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        p3 = make_float3(p.x, p.y, p.z);
        key = keys_sorted[s];
    }
    const unsigned full = 0xffffffffu;
    const unsigned vmask = __ballot_sync(full, valid);
    const int cnt = __popc(vmask);
    if (cnt == 0) return;  // warp-uniform
    // ---- leaf AABB ----  (NaN placeholders of the multi-GPU ghost region are not boxed: a leaf made only of them keeps
    // (+inf, -inf) and is never near anything; fminf/fmaxf alone would leave a NaN box, which every gap test passes)
    const bool boxed = valid && p3.x == p3.x;
    float3 alo = boxed ? p3 : make_float3(inf, inf, inf), ahi = boxed ? p3 : make_float3(-inf, -inf, -inf);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        alo.x = fminf(alo.x, __shfl_xor_sync(full, alo.x, o)); alo.y = fminf(alo.y, __shfl_xor_sync(full, alo.y, o));
        alo.z = fminf(alo.z, __shfl_xor_sync(full, alo.z, o)); ahi.x = fmaxf(ahi.x, __shfl_xor_sync(full, ahi.x, o));
        ahi.y = fmaxf(ahi.y, __shfl_xor_sync(full, ahi.y, o)); ahi.z = fmaxf(ahi.z, __shfl_xor_sync(full, ahi.z, o));
    }
    const uint32_t key0 = __shfl_sync(full, key, 0);
    if (lane == 0) {
        int leaf = s >> 5;
        leaf_lo[leaf] = make_float4(alo.x, alo.y, alo.z, __int_as_float(cnt));
        leaf_hi[leaf] = make_float4(ahi.x, ahi.y, ahi.z, __uint_as_float(key0));
    }
    // ---- sub-boxes: only the traversal of a WIDE leaf reads them (same test there, traverse.cu) ----
    if (!((ahi.x - alo.x > wide_limit) || (ahi.y - alo.y > wide_limit) || (ahi.z - alo.z > wide_limit))) return;  // warp-uniform
    // the 3 largest jumps between consecutive atoms of the run -> cut positions (cut c: between lane c and c+1)
    uint32_t knext = __shfl_down_sync(full, key, 1);
    bool pair_ok = (lane + 1 < cnt);
    unsigned long long jump = pair_ok ? (((unsigned long long)(key ^ knext) << 5) | (unsigned)(31 - lane)) : 0ull;
    unsigned cutmask = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        unsigned long long best = jump;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(full, best, o);
            best = other > best ? other : best;
        }
        if (best != 0ull) {  // unique winner: the lane id is folded into the value
            int wl = 31 - (int)(best & 31ull);
            cutmask |= 1u << wl;
            if (lane == wl) jump = 0ull;
        }
    }
    const int run = __popc(cutmask & ((1u << lane) - 1u));  // cuts strictly before this lane
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        bool mine = boxed && (run == r);
        float3 lo = mine ? p3 : make_float3(inf, inf, inf);
        float3 hi = mine ? p3 : make_float3(-inf, -inf, -inf);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo.x = fminf(lo.x, __shfl_xor_sync(full, lo.x, o));
            lo.y = fminf(lo.y, __shfl_xor_sync(full, lo.y, o));
            lo.z = fminf(lo.z, __shfl_xor_sync(full, lo.z, o));
            hi.x = fmaxf(hi.x, __shfl_xor_sync(full, hi.x, o));
            hi.y = fmaxf(hi.y, __shfl_xor_sync(full, hi.y, o));
            hi.z = fmaxf(hi.z, __shfl_xor_sync(full, hi.z, o));
        }
        if (lane == r) {  // empty runs keep (+inf, -inf): never "near" anything
            int leaf = s >> 5;
            leaf_sub[(size_t)leaf * 8 + 2 * r] = make_float4(lo.x, lo.y, lo.z, 0.f);
            leaf_sub[(size_t)leaf * 8 + 2 * r + 1] = make_float4(hi.x, hi.y, hi.z, 0.f);
        }
    }
}

// ---- unpack to the caller's layout, ORIGINAL atom order -------------------------------------------
// mode 0: positions, 1: velocities (+ pending half kick), 2: forces
// id == nullptr: rows in slot order; id_off is subtracted from the ids (multi-GPU: global ids -> hand-over rows)
__global__ void unpack_kernel(const float4* __restrict__ src, const int32_t* __restrict__ id, int n, int stride,
                              float* __restrict__ out, int mode, const float4* __restrict__ force, float half_dt, int id_off) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float4 v = src[s];
    if (mode == 1 && force != nullptr) {
        float4 f = force[s];
        float k = v.w * half_dt;
        v.x = fmaf(f.x, k, v.x);
        v.y = fmaf(f.y, k, v.y);
        v.z = fmaf(f.z, k, v.z);
    }
    float* o = out + (int64_t)(id ? id[s] - id_off : s) * stride;
    o[0] = v.x; o[1] = v.y; o[2] = v.z;
    if (stride == 4) o[3] = (mode == 0) ? 0.f : v.w;
}

__global__ void unsort4_kernel(const float4* __restrict__ src, const int32_t* __restrict__ id, int n, float4* __restrict__ dst, int id_off) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) dst[id[s] - id_off] = src[s];
}

// positions and (raw) velocities in one pass: the download half of nb200_leapfrog_host_async
__global__ void unpack_state_kernel(const float4* __restrict__ pos, const float4* __restrict__ vel, const int32_t* __restrict__ id, int n,
                                    int stride, float* __restrict__ out_pos, float* __restrict__ out_vel) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float4 p = pos[s], v = vel[s];
    const int64_t o = (int64_t)id[s] * stride;
    out_pos[o] = p.x; out_pos[o + 1] = p.y; out_pos[o + 2] = p.z;
    out_vel[o] = v.x; out_vel[o + 1] = v.y; out_vel[o + 2] = v.z;
    if (stride == 4) { out_pos[o + 3] = 0.f; out_vel[o + 3] = v.w; }
}

// KE = sum m v^2 / 2 (velocities synchronised with half_dt), PE = sum force.w
__global__ void energy_kernel(const float4* __restrict__ vel, const float4* __restrict__ force, int n, float half_dt,
                              double* __restrict__ out2) {
    double ke = 0.0, pe = 0.0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        float4 v = vel[s];
        float4 f = force[s];
        float k = v.w * half_dt;
        double vx = (double)v.x + (double)f.x * k, vy = (double)v.y + (double)f.y * k, vz = (double)v.z + (double)f.z * k;
        ke += 0.5 * (vx * vx + vy * vy + vz * vz) / (double)v.w;
        pe += (double)f.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ke += __shfl_xor_sync(0xffffffffu, ke, o);
        pe += __shfl_xor_sync(0xffffffffu, pe, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out2[0], ke);
        atomicAdd(&out2[1], pe);
    }
}

// ---- list reuse (Verlet skin): how far has any atom moved since the list was built? -------------------------------
// out[0]: largest squared displacement since the last violation report (float bits), out[1]: sticky flag "some atom
// moved farther than the limit while a reused list was in force"
__global__ void displacement_kernel(const float4* __restrict__ pos, const float4* __restrict__ ref, int n, float limit2,
                                    unsigned int* __restrict__ out) {
    float m = 0.f;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const float4 p = pos[s], r = ref[s];
        const float dx = p.x - r.x, dy = p.y - r.y, dz = p.z - r.z;
        m = fmaxf(m, dx * dx + dy * dy + dz * dz);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > limit2) {
        atomicMax(out, __float_as_uint(m));  // non-negative floats order like their bits
        out[1] = 1u;
    }
}

// ---- rescale_velocity! (Simulator.jl:119-144) -----------------------------------------------------------------
// pass 1: close a pending half kick (v += F/m * half_dt) and accumulate the reference's "temperature"
//   Ti = sum_i (2 / (3 N kb)) * |v_i| * m_i / 2        (kb = 1; the reference uses the SPEED, :133-137)
// or, with physical != 0, the kinetic temperature sum_i m_i v_i^2 / (3 N).  pass 2: v *= beta.
__global__ void thermo_sum_kernel(float4* __restrict__ vel, const float4* __restrict__ force, int n, float half_dt, int physical,
                                  double* __restrict__ out) {
    double acc = 0.0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        float4 v = vel[s];
        if (half_dt != 0.f) {
            const float4 f = force[s];
            const float k = v.w * half_dt;
            v.x = fmaf(f.x, k, v.x); v.y = fmaf(f.y, k, v.y); v.z = fmaf(f.z, k, v.z);
            vel[s] = v;
        }
        const double v2 = (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z;
        const double m = 1.0 / (double)v.w;
        acc += physical ? m * v2 : sqrt(v2) * m * 0.5;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

__global__ void thermo_scale_kernel(float4* __restrict__ vel, int n, const double* __restrict__ sum, double norm, float tf, float gamma) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    // beta = (1 + gamma * (Tf / Ti - 1)) ^ 0.5 (Simulator.jl:140): the base is a Float32 expression, the exponent a Float64
    // literal, so beta is a Float64 and `velocity .*= beta` multiplies in Float64 before the store rounds to Float32.
    // (Ti itself is accumulated in Float64 here and rounded once; the reference folds it in Float32 atom by atom — a
    // difference of a few ulp of Ti that depends on the atom order)
    const float ti = (float)(*sum * norm);
    const double beta = sqrt((double)__fadd_rn(1.0f, __fmul_rn(gamma, __fsub_rn(__fdiv_rn(tf, ti), 1.0f))));
    float4 v = vel[s];
    v.x = (float)((double)v.x * beta); v.y = (float)((double)v.y * beta); v.z = (float)((double)v.z * beta);
    vel[s] = v;
}

// ---- multi-GPU helpers (Morton-slab partition) --------------------------------------
__global__ void slab_box_init_kernel(int* __restrict__ box6) {
    if (threadIdx.x < 3) box6[threadIdx.x] = 0x7fffffff;       // min
    else if (threadIdx.x < 6) box6[threadIdx.x] = (int)0x80000000;  // max
}

// AABB of the owned atoms (the slab): warp reduce, then 6 atomics per warp
__global__ void slab_box_kernel(const float4* __restrict__ pos, int n, int* __restrict__ box6) {
    const float inf = __int_as_float(0x7f800000);
    float3 lo = make_float3(inf, inf, inf), hi = make_float3(-inf, -inf, -inf);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pos[i];
        lo = make_float3(fminf(lo.x, p.x), fminf(lo.y, p.y), fminf(lo.z, p.z));
        hi = make_float3(fmaxf(hi.x, p.x), fmaxf(hi.y, p.y), fmaxf(hi.z, p.z));
    }
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(full, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(full, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(full, lo.z, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(full, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(full, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(full, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&box6[0], f2ord(lo.x)); atomicMin(&box6[1], f2ord(lo.y)); atomicMin(&box6[2], f2ord(lo.z));
        atomicMax(&box6[3], f2ord(hi.x)); atomicMax(&box6[4], f2ord(hi.y)); atomicMax(&box6[5], f2ord(hi.z));
    }
}

// Ghosts from the all-gathered positions (NCCL exchange): every foreign atom within the cutoff of the slab box (and
// inside the slab's dilated occupancy grid) goes to the ghost pre-sort arrays (warp-aggregated atomic);
// gidx[g] remembers its position in the gathered array (the global handle of the atom).
__global__ void ghost_select_kernel(const float4* __restrict__ all_pos, long long n_all, long long own_begin, int n_own,
                                    const int* __restrict__ box6, float cutoff, float4* __restrict__ pos_out,
                                    int32_t* __restrict__ gidx_out, unsigned int* __restrict__ ghost_count,
                                    unsigned int ghost_capacity, const unsigned long long* __restrict__ grid, GridQ gq, BoxQ bq,
                                    uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    bool in_range = i < n_all;
    float4 p = in_range ? all_pos[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    bool own = in_range && i >= own_begin && i < own_begin + n_own;
    bool ghost = false;
    if (in_range && !own) {
        float lo[3] = {ord2f(box6[0]), ord2f(box6[1]), ord2f(box6[2])};
        float hi[3] = {ord2f(box6[3]), ord2f(box6[4]), ord2f(box6[5])};
        float gx = fmaxf(0.f, fmaxf(lo[0] - p.x, p.x - hi[0]));
        float gy = fmaxf(0.f, fmaxf(lo[1] - p.y, p.y - hi[1]));
        float gz = fmaxf(0.f, fmaxf(lo[2] - p.z, p.z - hi[2]));
        float r2 = cutoff * cutoff;
        ghost = gx * gx + gy * gy + gz * gz <= fmaf(r2, 4e-6f, r2) + 1e-37f;  // same conservative pad as the traversal
        if (ghost && grid) ghost = grid_point(grid, gq, p.x, p.y, p.z);         // and inside the slab's dilated occupancy grid
    }
    unsigned m = __ballot_sync(full, ghost);
    if (m) {
        unsigned base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(ghost_count, (unsigned)__popc(m));
        base = __shfl_sync(full, base, __ffs(m) - 1);
        if (ghost) {
            unsigned g = base + __popc(m & ((1u << lane) - 1u));
            if (g < ghost_capacity) {
                pos_out[g] = p;
                gidx_out[g] = (int32_t)i;
                keys[g] = morton30(p.x, p.y, p.z, bq);
                vals[g] = g;
            }
        }
    }
}

// gathered (global) index of the atom in every sorted slot of the two-segment search arrays: owned slots
// own_begin + hand-over id, ghost slots through the ghost's pre-sort index
__global__ void compose_kernel(const int32_t* __restrict__ id_sorted, int n_own, int ghost_base, int n, int own_begin,
                               const int32_t* __restrict__ ghost_gidx, int32_t* __restrict__ out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int v = -1;
    if (s < n_own) v = own_begin + id_sorted[s];
    else if (s >= ghost_base) v = ghost_gidx[id_sorted[s]];
    out[s] = v;
}

// ---- literal reference kernels ---------------------------------------------------------------------
// sum_forces! (Forces.jl:68-75)
__global__ void sum_forces_kernel(float* __restrict__ out, const float* __restrict__ f1, const float* __restrict__ f2,
                                  int64_t n3) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) out[i] = __fadd_rn(f1[i], f2[i]);
}

// Velocity-Verlet body + boundary_reflect! (Simulator.jl:198-223, 81-111), the reference's
// operation order, every operation a single IEEE binary32 rounding (no contraction):
//   a_t = F/m ; x = x + (v*dt + (a_t*dt^2)/2) ; a_tdt = Fnext/m ; v = v + ((a_t + a_tdt)*dt)/2
__global__ void verlet_literal_kernel(float* __restrict__ pos, float* __restrict__ vel, const float* __restrict__ f,
                                      const float* __restrict__ fnext, const float* __restrict__ mass, int n, float dt,
                                      Box3 box, int reflect) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * 3) return;
    int i = (int)(t / 3), d = (int)(t - (int64_t)i * 3);
    float m = mass[i];
    float a_t = __fdiv_rn(f[t], m);
    float dt2 = __fmul_rn(dt, dt);
    float t1 = __fmul_rn(vel[t], dt);
    float t2 = __fdiv_rn(__fmul_rn(a_t, dt2), 2.0f);
    float x = __fadd_rn(pos[t], __fadd_rn(t1, t2));
    float a_tdt = __fdiv_rn(fnext[t], m);
    float s = __fadd_rn(a_t, a_tdt);
    float v = __fadd_rn(vel[t], __fdiv_rn(__fmul_rn(s, dt), 2.0f));
    if (reflect) {
        if (box.lo[d] > x) { v = -v; x = box.lo[d]; }
        if (box.hi[d] < x) { v = -v; x = box.hi[d]; }
    }
    pos[t] = x;
    vel[t] = v;
}

}  // namespace

// =====================================================================================================
// launchers
// =====================================================================================================
int launch_pack(cudaStream_t s, const float* xyz_dev, int stride, const float* vel_dev, const float* mass_dev,
                const float* charge_dev, int n, float4* pos, float4* vel, int32_t* id) {
    pack_kernel<<<blocks_for(n), TPB, 0, s>>>(xyz_dev, stride, vel_dev, mass_dev, charge_dev, n, pos, vel, id);
    return 1;
}

int launch_refresh(cudaStream_t s, const float* xyz_dev, const float* vel_dev, int stride, const int32_t* id, int n,
                   float4* pos, float4* vel, const float* bmin, const float* bmax, int hilbert, uint32_t* keys, uint32_t* vals, int id_off) {
    BoxQ q;
    q.hilbert = hilbert;
    for (int d = 0; d < 3; ++d) { q.lo[d] = 0.f; q.scale[d] = 0.f; }
    if (keys) q = make_boxq(bmin, bmax, hilbert);
    refresh_kernel<<<blocks_for(n), TPB, 0, s>>>(xyz_dev, vel_dev, stride, id, n, pos, vel, q, keys, vals, id_off);
    return 1;
}

int launch_unpack_state(cudaStream_t s, const float4* pos, const float4* vel, const int32_t* id, int n, int stride, float* out_pos,
                        float* out_vel) {
    unpack_state_kernel<<<blocks_for(n), TPB, 0, s>>>(pos, vel, id, n, stride, out_pos, out_vel);
    return 1;
}

int launch_morton(cudaStream_t s, const float4* pos, int n, const float* bmin, const float* bmax, uint32_t* keys,
                  uint32_t* vals, int hilbert) {
    morton_kernel<<<blocks_for(n), TPB, 0, s>>>(pos, n, make_boxq(bmin, bmax, hilbert), keys, vals);
    return 1;
}

int launch_integrate(cudaStream_t s, float4* pos, float4* vel, const float4* force, int n, float kick_dt, float dt,
                     const float* bmin, const float* bmax, uint32_t* keys, uint32_t* vals, int hilbert, float4* pos_out,
                     const MgPublish* pub) {
    Box3 b;
    for (int d = 0; d < 3; ++d) { b.lo[d] = bmin[d]; b.hi[d] = bmax[d]; }
    PublishArgs pa = {};
    if (pub) {
        pa.pub_pos = pub->pub_pos; pa.pub_id = pub->pub_id; pa.id_in = pub->id_in; pa.pub_box = pub->pub_box; pa.n_pub = pub->n_pub;
        pa.slab_box6 = pub->slab_box6; pa.slab_box6_next = pub->slab_box6_next;
        pa.g_pos = pub->g_pos; pa.g_keys = pub->g_keys; pa.g_vals = pub->g_vals; pa.g_fill = pub->g_pos ? pub->g_fill : 0;
        pa.flag = pub->flag; pa.done = pub->done;
        pa.prev_keys = pub->prev_keys; pa.split = pub->split; pa.world = pub->world; pa.rank = pub->rank; pa.mq = make_boxq(bmin, bmax, 0);
        integrate_kernel<true><<<blocks_for((int64_t)n + pa.g_fill), TPB, 0, s>>>(pos, pos_out ? pos_out : pos, vel, force, n, kick_dt, dt, b,
                                                                                 make_boxq(bmin, bmax, hilbert), keys, vals, pa);
    } else {
        integrate_kernel<false><<<blocks_for(n), TPB, 0, s>>>(pos, pos_out ? pos_out : pos, vel, force, n, kick_dt, dt, b,
                                                             make_boxq(bmin, bmax, hilbert), keys, vals, pa);
    }
    return 1;
}

int launch_reorder(cudaStream_t s, const uint32_t* perm, const uint32_t* keys_sorted, const float4* pos_in,
                   const float4* vel_in, const int32_t* id_in, float4* pos_out, float4* vel_out, int32_t* id_out,
                   float4* force_zero, float4* leaf_lo, float4* leaf_hi, float4* leaf_sub, int n, float cutoff,
                   const Housekeeping* hk) {
    Housekeeping none = {};
    if (n <= 0) return 0;
    reorder_kernel<<<blocks_for(n), TPB, 0, s>>>(perm, keys_sorted, pos_in, vel_in, id_in, pos_out, vel_out, id_out,
                                                force_zero, leaf_lo, leaf_hi, leaf_sub, n, wide_leaf_limit(cutoff), hk ? *hk : none);
    return 1;
}

int launch_unpack(cudaStream_t s, const float4* src, const int32_t* id, int n, int stride, float* out_dev, int mode,
                  const float4* force, float half_dt, int id_off) {
    unpack_kernel<<<blocks_for(n), TPB, 0, s>>>(src, id, n, stride, out_dev, mode, force, half_dt, id_off);
    return 1;
}

int launch_energy(cudaStream_t s, const float4* vel, const float4* force, int n, float half_dt, double* out2) {
    cudaMemsetAsync(out2, 0, 2 * sizeof(double), s);
    int blocks = min(blocks_for(n), 148 * 8);
    energy_kernel<<<blocks, TPB, 0, s>>>(vel, force, n, half_dt, out2);
    return 1;
}

int launch_displacement_check(cudaStream_t s, const float4* pos, const float4* pos_ref, int n, float limit2, unsigned int* out2) {
    displacement_kernel<<<min(blocks_for(n), 148 * 8), TPB, 0, s>>>(pos, pos_ref, n, limit2, out2);
    return 1;
}

int launch_rescale_velocity(cudaStream_t s, float4* vel, const float4* force, int n, float half_dt, float tf, float gamma, int physical,
                            double* sum_dev) {
    cudaMemsetAsync(sum_dev, 0, sizeof(double), s);
    thermo_sum_kernel<<<min(blocks_for(n), 148 * 8), TPB, 0, s>>>(vel, force, n, half_dt, physical, sum_dev);
    // literal: Ti = (2 / (3 N)) * sum |v| m / 2; physical: T = sum m v^2 / (3 N)
    const double norm = physical ? 1.0 / (3.0 * (double)n) : 2.0 / (3.0 * (double)n);
    thermo_scale_kernel<<<blocks_for(n), TPB, 0, s>>>(vel, n, sum_dev, norm, tf, gamma);
    return 2;
}

int launch_slab_box_init(cudaStream_t s, int* box6) {
    slab_box_init_kernel<<<1, 32, 0, s>>>(box6);
    return 1;
}

int launch_slab_box(cudaStream_t s, const float4* pos, int n, int* box6, bool init) {
    if (init) slab_box_init_kernel<<<1, 32, 0, s>>>(box6);
    if (n > 0) slab_box_kernel<<<min(blocks_for(n), 148 * 4), TPB, 0, s>>>(pos, n, box6);
    return 2;
}

int launch_ghost_select(cudaStream_t s, const float4* all_pos, int64_t n_all, int64_t own_begin, int n_own, const int* box6,
                        float cutoff, float4* gpos, int32_t* ggidx, unsigned int* ghost_count, int64_t ghost_capacity,
                        const unsigned long long* grid, const float* bmin, const float* bmax, int hilbert, uint32_t* gkeys,
                        uint32_t* gvals) {
    cudaMemsetAsync(ghost_count, 0, sizeof(unsigned int), s);
    GridQ gq = grid ? make_gridq(bmin, bmax) : GridQ();
    ghost_select_kernel<<<blocks_for(n_all), TPB, 0, s>>>(all_pos, n_all, own_begin, n_own, box6, cutoff, gpos, ggidx, ghost_count,
                                                        (unsigned int)ghost_capacity, grid, gq, make_boxq(bmin, bmax, hilbert), gkeys,
                                                        gvals);
    return 1;
}

int launch_compose(cudaStream_t s, const int32_t* id_sorted, int n_own, int ghost_base, int n, int own_begin, const int32_t* ghost_gidx,
                   int32_t* out) {
    compose_kernel<<<blocks_for(n), TPB, 0, s>>>(id_sorted, n_own, ghost_base, n, own_begin, ghost_gidx, out);
    return 1;
}

int launch_unsort4(cudaStream_t s, const float4* src, const int32_t* id, int n, float4* dst, int id_off) {
    unsort4_kernel<<<blocks_for(n), TPB, 0, s>>>(src, id, n, dst, id_off);
    return 1;
}

int launch_sum_forces(cudaStream_t s, float* out, const float* f1, const float* f2, int64_t n3) {
    sum_forces_kernel<<<blocks_for(n3), TPB, 0, s>>>(out, f1, f2, n3);
    return 1;
}

int launch_verlet_literal(cudaStream_t s, float* pos, float* vel, const float* f, const float* fnext, const float* mass,
                          int n, float dt, const float* bmin, const float* bmax, int reflect) {
    Box3 b;
    for (int d = 0; d < 3; ++d) { b.lo[d] = reflect ? bmin[d] : 0.f; b.hi[d] = reflect ? bmax[d] : 0.f; }
    verlet_literal_kernel<<<blocks_for((int64_t)n * 3), TPB, 0, s>>>(pos, vel, f, fnext, mass, n, dt, b, reflect);
    return 1;
}

// Tuning (NB200_CARVEOUT): one preferred shared-memory carve-out for every kernel of the slab step, both streams.  The owned
// pass of the traversal keeps ~176 KB of shared memory per SM resident while the ghost side's kernels run beside it; an SM
// changes its L1 / shared-memory split only when it is empty.
void carveout_atoms(int pct) {
    cudaFuncSetAttribute(reorder_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(integrate_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(integrate_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

}  // namespace nb200
