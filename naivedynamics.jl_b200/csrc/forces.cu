// forces.cu — pair-force kernels.
//
// (1) force_tiles_kernel: the fused Lennard-Jones 12-6 + Coulomb kernel over the TILE list (nb200_internal.cuh):
//     one warp per tile group, lane <-> atom of the group's leaf.  Per tile the 32 target positions are gathered
//     once (coalesced: a block of targets comes from at most a few leaves) into shared memory; every lane walks the
//     set bits of its hit mask: force on its query atom in registers, reaction on the partner — half list — with one
//     16-byte vector reduction (red.global.add.v4.f32) into L2-resident lines ("sorted-order scatter"); in the self
//     tile the symmetric mask (mask | transposed mask) gives every atom its complete row and nothing is sent.
//     The step loop normally evaluates the same tiles inside the traversal (traverse.cu, FUSED); this kernel serves
//     energies on demand, reused (skin) lists, and the un-fused configuration.
//     Replaces the role of force_lennardjones!/force_coulomb!/sum_forces! inside simulate!
//     (Simulator.jl:192-195) with physical formulas — see DESIGN.md "Forces" for why the literal
//     Forces.jl expressions cannot drive an MD loop.
// (2) lj_literal / coulomb_literal: Forces.jl:6-66 reproduced as written, behind the reference's
//     own entry-point signatures.
#include <type_traits>

#include "nb200_internal.cuh"
#include "pair_force.cuh"

namespace nb200 {

namespace {

constexpr int FORCE_WARPS = 8;

// WITH_PE = false is the step loop's variant: the potential energy is only accumulated when somebody asks
// for it (nb200_get_energies re-runs the kernel with WITH_PE = true on the same list).  Each atom of a pair
// gets half of the pair energy in .w.
// HALF: the list holds each pair once; the reaction goes to the partner (see above).
// CHECK: skin list — the exact predicate at the force cutoff is re-applied per pair (pair_eval).
// ghost_base (multi-GPU, GHOSTS): targets in slots >= ghost_base are ghosts and receive nothing (their owners compute that
// force themselves); every query leaf is owned.
template <bool WITH_PE, bool HALF, bool CHECK, bool GHOSTS>
__global__ void __launch_bounds__(FORCE_WARPS * 32, 5)
    force_tiles_kernel(const GroupHdr* __restrict__ groups, const int32_t* __restrict__ tiles, const Counters* __restrict__ ctr,
                       unsigned int group_capacity, const float4* __restrict__ pos, float4* __restrict__ force, int n, FFDev ff,
                       int ghost_base) {
    __shared__ float4 s_t[FORCE_WARPS][32];   // targets of the current tile
    __shared__ int32_t s_i[FORCE_WARPS][32];  // and their tile words (sorted slot of the target, -1 = none)
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float4* __restrict__ tp = s_t[threadIdx.x >> 5];
    int32_t* __restrict__ ti = s_i[threadIdx.x >> 5];
    // pair-loop constants pinned in registers, shared-memory addresses of this warp's target buffers (pair_force.cuh)
    __shared__ unsigned long long s_pin[FORCE_WARPS][3];
    const unsigned tb = (unsigned)__cvta_generic_to_shared(tp), ib = (unsigned)__cvta_generic_to_shared(ti);
    PairConsts pc = {0.f, 0.f, 0.f, 0.f};
    unsigned long long fbase = 0;
    float rc2 = 0.f;
    const bool has_q = ff.kcoul != 0.0f;
    if (!WITH_PE) {
        unsigned long long* pin = s_pin[threadIdx.x >> 5];
        const unsigned long long c = pinned(&pin[0], ((unsigned long long)__float_as_uint(ff.eps24) << 32) | __float_as_uint(ff.sigma2));
        const float e24 = __uint_as_float((unsigned)(c >> 32));
        pc.sigma2 = __uint_as_float((unsigned)c);
        pc.eps48 = 2.0f * e24;
        pc.neps24 = -e24;
        fbase = pinned(&pin[1], (unsigned long long)__cvta_generic_to_global(force));
        rc2 = __uint_as_float((unsigned)pinned(&pin[2], (unsigned long long)__float_as_uint(ff.rc2)));
    }
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned ngrp = min(ctr->n_segments(), group_capacity);
    for (unsigned g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ngrp; g += nwarps) {
        const GroupHdr H = groups[g];
        const int nt = (int)(H.ntiles & ~GROUP_SELF);
        if (nt == 0) continue;
        const bool has_self = (H.ntiles & GROUP_SELF) != 0u;
        const int ia = H.leaf * LEAF + lane;
        const bool valid = ia < n;
        const bool own_i = valid && ia < ghost_base;
        const float4 pi = valid ? pos[ia] : make_float4(0.f, 0.f, 0.f, 0.f);
        float fx = 0.f, fy = 0.f, fz = 0.f, pe = 0.f;
        const int32_t* __restrict__ T = tiles + H.base_tile * TILE_WORDS;
        int tj_n = T[lane];
        unsigned mask_n = (unsigned)T[32 + lane];
        for (int k = 0; k < nt; ++k) {
            const int tjw = tj_n;
            const unsigned mask = mask_n;
            if (k + 1 < nt) {  // next tile's words in flight while this one is evaluated
                tj_n = T[(k + 1) * TILE_WORDS + lane];
                mask_n = (unsigned)T[(k + 1) * TILE_WORDS + 32 + lane];
            }
            const int tj = tjw & 0x7fffffff;
            const float4 pt = tj < n ? __ldg(&pos[tj]) : pi;
            __syncwarp(full);
            tp[lane] = pt;
            ti[lane] = tjw;
            __syncwarp(full);
            const bool self_tile = has_self && k == 0;
            unsigned mr = mask;
            if (HALF && self_tile) mr |= transpose32(mask, lane);  // complete row of my atom inside the leaf; nothing to send
            if (!own_i) mr = 0u;
            const bool react = HALF && !self_tile;
            if (!WITH_PE) {
                // the step loop's variant: the tight pair loop of pair_force.cuh, one instance per (reaction, charges)
                auto pair_loop = [&](auto react_c, auto q_c) {
                    constexpr bool REACT = decltype(react_c)::value, Q = decltype(q_c)::value;
                    while (mr) {
                        const int b = top_bit(mr);
                        mr ^= bit_at(b);
                        const float4 tq = lds128(tb + 16u * (unsigned)b);
                        float dx, dy, dz, r2;
                        bool act = true;
                        if (CHECK) {  // skin list: the reference's predicate at the force cutoff, no contraction (pair_eval)
                            dx = __fsub_rn(pi.x, tq.x); dy = __fsub_rn(pi.y, tq.y); dz = __fsub_rn(pi.z, tq.z);
                            r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                            act = r2 < rc2;
                            r2 = act ? r2 : 1.0f;
                        } else {
                            dx = pi.x - tq.x; dy = pi.y - tq.y; dz = pi.z - tq.z;
                            r2 = dx * dx + dy * dy + dz * dz;
                        }
                        float fs = pair_fs<Q>(pc, r2, tq.w);
                        if (CHECK) fs = act ? fs : 0.0f;
                        fx = fmaf(fs, dx, fx); fy = fmaf(fs, dy, fy); fz = fmaf(fs, dz, fz);
                        if (REACT) {
                            const int tw = (int)lds32(ib + 4u * (unsigned)b);
                            if (!GHOSTS || tw < ghost_base) red_add4(fbase, (unsigned)tw, -fs * dx, -fs * dy, -fs * dz);
                        }
                    }
                };
                using std::true_type;
                using std::false_type;
                pc.kq = ff.kcoul * pi.w;
                if (react) { if (has_q) pair_loop(true_type{}, true_type{}); else pair_loop(true_type{}, false_type{}); }
                else       { if (has_q) pair_loop(false_type{}, true_type{}); else pair_loop(false_type{}, false_type{}); }
            } else {
                while (mr) {  // energies on demand: the general pair_eval, half of the pair energy to either atom
                    const int b = top_bit(mr);
                    mr ^= 1u << b;
                    float fs, dx, dy, dz, u;
                    pair_eval<true, CHECK>(pi, tp[b], ff, fs, dx, dy, dz, u);
                    fx = fmaf(fs, dx, fx); fy = fmaf(fs, dy, fy); fz = fmaf(fs, dz, fz);
                    pe = fmaf(0.5f, u, pe);
                    if (react) {
                        const int tw = ti[b];
                        if (tw < ghost_base) atomicAdd(&force[(unsigned)tw], make_float4(-fs * dx, -fs * dy, -fs * dz, 0.5f * u));
                    }
                }
            }
        }
        if (own_i) atomicAdd(&force[ia], make_float4(fx, fy, fz, pe));
    }
}

// ---- literal force_lennardjones! (Forces.jl:6-45) ---------------------------------------------------------
// force[a] .+= (24*eps ./ d) .* ((2*sigma ./ d) .^ 12 .- (sigma ./ d) .^ 6), eps = -1f10 (Float32),
// sigma = 1e-4 (Float64): `24*eps/d` is a Float32 expression, the bracket and the product are Float64.
// The same scalar goes to x, y and z, so one double accumulator per atom is enough; the reference
// accumulates in Float32 pair by pair — the results agree to rounding (checked at 1e-5 relative).
__global__ void lj_literal_kernel(const int32_t* __restrict__ a, const float* __restrict__ d, int64_t np, int index_base,
                                  int n, double* __restrict__ acc) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= np) return;
    int i = a[e] - index_base;
    if (i < 0 || i >= n) return;
    const float eps = -1e10f;
    const double sigma = 0.0001;
    float dd = d[e];
    float pre = __fdiv_rn(__fmul_rn(24.0f, eps), dd);
    double br = pow((2.0 * sigma) / (double)dd, 12.0) - pow(sigma / (double)dd, 6.0);
    atomicAdd(&acc[i], (double)pre * br);
}

__global__ void lj_literal_finish_kernel(const double* __restrict__ acc, int n, float* __restrict__ force) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = (float)acc[i];
    force[3 * (int64_t)i] = v;
    force[3 * (int64_t)i + 1] = v;
    force[3 * (int64_t)i + 2] = v;
}

// ---- literal force_coulomb! (Forces.jl:46-66) ---------------------------------------------------------------
// force[a] .+= k*q_a*q_b ./ d.^2 ; force[b] .-= force[a]   — the second statement subtracts the RUNNING
// force[a], so the result is defined only for the sequential list order.  One thread walks the list in
// order (single IEEE operations, like Julia); the three components are always equal, so it keeps one
// scalar per atom in `force[3i]` and the finish kernel replicates it.
__global__ void coulomb_literal_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, const float* __restrict__ d,
                                       int64_t np, int index_base, const float* __restrict__ charge, int n,
                                       float* __restrict__ force) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int64_t e = 0; e < np; ++e) {
        int i = a[e] - index_base, j = b[e] - index_base;
        if (i < 0 || i >= n || j < 0 || j >= n) continue;
        float dd = d[e];
        float num = __fmul_rn(charge[i], charge[j]);  // k = 1
        float v = __fdiv_rn(num, __fmul_rn(dd, dd));
        float fi = __fadd_rn(force[3 * (int64_t)i], v);
        force[3 * (int64_t)i] = fi;
        // i == j cannot occur in a pair list; if it did Julia would read the updated value too
        force[3 * (int64_t)j] = __fsub_rn(force[3 * (int64_t)j], fi);
    }
}

__global__ void replicate3_kernel(float* __restrict__ force, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = force[3 * (int64_t)i];
    force[3 * (int64_t)i + 1] = v;
    force[3 * (int64_t)i + 2] = v;
}

}  // namespace

int launch_force(cudaStream_t s, int sm_count, const GroupHdr* segs, const int32_t* entries, const Counters* counters,
                 int64_t seg_capacity, const float4* pos, float4* force, int n, ForceField ff, bool with_pe, bool half,
                 bool check_cutoff, int ghost_base) {
    const FFDev d = make_ffdev(ff);
    // one warp per group in the common case (one or two groups per leaf); the grid-stride loop covers the rest
    const int n_leaves = (n + LEAF - 1) / LEAF;
    int blocks = (2 * n_leaves + FORCE_WARPS - 1) / FORCE_WARPS;
    if (blocks < sm_count) blocks = sm_count;
    typedef void (*Kern)(const GroupHdr*, const int32_t*, const Counters*, unsigned int, const float4*, float4*, int, FFDev, int);
    const bool ghosts = ghost_base != 0x7fffffff;  // multi-GPU: targets behind ghost_base receive nothing
    static const Kern table[16] = {
        force_tiles_kernel<false, false, false, false>, force_tiles_kernel<false, false, false, true>,
        force_tiles_kernel<false, false, true, false>,  force_tiles_kernel<false, false, true, true>,
        force_tiles_kernel<false, true, false, false>,  force_tiles_kernel<false, true, false, true>,
        force_tiles_kernel<false, true, true, false>,   force_tiles_kernel<false, true, true, true>,
        force_tiles_kernel<true, false, false, false>,  force_tiles_kernel<true, false, false, true>,
        force_tiles_kernel<true, false, true, false>,   force_tiles_kernel<true, false, true, true>,
        force_tiles_kernel<true, true, false, false>,   force_tiles_kernel<true, true, false, true>,
        force_tiles_kernel<true, true, true, false>,    force_tiles_kernel<true, true, true, true>};
    const Kern kern = table[(with_pe ? 8 : 0) | (half ? 4 : 0) | (check_cutoff ? 2 : 0) | (ghosts ? 1 : 0)];
    kern<<<blocks, FORCE_WARPS * 32, 0, s>>>(segs, entries, counters, (unsigned int)seg_capacity, pos, force, n, d, ghost_base);
    return 1;
}

int launch_lj_literal(cudaStream_t s, const int32_t* a, const float* d, int64_t np, int index_base, int n, double* acc,
                      float* force) {
    int launches = 0;
    cudaMemsetAsync(acc, 0, sizeof(double) * (size_t)n, s);
    if (np > 0) {
        lj_literal_kernel<<<(unsigned)((np + 255) / 256), 256, 0, s>>>(a, d, np, index_base, n, acc);
        ++launches;
    }
    lj_literal_finish_kernel<<<(n + 255) / 256, 256, 0, s>>>(acc, n, force);
    return launches + 1;
}

int launch_coulomb_literal(cudaStream_t s, const int32_t* a, const int32_t* b, const float* d, int64_t np, int index_base,
                           const float* charge, int n, float* force) {
    cudaMemsetAsync(force, 0, sizeof(float) * 3 * (size_t)n, s);
    coulomb_literal_kernel<<<1, 32, 0, s>>>(a, b, d, np, index_base, charge, n, force);
    replicate3_kernel<<<(n + 255) / 256, 256, 0, s>>>(force, n);
    return 2;
}

}  // namespace nb200
