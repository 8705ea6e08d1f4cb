// forces.cu — pair-force kernels.
//
// (1) force_kernel: the step loop's fused Lennard-Jones 12-6 + Coulomb kernel over the neighbour list
//     (half list: reaction scattered to the partner in sorted order; directed list: owner-computes):
//     one warp per list segment, lane <-> atom of the segment's leaf,
//     each lane walks its own row (rounds are located with one ballot, reads are coalesced), gathers
//     the partner position (Morton order keeps these in L1/L2), accumulates force and energy in
//     registers and adds the result to force[] in sorted order with ONE 16-B vector atomic per atom
//     per segment (a leaf normally has a single segment).  No atomics on the pair path.
//     Replaces the role of force_lennardjones!/force_coulomb!/sum_forces! inside simulate!
//     (Simulator.jl:192-195) with physical formulas — see DESIGN.md "Forces" for why the literal
//     Forces.jl expressions cannot drive an MD loop.
// (2) lj_literal / coulomb_literal: Forces.jl:6-66 reproduced as written, behind the reference's
//     own entry-point signatures.
#include "nb200_internal.cuh"

namespace nb200 {

namespace {

struct FFDev {
    float sigma2, eps24, eps4, ulj_rc, kcoul, inv_rc_shift;
    float rc2;  // fl(cutoff * cutoff): CHECK variants drop listed pairs that are outside the cutoff now (skin list)
};

// Returns the pair's scalar force factor fs and separation d = r_i - r_j (force on i = fs * d, reaction on j = -fs * d)
// and, with WITH_PE, the pair energy.
// CHECK: the list was built with a larger cutoff (Verlet skin, nb200_set_list_reuse) — keep exactly the pairs the
// search itself would keep at the force cutoff now: the reference's predicate, no contraction (traverse.cu).
template <bool WITH_PE, bool CHECK>
__device__ __forceinline__ void pair_eval(const float4& pi, const float4& pj, const FFDev& ff, bool& act, float& fs, float& dx,
                                          float& dy, float& dz, float& u) {
    float r2;
    if (CHECK) {
        dx = __fsub_rn(pi.x, pj.x); dy = __fsub_rn(pi.y, pj.y); dz = __fsub_rn(pi.z, pj.z);
        r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        act = act && r2 < ff.rc2;
    } else {
        dx = pi.x - pj.x; dy = pi.y - pj.y; dz = pi.z - pj.z;
        r2 = dx * dx + dy * dy + dz * dz;
    }
    r2 = act ? r2 : 1.0f;
    // rsqrt.approx is within 2 ulp; 1/r^2 = (1/r)^2 is then within ~4 ulp (5e-7), far inside the 1e-5 budget,
    // and replaces an IEEE division plus a sqrt + division
    float inv_r = rsqrtf(r2);
    float inv_r2 = inv_r * inv_r;
    float s2 = ff.sigma2 * inv_r2;
    float s6 = s2 * s2 * s2;
    float s12 = s6 * s6;
    fs = ff.eps24 * (2.0f * s12 - s6) * inv_r2;
    u = 0.f;
    if (WITH_PE) u = ff.eps4 * (s12 - s6) - ff.ulj_rc;
    if (ff.kcoul != 0.0f) {
        float qq = ff.kcoul * pi.w * pj.w;
        fs = fmaf(qq * inv_r, inv_r2, fs);
        if (WITH_PE) u = fmaf(qq, inv_r - ff.inv_rc_shift, u);
    }
    fs = act ? fs : 0.0f;
    if (WITH_PE) u = act ? u : 0.0f;
}

// WITH_PE = false is the step loop's variant: the potential energy is only accumulated when somebody asks
// for it (nb200_get_energies re-runs the kernel with WITH_PE = true on the same list).
// HALF = true: the list holds each pair once (row of the Morton-earlier atom); the reaction force goes to the
// partner with one 16-byte vector reduction (red.global.add.v4.f32) — partners are the next few leaves in
// sorted order, so the reductions land in L2-resident lines ("sorted-order scatter").  Each atom of a pair
// gets half of the pair energy in .w.
// 5 blocks of 256 threads per SM (<= 51 registers): measured best (4: 0.155 ms, 5: 0.153, 6: 0.225 with spills)
template <bool WITH_PE, bool HALF, bool CHECK>
__global__ void __launch_bounds__(256, 5)
    force_kernel(const SegHdr* __restrict__ segs, const int32_t* __restrict__ entries, const Counters* __restrict__ ctr,
                 unsigned int seg_capacity, const float4* __restrict__ pos, float4* __restrict__ force, int n, FFDev ff) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned nseg = min(ctr->n_segments(), seg_capacity);
    for (unsigned seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; seg < nseg; seg += nwarps) {
        const SegHdr* H = &segs[seg];
        if (H->total == 0) continue;
        const int ia = H->leaf * LEAF + lane;
        const int c = H->cnt[lane];
        const bool valid = ia < n;
        const float4 pi = valid ? pos[ia] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int maxc = __reduce_max_sync(full, c);
        const int32_t* __restrict__ row = entries + H->base + lane;
        float fx = 0.f, fy = 0.f, fz = 0.f, pe = 0.f;
        // Groups of 4 rounds.  The entry indices of group g+1 are loaded while group g's partner positions
        // are in flight, so each group exposes ONE gather latency instead of an index load followed by a
        // dependent gather.
        auto load_idx = [&](int k0, int (&j)[4], bool (&act)[4]) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                act[u] = (k0 + u) < c;
                j[u] = act[u] ? __ldg(&row[(k0 + u) * 32]) : ia;
            }
        };
        int jn[4];
        bool an[4];
        load_idx(0, jn, an);
        for (int k = 0; k < maxc; k += 4) {
            int j[4];
            bool act[4];
            float4 pj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { j[u] = jn[u]; act[u] = an[u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) pj[u] = act[u] ? __ldg(&pos[j[u]]) : pi;
            if (k + 4 < maxc) load_idx(k + 4, jn, an);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float fs, dx, dy, dz, pu;
                pair_eval<WITH_PE, CHECK>(pi, pj[u], ff, act[u], fs, dx, dy, dz, pu);
                fx = fmaf(fs, dx, fx); fy = fmaf(fs, dy, fy); fz = fmaf(fs, dz, fz);
                if (WITH_PE) pe = fmaf(0.5f, pu, pe);
                // reaction: (-fs) * d, one multiply per component with a negated operand instead of a product and a negation
                if (HALF && act[u])  // (CHECK: pair_eval cleared act[u] for a listed pair that is outside the cutoff now)
                    atomicAdd(&force[j[u]], make_float4(-fs * dx, -fs * dy, -fs * dz, WITH_PE ? 0.5f * pu : 0.f));
            }
        }
        if (valid && c > 0) atomicAdd(&force[ia], make_float4(fx, fy, fz, pe));
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

int launch_force(cudaStream_t s, int sm_count, const SegHdr* segs, const int32_t* entries, const Counters* counters,
                 int64_t seg_capacity, const float4* pos, float4* force, int n, ForceField ff, bool with_pe, bool half,
                 bool check_cutoff) {
    FFDev d;
    d.sigma2 = ff.sigma * ff.sigma;
    d.eps24 = 24.0f * ff.eps;
    d.eps4 = 4.0f * ff.eps;
    d.kcoul = ff.kcoul;
    double src2 = (double)ff.sigma * ff.sigma / ((double)ff.cutoff * ff.cutoff);
    double src6 = src2 * src2 * src2;
    d.ulj_rc = ff.shift ? (float)(4.0 * (double)ff.eps * (src6 * src6 - src6)) : 0.0f;
    d.inv_rc_shift = ff.shift ? 1.0f / ff.cutoff : 0.0f;
    d.rc2 = ff.cutoff * ff.cutoff;  // single Float32 product, like squared_radius in the traversal
    // one warp per segment in the common case (one segment per leaf); the grid-stride loop covers the rest
    const int n_leaves = (n + LEAF - 1) / LEAF;
    int blocks = (n_leaves + n_leaves / 8 + 7) / 8;
    if (blocks < sm_count) blocks = sm_count;
    typedef void (*Kern)(const SegHdr*, const int32_t*, const Counters*, unsigned int, const float4*, float4*, int, FFDev);
    static const Kern table[8] = {force_kernel<false, false, false>, force_kernel<false, false, true>, force_kernel<false, true, false>,
                                  force_kernel<false, true, true>,   force_kernel<true, false, false>, force_kernel<true, false, true>,
                                  force_kernel<true, true, false>,   force_kernel<true, true, true>};
    const Kern kern = table[(with_pe ? 4 : 0) | (half ? 2 : 0) | (check_cutoff ? 1 : 0)];
    kern<<<blocks, 256, 0, s>>>(segs, entries, counters, (unsigned int)seg_capacity, pos, force, n, d);
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
