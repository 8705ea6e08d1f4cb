// System setup on the device: the draws of collect_objects / generate_positions
// (MDInput.jl:175-190, 305-336) and the re-draw of atoms that landed too close to another one
// (unique_pairs_prune / generate_pruned_positions!, MDInput.jl:228-283), so a million-atom system never exists on
// the host and never goes through the reference's O(N^2) pair loop.
//
// Random numbers: Philox4x32-10 (Salmon et al., SC'11) keyed by the caller's seed, counter = (atom, stream, round, 0).
// A counter-based generator makes every draw a pure function of (seed, atom, round): the kernels need no state, the
// result does not depend on the launch shape, and the tests restate the same draws in numpy, so the generated
// system is checked bit for bit.  (The reference draws from Julia's unseeded global RNG: there is no
// stream to reproduce, only the distributions and the arithmetic that follows the draws.)
//
//   stream 0: x = u53(w0,w1), y = u53(w2,w3)        stream 1: z = u53(w0,w1)        (round = number of re-draws so far)
//   stream 2: mass = u53(w0,w1), charge = u53(w2,w3)   stream 3: veldist x,y,z = u24(w0), u24(w1), u24(w2)   (round 0)
//
// u53 is a Float64 in [0,1) with 53 random bits (Julia's rand(Float64) grid), u24 a Float32 in [0,1) with 24
// (rand(Float32)).  rand(Uniform(a,b)) is a + (b-a)*u with (b-a) in the type of the bounds (Float32) and the
// rest in Float64 (Distributions.jl), rounded to Float32 when it is stored in the Float32 collection.
#include "nb200_internal.cuh"

namespace nb200 {
namespace {

constexpr int TPB = 256;
inline int blocks_for(int64_t n) { return (int)((n + TPB - 1) / TPB); }

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)((((unsigned long long)hi << 32) | lo) >> 11) * 0x1p-53;
}
__device__ __forceinline__ float u24(uint32_t w) { return (float)(w >> 8) * 0x1p-24f; }

// rand(Uniform(a, b)) with Float32 bounds: a + (b - a) * u, the width in Float32, the rest in Float64
__device__ __forceinline__ double uniform64(float a, float b, double u) {
    return __dadd_rn((double)a, __dmul_rn((double)__fsub_rn(b, a), u));
}

struct SetupRanges {
    float lo[3], hi[3];
    float minmass, maxmass, mincharge, maxcharge;
};

__device__ __forceinline__ void draw_position(uint32_t i, uint32_t round, uint32_t k0, uint32_t k1, const SetupRanges& g,
                                              float* __restrict__ out4) {
    const uint4 a = philox4x32_10(make_uint4(i, 0u, round, 0u), k0, k1);
    const uint4 b = philox4x32_10(make_uint4(i, 1u, round, 0u), k0, k1);
    float4 p;
    p.x = (float)uniform64(g.lo[0], g.hi[0], u53(a.x, a.y));
    p.y = (float)uniform64(g.lo[1], g.hi[1], u53(a.z, a.w));
    p.z = (float)uniform64(g.lo[2], g.hi[2], u53(b.x, b.y));
    p.w = 0.f;
    *reinterpret_cast<float4*>(out4) = p;
}

__device__ __forceinline__ double draw_mass64(uint32_t i, uint32_t k0, uint32_t k1, const SetupRanges& g, double* charge64) {
    const uint4 m = philox4x32_10(make_uint4(i, 2u, 0u, 0u), k0, k1);
    *charge64 = uniform64(g.mincharge, g.maxcharge, u53(m.z, m.w));
    return uniform64(g.minmass, g.maxmass, u53(m.x, m.y));
}

// positions (stride 4), mass, charge and the raw velocity draws (stride 4) in the staging layout nb200_set_system
// uploads into; sum3[d] accumulates the three veldist sums.  Every draw is a multiple of 2^-24 below 1, so the fp64
// sum is exact (and therefore independent of the order of the atomics) for n < 2^29.
__global__ void __launch_bounds__(TPB) setup_draw_kernel(int n, uint32_t k0, uint32_t k1, SetupRanges g, float* __restrict__ sx,
                                                         float* __restrict__ sv, float* __restrict__ sm, float* __restrict__ sq,
                                                         double* __restrict__ sum3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (i < n) {
        draw_position((uint32_t)i, 0u, k0, k1, g, sx + (int64_t)i * 4);
        double q64;
        const double m64 = draw_mass64((uint32_t)i, k0, k1, g, &q64);
        sm[i] = (float)m64;
        sq[i] = (float)q64;
        const uint4 v = philox4x32_10(make_uint4((uint32_t)i, 3u, 0u, 0u), k0, k1);
        const float4 w = make_float4(u24(v.x), u24(v.y), u24(v.z), 0.f);
        *reinterpret_cast<float4*>(sv + (int64_t)i * 4) = w;
        s0 = w.x; s1 = w.y; s2 = w.z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sum3[0], s0);
        atomicAdd(&sum3[1], s1);
        atomicAdd(&sum3[2], s2);
    }
}

// MDInput.jl:319-336.  randomvelocity: veldist ./= sum(veldist) in Float32, then
// velocity = temperature * veldist * 3 * objectcount * kb / mass — a left fold in Float32 (kb = 1) until the division
// by the Float64 mass draw, rounded to Float32 on store.  Otherwise temperature / objectcount * 3 * objectcount / mass.
// sum(veldist) is taken as the correctly rounded Float32 of the exact sum (Julia's pairwise Float32 sum is within a few ulp).
__global__ void __launch_bounds__(TPB) setup_velocity_kernel(int n, uint32_t k0, uint32_t k1, SetupRanges g, float temperature,
                                                             int randomvelocity, const double* __restrict__ sum3,
                                                             float* __restrict__ sv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double q64;
    const double m64 = draw_mass64((uint32_t)i, k0, k1, g, &q64);
    const float nf = __int2float_rn(n);
    float4 w = *reinterpret_cast<const float4*>(sv + (int64_t)i * 4);
    float c[3] = {w.x, w.y, w.z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float t;
        if (randomvelocity) {
            const float share = __fdiv_rn(c[d], (float)sum3[d]);
            t = __fmul_rn(__fmul_rn(__fmul_rn(temperature, share), 3.0f), nf);
        } else {
            t = __fmul_rn(__fmul_rn(__fdiv_rn(temperature, nf), 3.0f), nf);
        }
        c[d] = (float)__ddiv_rn((double)t, m64);
    }
    *reinterpret_cast<float4*>(sv + (int64_t)i * 4) = make_float4(c[0], c[1], c[2], 0.f);
}

// unique_pairs_prune (MDInput.jl:228-258) marks a[i] of every too-close pair (i < j): the atom with the lower id
__global__ void __launch_bounds__(TPB) prune_mark_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t np,
                                                         int32_t* __restrict__ mark) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = a[k], j = b[k];
        mark[i < j ? i : j] = 1;
    }
}

// generate_pruned_positions! (MDInput.jl:260-283): every marked atom gets a fresh position (generate_onePosition)
__global__ void __launch_bounds__(TPB) prune_redraw_kernel(int n, uint32_t round, uint32_t k0, uint32_t k1, SetupRanges g,
                                                           int32_t* __restrict__ mark, float* __restrict__ sx,
                                                           unsigned long long* __restrict__ redrawn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool hit = i < n && mark[i] != 0;
    if (hit) {
        draw_position((uint32_t)i, round, k0, k1, g, sx + (int64_t)i * 4);
        mark[i] = 0;
    }
    const unsigned int votes = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && votes) atomicAdd(redrawn, (unsigned long long)__popc(votes));
}

SetupRanges make_ranges(const float* bmin, const float* bmax, float minmass, float maxmass, float mincharge, float maxcharge) {
    SetupRanges g;
    for (int d = 0; d < 3; ++d) { g.lo[d] = bmin[d]; g.hi[d] = bmax[d]; }
    g.minmass = minmass; g.maxmass = maxmass; g.mincharge = mincharge; g.maxcharge = maxcharge;
    return g;
}

}  // namespace

int launch_setup_draw(cudaStream_t s, int n, uint64_t seed, const float* bmin, const float* bmax, float minmass, float maxmass,
                      float mincharge, float maxcharge, float temperature, int randomvelocity, float* sx, float* sv, float* sm,
                      float* sq, double* sum3) {
    const SetupRanges g = make_ranges(bmin, bmax, minmass, maxmass, mincharge, maxcharge);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    cudaMemsetAsync(sum3, 0, 3 * sizeof(double), s);
    setup_draw_kernel<<<blocks_for(n), TPB, 0, s>>>(n, k0, k1, g, sx, sv, sm, sq, sum3);
    setup_velocity_kernel<<<blocks_for(n), TPB, 0, s>>>(n, k0, k1, g, temperature, randomvelocity, sum3, sv);
    return 2;
}

int launch_prune_redraw(cudaStream_t s, int n, uint64_t seed, uint32_t round, const float* bmin, const float* bmax,
                        const int32_t* pair_a, const int32_t* pair_b, int64_t np, int32_t* mark, float* sx,
                        unsigned long long* redrawn) {
    const SetupRanges g = make_ranges(bmin, bmax, 0.f, 0.f, 0.f, 0.f);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    int64_t blocks = (np + TPB - 1) / TPB;
    if (blocks > 148 * 16) blocks = 148 * 16;
    prune_mark_kernel<<<(int)blocks, TPB, 0, s>>>(pair_a, pair_b, np, mark);
    prune_redraw_kernel<<<blocks_for(n), TPB, 0, s>>>(n, round, k0, k1, g, mark, sx, redrawn);
    return 2;
}

}  // namespace nb200
