// pair_force.cuh — the physical pair interaction of the step loop (Lennard-Jones 12-6 + Coulomb)
// and the tile primitives shared by its two users: the force kernel that reads the tile list (forces.cu) and the
// traversal with fused forces (traverse.cu), which evaluates the same tiles while they are still in shared memory.
// Replaces the role of force_lennardjones!/force_coulomb!/sum_forces! inside simulate! (Simulator.jl:192-195,
// Forces.jl:6-75) with physical formulas over the bit-exact pair set.
#pragma once
#include "nb200_internal.cuh"

namespace nb200 {

struct FFDev {
    float sigma2, eps24, eps4, ulj_rc, kcoul, inv_rc_shift;
    float rc2;  // fl(cutoff * cutoff): CHECK variants drop listed pairs that are outside the cutoff now (skin list)
};

inline FFDev make_ffdev(const ForceField& ff) {
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
    return d;
}

#ifdef __CUDACC__

__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Scalar force factor fs and separation d = r_i - r_j of one LISTED pair (force on i = fs * d, reaction on j = -fs * d)
// and, with WITH_PE, the pair energy.
// CHECK: the list was built with a larger cutoff (Verlet skin, nb200_set_list_reuse) — keep exactly the pairs the
// search itself would keep at the force cutoff now: the reference's predicate, no contraction (traverse.cu); a pair
// outside the cutoff gets fs = u = 0.
template <bool WITH_PE, bool CHECK>
__device__ __forceinline__ void pair_eval(const float4& pi, const float4& pj, const FFDev& ff, float& fs, float& dx, float& dy,
                                          float& dz, float& u) {
    float r2;
    bool act = true;
    if (CHECK) {
        dx = __fsub_rn(pi.x, pj.x); dy = __fsub_rn(pi.y, pj.y); dz = __fsub_rn(pi.z, pj.z);
        r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        act = r2 < ff.rc2;
        r2 = act ? r2 : 1.0f;
    } else {
        dx = pi.x - pj.x; dy = pi.y - pj.y; dz = pi.z - pj.z;
        r2 = dx * dx + dy * dy + dz * dz;
    }
    // rsqrt.approx is within 2 ulp; 1/r^2 = (1/r)^2 is then within ~4 ulp (5e-7), far inside the 1e-5 budget,
    // and replaces an IEEE division plus a sqrt + division
    const float inv_r = rsqrt_fast(r2);
    const float inv_r2 = inv_r * inv_r;
    const float s2 = ff.sigma2 * inv_r2;
    const float s6 = s2 * s2 * s2;
    // 24 eps (2 s12 - s6) / r^2 = [24 eps s6 / r^2] * (2 s6 - 1)
    fs = (ff.eps24 * inv_r2) * (s6 * fmaf(2.0f, s6, -1.0f));
    u = 0.f;
    if (WITH_PE) u = ff.eps4 * (s6 * (s6 - 1.0f)) - ff.ulj_rc;
    if (ff.kcoul != 0.0f) {
        const float qq = ff.kcoul * pi.w * pj.w;
        fs = fmaf(qq * inv_r, inv_r2, fs);
        if (WITH_PE) u = fmaf(qq, inv_r - ff.inv_rc_shift, u);
    }
    if (CHECK) {
        fs = act ? fs : 0.0f;
        if (WITH_PE) u = act ? u : 0.0f;
    }
}

// 32x32 bit-matrix transpose across a warp: lane r holds row r (bit c = column c); afterwards lane r holds column r
// (bit c = old bit r of lane c).  Five block-swap steps (16, 8, 4, 2, 1), one shuffle each.
__device__ __forceinline__ unsigned transpose32(unsigned x, int lane) {
    const unsigned full = 0xffffffffu;
    unsigned m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const unsigned y = __shfl_xor_sync(full, x, j);
        // lanes without bit j keep their low half-blocks and take the partner's low half-blocks shifted up;
        // lanes with bit j keep their high half-blocks and take the partner's high half-blocks shifted down
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
        m ^= m << (j >> 1);
    }
    return x;
}

// ---- building blocks of the per-lane pair loop (traverse.cu, forces.cu) -----------------------------------------------------
// As plain CUDA C the loop body came out at 41 SASS instructions per pair: ptxas re-loaded sigma^2, 24 eps and the force
// pointer from the constant bank and re-made the constants 1 and 2.0 in EVERY iteration, and did not unswitch the loop.  With
// the constants pinned in registers, the LJ bracket as one FFMA (48 eps s6 - 24 eps), BMSK for the bit, explicit shared-memory
// addresses (target b is one LEA away) and one loop instance per (reaction, charges) combination it is 34.

// A value neither nvcc nor ptxas can see through (written to shared memory, read back volatile): it stays in a register
// instead of being re-loaded from the constant bank inside the pair loop — one issue slot per constant and iteration.
__device__ __forceinline__ unsigned long long pinned(volatile unsigned long long* slot, unsigned long long v) {
    *slot = v;
    return *slot;
}
// shared-memory loads at explicit 32-bit addresses (kept literally: no re-derivation of the address from an index)
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned lds32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// reaction force: one 16-byte vector reduction, address = base + 16 * slot
__device__ __forceinline__ void red_add4(unsigned long long base, unsigned slot, float x, float y, float z) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + 16ull * slot), "f"(x), "f"(y), "f"(z), "f"(0.f) : "memory");
}
__device__ __forceinline__ unsigned bit_at(int b) {  // 1u << b without a register for the constant 1 (BMSK)
    unsigned m;
    asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(m) : "r"(b));
    return m;
}
// The LJ 12-6 + Coulomb force factor with pinned constants: fs = [24 eps s6 (2 s6 - 1) + k qi qj / r] / r^2, s6 = (sigma^2 / r^2)^3
struct PairConsts {
    float sigma2, eps48, neps24;  // sigma^2, 48 eps, -24 eps
    float kq;                     // k * (charge of the query atom)
};
template <bool Q>
__device__ __forceinline__ float pair_fs(const PairConsts& c, float r2, float qj) {
    const float inv_r = rsqrt_fast(r2);
    const float inv_r2 = inv_r * inv_r;
    const float s2 = c.sigma2 * inv_r2;
    const float s6 = s2 * s2 * s2;
    float v = fmaf(s6, c.eps48, c.neps24) * s6;
    if (Q) v = fmaf(c.kq * qj, inv_r, v);
    return v * inv_r2;
}

__device__ __forceinline__ int top_bit(unsigned m) {  // index of the highest set bit (FLO); m != 0
    int hb;
    asm("bfind.u32 %0, %1;" : "=r"(hb) : "r"(m));
    return hb;
}

#endif  // __CUDACC__

}  // namespace nb200
