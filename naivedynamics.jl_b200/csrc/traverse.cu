// traverse.cu — warp-cooperative cutoff traversal of the LBVH, emitting the compacted neighbour list.
//
// Replaces leafneighbor_traverse + aabb_overlap_test + one/twocluster_proximitytest!
// (BVHTraverse.jl:1236-1323, 1094-1098, 1021-1055).  Like the reference's live path it is a
// LEAF-query traversal (its own finding: leaf-vs-tree beats atom-vs-tree, devdiary.md:1415), but
// re-shaped for a warp:
//   * one warp owns one query leaf A = 32 curve-consecutive atoms, lane <-> atom;
//   * the tree walk is cooperative and starts from a precomputed frontier of the tree's first levels: up to 32
//     internal nodes are popped from a short shared-memory stack per round, every lane tests BOTH child boxes of its
//     node (they live in the 64-B parent), hit internal children are pushed, hit leaf children become candidates;
//   * the atoms of the candidate leaves are gathered three leaves at a time, tested against A's box and compacted
//     into an SoA target buffer in shared memory; then every lane tests its query atom against all buffered targets
//     with the reference's exact predicate (packed FADD2/FMUL2, hit bits funnel-shifted into a per-lane mask);
//   * the per-lane hit masks ARE the list: every 32-target block goes out as one 256-byte TILE (target slots + one
//     mask word per query atom, layout in nb200_internal.cuh) with two coalesced 128-byte stores; one packed
//     atomicAdd per drain pass reserves the tiles of that pass.  Nothing is expanded per hit;
//   * FUSED: the pair forces of a tile are evaluated right here, while its targets are still in shared memory
//     (pair_force.cuh) — the list is still written (energies, export and list reuse read it) but the step loop
//     never reads it back;
//   * two list forms: HALF (default; each pair once, in the row of its curve-earlier atom, like the reference's
//     traversal) and DIRECTED (each pair in both rows; deterministic force sums).
// The box tests are conservative (cutoff^2 padded by 4e-6 relative, far above the 5-ulp worst case
// of the fp32 distance evaluation), so the emitted set is exactly the brute-force set of
//   fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)) < fl(r*r)              (BVHTraverse.jl:1026-1027,1248)
// — no FMA contraction: the distance uses __fmul_rn/__fadd_rn/__fsub_rn.
#include <type_traits>

#include "nb200_internal.cuh"
#include "pair_force.cuh"

namespace nb200 {

namespace {

#ifndef NB200_TRAV_WARPS
#define NB200_TRAV_WARPS 2
#endif
#ifndef NB200_MINBLOCKS
#define NB200_MINBLOCKS 14
#endif
#ifndef NB200_GHOST_PASS_BLOCKS
#define NB200_GHOST_PASS_BLOCKS 12  // blocks per SM of the ghost pass (one warp per boundary leaf while they last)
#endif
#ifndef NB200_MINBLOCKS_FUSED
#define NB200_MINBLOCKS_FUSED 14  // 72 registers, no spills (15: 64 registers with spills); 14 x 14.75 KB of shared memory per SM
#endif
constexpr int TRAV_WARPS = NB200_TRAV_WARPS;
#ifndef NB200_TGT_CAP
#define NB200_TGT_CAP 192
#endif
#ifndef NB200_GATHER
#define NB200_GATHER 3
#endif
constexpr int STACK = 192;   // wide pops while sp <= 96, then one node per round: 96 + 32 + 64 (tree depth) = 192
constexpr int STACK_WIDE_LIMIT = 96;
constexpr int CAND = 64;     // a round pops <= 32 nodes -> <= 64 leaf candidates
constexpr int TGT_CAP = NB200_TGT_CAP; // gathered target atoms per distance pass
constexpr int GATHER = NB200_GATHER;  // candidate leaves gathered per batch (16 B in flight per lane each)

template <bool FUSED>
struct __align__(16) WarpSmemT {
    float tx[TGT_CAP + 4];      // targets, SoA: the distance pass reads 4 consecutive targets per LDS.128 (broadcast)
    float ty[TGT_CAP + 4];      //   (+4 sentinels for the unrolled loop)
    float tz[TGT_CAP + 4];
    int32_t tidx[TGT_CAP];      // sorted slot of each target
    float4 sub[8];              // the query leaf's 4 sub-boxes (lo, hi) — only read for wide leaves
    int32_t stack[STACK];
    int32_t cand[CAND];
    float4 t4[FUSED ? TGT_CAP : 1];  // FUSED: the targets again as (x, y, z, charge): one LDS.128 per evaluated pair
    unsigned long long pin[2];       // FUSED: pair-loop constants on their way into registers (pinned())
};

struct FusedArgs {
    FFDev ff;
    float4* force;  // zeroed by reorder_kernel; own sums and reactions are added with 16-B vector reductions
};

__device__ __forceinline__ float gap(float alo, float ahi, float blo, float bhi) {
    return fmaxf(0.f, fmaxf(alo - bhi, blo - ahi));
}

// squared gap between the boxes [alo, ahi] and [blo, bhi] against the padded squared cutoff
__device__ __forceinline__ bool box_near(const float3& alo, const float3& ahi, const float3& blo, const float3& bhi, float r2pad) {
    float gx = gap(alo.x, ahi.x, blo.x, bhi.x);
    float gy = gap(alo.y, ahi.y, blo.y, bhi.y);
    float gz = gap(alo.z, ahi.z, blo.z, bhi.z);
    return gx * gx + gy * gy + gz * gz <= r2pad;
}
__device__ __forceinline__ float3 xyz(const float4& v) { return make_float3(v.x, v.y, v.z); }

__device__ __forceinline__ float dist2_exact(const float4& a, const float4& b) {
    float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- packed fp32 (Blackwell FADD2 / FMUL2: two IEEE binary32 operations per issue slot) ----------------------
// The distance pass is bound by instruction issue, not by the FMA pipe, so the subtractions and squares of
// two targets share one instruction.  The two ADDITIONS stay scalar on purpose: ptxas 12.9 contracts
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (even with -fmad=false), which would break the bit-exact predicate;
// a scalar __fadd_rn of an FMUL2 half is never contracted (checked in SASS: no FFMA in the loop).  (Also tried: the
// additions as fma.rn.f32x2(a, 1, b) — ptxas folds the multiplication by one and contracts that into FFMA2 as well.)
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// Squared distances query <-> two targets, each exactly fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)).
__device__ __forceinline__ void dist2_pair(unsigned long long qx, unsigned long long qy, unsigned long long qz, float x0, float x1,
                                           float y0, float y1, float z0, float z1, float& d0, float& d1) {
    unsigned long long dx = sub2(qx, pk2(x0, x1)), dy = sub2(qy, pk2(y0, y1)), dz = sub2(qz, pk2(z0, z1));
    float xa, xb, ya, yb, za, zb;
    unpk2(mul2(dx, dx), xa, xb);
    unpk2(mul2(dy, dy), ya, yb);
    unpk2(mul2(dz, dz), za, zb);
    d0 = __fadd_rn(__fadd_rn(xa, ya), za);
    d1 = __fadd_rn(__fadd_rn(xb, yb), zb);
}

// HALF = true : every unordered pair is emitted ONCE, in the row of its Morton-earlier atom — exactly the
//               reference's rule "a query leaf walks only the Morton-later part of the tree"
//               (BVHTraverse.jl:1267-1309) — subtrees whose last leaf does not come after the query leaf are
//               skipped with the leaf range the node already stores.  Forces: the reaction goes to the partner.
// HALF = false: directed list (each pair in both rows), owner-computes forces, deterministic sums.
//
// Per warp (= query leaf A) the kernel alternates two phases until the tree is exhausted:
//   FILL : tree-walk rounds produce candidate leaves; their atoms are gathered GATHER leaves at a time, tested against
//          A's box and compacted into the SoA target buffer (the leaf's own atoms are the first 32 targets);
//   DRAIN: every lane tests its query atom against all buffered targets (exact predicate -> per-lane hit masks);
//          each 32-target block goes out as one tile (and, FUSED, its pair forces are evaluated on the spot).
// MG (multi-GPU slab): the GHOST PASS.  The search arrays hold two sorted segments with a tree each — the rank's owned
// atoms (leaves [0, nL), resident in curve order like a single-GPU system) and this step's ghosts (leaves behind them;
// the ghost tree's ids carry that offset, lbvh_build.cu).  The owned segment is searched by the plain kernel as soon as
// its tree stands; the ghost pass runs when the ghost tree is ready (the halo exchange hides under the first pass):
// the same owned leaves query, `frontier` is the ghost tree's, there is no self tile, and no target gets a reaction
// (a ghost's force is its owner's business).  Every pair with at least one owned atom is emitted once by the two
// passes together, a pair of two ghosts (it belongs to other ranks) never is.
template <bool HALF, bool MG, bool FUSED>
__device__ __forceinline__ void traverse_leaf(const int A, WarpSmemT<FUSED>& S, const Node* __restrict__ nodes, const int32_t* __restrict__ frontier,
                                              const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                                              const float4* __restrict__ leaf_sub, const float4* __restrict__ pos, int n, float cutoff,
                                              int32_t* __restrict__ tiles, unsigned long long tile_capacity, GroupHdr* __restrict__ groups,
                                              unsigned int group_capacity, Counters* __restrict__ ctr, long long* __restrict__ dbg,
                                              int n_query, const FusedArgs& fa) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;

    const int ia = A * LEAF + lane;
    const bool valid_i = ia < (MG ? n_query : n);  // (MG: the pad slots of the last owned leaf hold NaN and never query)
    const float inf = __int_as_float(0x7f800000);
    const float4 pi = ia < n ? pos[ia] : make_float4(inf, 0.f, 0.f, 0.f);
    const unsigned long long qx2 = pk2(pi.x, pi.x), qy2 = pk2(pi.y, pi.y), qz2 = pk2(pi.z, pi.z);
    const float r2 = __fmul_rn(cutoff, cutoff);  // squared_radius = neighbor_distance^2 in Float32
    const int r2_bits = __float_as_int(r2);
    const float r2pad = fmaf(r2, 4e-6f, r2) + 1e-37f;
    const float3 alo = xyz(leaf_lo[A]), ahi = xyz(leaf_hi[A]);
    // a "wide" leaf: its Morton run crosses a coarse cell boundary, the AABB is much larger than the 4 sub-boxes
    // of its sub-runs; only then do the (more expensive) sub-box tests pay for themselves
    const float wlim = wide_leaf_limit(cutoff);
    const bool wide = (ahi.x - alo.x > wlim) || (ahi.y - alo.y > wlim) || (ahi.z - alo.z > wlim);
    if (wide && lane < 8) S.sub[lane] = leaf_sub[(size_t)A * 8 + lane];  // written by reorder_kernel for wide leaves only
    // the leaf's own atoms are the first 32 targets
    S.tx[lane] = pi.x; S.ty[lane] = pi.y; S.tz[lane] = pi.z;
    S.tidx[lane] = ia;
    if (FUSED) S.t4[lane] = pi;
    // Start from the tree's precomputed frontier (<= 32 entries of the first levels) instead of the root: internal
    // nodes go on the stack, the rare leaf entries of a small tree are box-tested here and become candidates.
    int sp = 0, ncand = 0, cpos = 0;  // stack size, candidates of the last round, next one to gather
    __syncwarp(full);                 // (S.sub)
    {
        const int nf = frontier[0];
        const int e = lane < nf ? frontier[1 + lane] : 0;
        // the entry's own box first (they sit behind the list, lbvh_build.cu): far entries are never pushed, and a leaf
        // that is near none — most leaves in the ghost pass — is done after this one test
        const float4* __restrict__ fbox = reinterpret_cast<const float4*>(frontier + 64);
        bool near_e = false;
        if (lane < nf) {
            const float3 elo = xyz(__ldg(&fbox[2 * lane])), ehi = xyz(__ldg(&fbox[2 * lane + 1]));
            near_e = box_near(alo, ahi, elo, ehi, r2pad);
            if (wide && near_e) {
                near_e = false;
#pragma unroll
                for (int r = 0; r < 4; ++r) near_e = near_e || box_near(xyz(S.sub[2 * r]), xyz(S.sub[2 * r + 1]), elo, ehi, r2pad);
            }
        }
        const bool is_node = near_e && e >= 0;
        bool is_cand = false;
        if (near_e && e < 0) {
            const int B = ~e;
            is_cand = HALF ? B > A : B != A;
        }
        const unsigned mn = __ballot_sync(full, is_node), mc = __ballot_sync(full, is_cand);
        if (is_node) S.stack[__popc(mn & lt_mask)] = e;
        if (is_cand) S.cand[__popc(mc & lt_mask)] = ~e;
        sp = __popc(mn);
        ncand = __popc(mc);
    }
    // self tile (target t sits in bit t): HALF keeps the partners after me, directed drops only myself
    const unsigned self_mask = HALF ? (0xfffffffeu << lane) : ~(1u << lane);
    __syncwarp(full);

    int ntgt = MG ? 0 : 32;        // (ghost pass: no self tile)
    bool first_drain = !MG;
    int hits = 0;  // set bits of my masks
    float fx = 0.f, fy = 0.f, fz = 0.f;  // FUSED: force on my query atom
    long long dbg_t0 = dbg ? clock64() : 0, dbg_cand = 0, dbg_rounds = 0, dbg_targets = 0;

    // pair-loop constants, pinned in registers (see the pair loop)
    PairConsts pc = {0.f, 0.f, 0.f, 0.f};
    unsigned long long fbase = 0;
    bool has_q = false;
    if (FUSED) {
        const unsigned long long c = pinned(&S.pin[0], ((unsigned long long)__float_as_uint(fa.ff.eps24) << 32) | __float_as_uint(fa.ff.sigma2));
        const float e24 = __uint_as_float((unsigned)(c >> 32));
        pc.sigma2 = __uint_as_float((unsigned)c);
        pc.eps48 = 2.0f * e24;
        pc.neps24 = -e24;
        pc.kq = fa.ff.kcoul * pi.w;
        has_q = fa.ff.kcoul != 0.0f;
        fbase = pinned(&S.pin[1], (unsigned long long)__cvta_generic_to_global(fa.force));
    }
    auto near_sub = [&](const float3& blo, const float3& bhi) {
        bool hit = false;
#pragma unroll
        for (int r = 0; r < 4; ++r) hit = hit || box_near(xyz(S.sub[2 * r]), xyz(S.sub[2 * r + 1]), blo, bhi, r2pad);
        return hit;
    };

    for (;;) {
        // =============================== FILL ===============================
        bool more = true;
        while (ntgt <= TGT_CAP - GATHER * 32) {
            if (cpos < ncand) {
                // ---- gather up to GATHER candidate leaves, keep the atoms near A's box ----
                float4 pc[GATHER];
                int jb[GATHER];
                bool near[GATHER];
#pragma unroll
                for (int u = 0; u < GATHER; ++u) {
                    jb[u] = (cpos + u < ncand) ? S.cand[cpos + u] * LEAF + lane : n;
                    if (jb[u] < n) pc[u] = __ldg(&pos[jb[u]]);
                }
#pragma unroll
                for (int u = 0; u < GATHER; ++u) {
                    near[u] = false;
                    if (jb[u] < n) {
                        const float3 p = xyz(pc[u]);
                        near[u] = box_near(alo, ahi, p, p, r2pad);
                        if (wide && near[u]) near[u] = near_sub(p, p);
                    }
                }
                unsigned msk[GATHER];
#pragma unroll
                for (int u = 0; u < GATHER; ++u) msk[u] = __ballot_sync(full, near[u]);  // independent votes, no branch between
#pragma unroll
                for (int u = 0; u < GATHER; ++u) {
                    if (near[u]) {
                        const int k = ntgt + __popc(msk[u] & lt_mask);
                        S.tx[k] = pc[u].x; S.ty[k] = pc[u].y; S.tz[k] = pc[u].z;
                        S.tidx[k] = jb[u];
                        if (FUSED) S.t4[k] = pc[u];
                    }
                    ntgt += __popc(msk[u]);
                }
                dbg_cand += min(GATHER, ncand - cpos);
                cpos += GATHER;
            } else if (sp > 0) {
                // ---- one cooperative round of the tree walk: <= 32 nodes popped, both children tested ----
                // Above STACK_WIDE_LIMIT one node is popped per round (net growth <= 1), so the stack holds trees of
                // any depth up to STACK - STACK_WIDE_LIMIT - 32 levels of two-sided hits; beyond that the warp gives
                // up loudly (sticky flag -> NB200_ERR_STATE at the next sync) instead of overrunning shared memory.
                const int m = (sp > STACK_WIDE_LIMIT) ? 1 : min(sp, 32);
                ++dbg_rounds;
                const bool have = lane < m;
                const int nd = have ? S.stack[sp - 1 - lane] : 0;
                __syncwarp(full);
                bool pushL = false, pushR = false, candL = false, candR = false;
                int left_id = 0, right_id = 0;
                if (have) {
                    const float4* np = reinterpret_cast<const float4*>(&nodes[nd]);
                    const float4 c0 = __ldg(np), c1 = __ldg(np + 1), c2 = __ldg(np + 2), c3 = __ldg(np + 3);
                    left_id = __float_as_int(c0.w);
                    right_id = __float_as_int(c1.w);
                    bool hitL = box_near(alo, ahi, xyz(c0), xyz(c1), r2pad);
                    bool hitR = box_near(alo, ahi, xyz(c2), xyz(c3), r2pad);
                    // Karras numbering (lbvh_build.cu): the left child ends at leaf p = its own id, the right child
                    // at the node's last leaf.  HALF: a subtree that does not reach past my leaf holds no later
                    // partner; directed: only my own leaf is excluded (it is the self tile already).
                    const int left_last = left_id >= 0 ? left_id : ~left_id;
                    if (HALF) {
                        hitL = hitL && left_last > A;
                        hitR = hitR && __float_as_int(c3.w) > A;
                    } else {
                        hitL = hitL && left_id != ~A;
                        hitR = hitR && right_id != ~A;
                    }
                    if (wide) {
                        if (hitL) hitL = near_sub(xyz(c0), xyz(c1));
                        if (hitR) hitR = near_sub(xyz(c2), xyz(c3));
                    }
                    pushL = hitL && left_id >= 0;
                    candL = hitL && left_id < 0;
                    pushR = hitR && right_id >= 0;
                    candR = hitR && right_id < 0;
                }
                const unsigned bL = __ballot_sync(full, pushL), bR = __ballot_sync(full, pushR);
                const unsigned cL = __ballot_sync(full, candL), cR = __ballot_sync(full, candR);
                const int newsp = sp - m;
                if (newsp + __popc(bL) + __popc(bR) > STACK) {  // cannot happen in wide mode (96 + 64 <= 192): m == 1 here
                    if (lane == 0) atomicExch(&ctr->stack_overflow, 1u);
                    pushL = pushR = false;  // drop the subtree: the list is incomplete and the sync says so
                    sp = newsp;
                } else {
                    if (pushL) S.stack[newsp + __popc(bL & lt_mask)] = left_id;
                    if (pushR) S.stack[newsp + __popc(bL) + __popc(bR & lt_mask)] = right_id;
                    sp = newsp + __popc(bL) + __popc(bR);
                }
                if (candL) S.cand[__popc(cL & lt_mask)] = ~left_id;
                if (candR) S.cand[__popc(cL) + __popc(cR & lt_mask)] = ~right_id;
                ncand = __popc(cL) + __popc(cR);
                cpos = 0;
                __syncwarp(full);
            } else {
                more = false;
                break;
            }
        }
        // =============================== DRAIN ===============================
        __syncwarp(full);
        dbg_targets += ntgt;
        if (lane < 4) {  // sentinels: d2 = +inf, never a hit
            S.tx[ntgt + lane] = inf;
            S.ty[ntgt + lane] = 0.f;
            S.tz[ntgt + lane] = 0.f;
        }
        // one packed atomic reserves this pass's tiles and its group header
        const int nt = (ntgt + 31) >> 5;
        unsigned long long tile0 = 0;
        bool fits = false;
        if (nt > 0) {
            unsigned long long a = 0;
            if (lane == 0) a = atomicAdd(&ctr->alloc, ((unsigned long long)nt << SEG_BITS) | 1ull);
            a = __shfl_sync(full, a, 0);
            tile0 = a >> SEG_BITS;
            const unsigned int grp = (unsigned int)(a & ((1ull << SEG_BITS) - 1ull));
            fits = (tile0 + (unsigned long long)nt <= tile_capacity) && grp < group_capacity;
            if (lane == 0) {
                if (grp < group_capacity) {
                    const unsigned int w1 = fits ? ((unsigned int)nt | (first_drain ? GROUP_SELF : 0u)) : 0u;
                    *reinterpret_cast<uint4*>(&groups[grp]) = make_uint4((unsigned int)A, w1, (unsigned int)tile0, (unsigned int)(tile0 >> 32));
                }
                if (!fits) {
                    atomicExch(&ctr->overflow, 1u);
                    atomicExch(&ctr->overflow_sticky, 1u);
                }
            }
        }
        __syncwarp(full);
        for (int t0 = 0; t0 < ntgt; t0 += 32) {
            // exact predicate for 32 targets -> hit mask: d2 >= 0 and r2 >= 0, so the integer difference of the
            // float bit patterns is negative iff d2 < r2; its sign bit is funnel-shifted into the mask.
            // The quads are walked from the END of the block, so target t0 + b ends up in bit b.
            const int quads = (min(32, ntgt - t0) + 3) >> 2;
            unsigned mask = 0;
            auto quad = [&](int u) {
                const float4 X = *reinterpret_cast<const float4*>(&S.tx[t0 + 4 * u]);
                const float4 Y = *reinterpret_cast<const float4*>(&S.ty[t0 + 4 * u]);
                const float4 Z = *reinterpret_cast<const float4*>(&S.tz[t0 + 4 * u]);
                float d0, d1, d2, d3;
                dist2_pair(qx2, qy2, qz2, X.x, X.y, Y.x, Y.y, Z.x, Z.y, d0, d1);
                dist2_pair(qx2, qy2, qz2, X.z, X.w, Y.z, Y.w, Z.z, Z.w, d2, d3);
                mask = __funnelshift_l((unsigned)(__float_as_int(d3) - r2_bits), mask, 1);
                mask = __funnelshift_l((unsigned)(__float_as_int(d2) - r2_bits), mask, 1);
                mask = __funnelshift_l((unsigned)(__float_as_int(d1) - r2_bits), mask, 1);
                mask = __funnelshift_l((unsigned)(__float_as_int(d0) - r2_bits), mask, 1);
            };
            if (quads == 8) {  // a full block (most of them): straight-line code, no loop counter
#pragma unroll
                for (int u = 7; u >= 0; --u) quad(u);
            } else {
#pragma unroll 1
                for (int u = quads - 1; u >= 0; --u) quad(u);
            }
            if (!valid_i) mask = 0u;
            const bool self_tile = first_drain && t0 == 0;
            const unsigned raw = mask;  // self tile: symmetric (the predicate is), diagonal included
            if (self_tile) mask &= self_mask;
            hits += __popc(mask);
            if (fits) {
                int32_t* __restrict__ T = tiles + (tile0 + (unsigned)(t0 >> 5)) * TILE_WORDS;
                T[lane] = (t0 + lane < ntgt) ? S.tidx[t0 + lane] : -1;
                T[32 + lane] = (int32_t)mask;
            }
            if (FUSED) {
                // ---- pair forces of this tile, targets still in shared memory ----
                // My mask bits are the pairs whose force on MY atom I evaluate.  Half list: the reaction goes to the
                // partner with one 16-byte vector reduction — except in the self tile, where the symmetric mask gives
                // every atom its complete row inside the leaf.  Directed list: no reactions at all.
                // MG (ghost pass): every target is a ghost and gets nothing — its owner computes that force itself.
                // (Tried and measured slower, see DESIGN.md: reaction recomputed by the target lane from the transposed
                //  masks; one loop per drain pass or per leaf over a lane's whole row, with and without a partner lane
                //  taking over part of a long row; pairs dealt out evenly with per-query sums in shared memory.  The
                //  extra cursor work per pair costs what the better lane balance saves.)
                unsigned mr = (HALF && self_tile) ? (raw & ~(1u << lane)) : mask;
                const bool react = HALF && !self_tile && !MG;
                // (loop shape: see pair_force.cuh — 34 SASS instructions per pair with reaction and charges)
                const unsigned tb = (unsigned)__cvta_generic_to_shared(&S.t4[t0]);
                const unsigned ib = (unsigned)__cvta_generic_to_shared(&S.tidx[t0]);
                auto pair_loop = [&](auto react_c, auto q_c) {  // (unswitched by hand: no branch inside the loop)
                    constexpr bool REACT = decltype(react_c)::value, Q = decltype(q_c)::value;
                    while (mr) {
                        const int b = top_bit(mr);
                        mr ^= bit_at(b);
                        const float4 T = lds128(tb + 16u * (unsigned)b);
                        const float dx = pi.x - T.x, dy = pi.y - T.y, dz = pi.z - T.z;
                        const float fs = pair_fs<Q>(pc, dx * dx + dy * dy + dz * dz, T.w);
                        fx = fmaf(fs, dx, fx); fy = fmaf(fs, dy, fy); fz = fmaf(fs, dz, fz);
                        if (REACT) red_add4(fbase, lds32(ib + 4u * (unsigned)b), -fs * dx, -fs * dy, -fs * dz);
                    }
                };
                using std::true_type;
                using std::false_type;
                if (react) { if (has_q) pair_loop(true_type{}, true_type{}); else pair_loop(true_type{}, false_type{}); }
                else       { if (has_q) pair_loop(false_type{}, true_type{}); else pair_loop(false_type{}, false_type{}); }
            }
        }
        ntgt = 0;
        first_drain = false;
        __syncwarp(full);
        if (!more) break;
    }
    const int n_emitted = __reduce_add_sync(full, hits);
    if (lane == 0 && n_emitted > 0) atomicAdd(&ctr->n_valid, (unsigned long long)n_emitted);
    if (FUSED && valid_i) atomicAdd(&fa.force[ia], make_float4(fx, fy, fz, 0.f));
    if (dbg && lane == 0) {
        dbg[4 * A + 0] = clock64() - dbg_t0;
        dbg[4 * A + 1] = dbg_cand;
        dbg[4 * A + 2] = dbg_rounds;
        dbg[4 * A + 3] = dbg_targets;
    }
}

// Owned leaves that are near the ghost tree at all (their box within the cutoff of a ghost frontier entry's box): the only
// leaves the ghost pass has to visit.  One thread per owned leaf, warp-aggregated append.  (The test is the prologue's own
// pre-test without its sub-box refinement, i.e. conservative: a leaf that is not listed would have pushed nothing.)
__global__ void __launch_bounds__(256)
    boundary_leaves_kernel(const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi, int nL, const int32_t* __restrict__ frontier,
                           float cutoff, int32_t* __restrict__ list, unsigned int* __restrict__ count) {
    __shared__ float4 s_box[64];
    __shared__ int s_nf;
    if (threadIdx.x == 0) s_nf = min(frontier[0], 32);
    if (threadIdx.x < 64) s_box[threadIdx.x] = reinterpret_cast<const float4*>(frontier + 64)[threadIdx.x];
    __syncthreads();
    const int A = blockIdx.x * blockDim.x + threadIdx.x;
    const float r2 = __fmul_rn(cutoff, cutoff);
    const float r2pad = fmaf(r2, 4e-6f, r2) + 1e-37f;
    bool near_any = false;
    if (A < nL) {
        const float3 alo = xyz(leaf_lo[A]), ahi = xyz(leaf_hi[A]);
        for (int e = 0; e < s_nf; ++e) near_any = near_any || box_near(alo, ahi, xyz(s_box[2 * e]), xyz(s_box[2 * e + 1]), r2pad);
    }
    const unsigned m = __ballot_sync(0xffffffffu, near_any);
    if (m) {
        const int lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(count, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (near_any) list[base + __popc(m & ((1u << lane) - 1u))] = A;
    }
}

// One warp per query leaf.  The ghost pass (MG) visits only the listed boundary leaves (blist / bcount; ~9 % of the owned
// leaves at 1 M atoms per GPU and two ranks), one warp per leaf while they last: its earlier form — a persistent grid of 3
// blocks per SM that looped over ALL owned leaves — executed 4 % of the owned pass's instructions but ran for half of its
// duration (each warp worked through its boundary leaves one after the other, latency-bound) and held a quarter of the block
// slots of every SM all that time.
template <bool HALF, bool MG, bool FUSED>
__global__ void __launch_bounds__(TRAV_WARPS * 32, FUSED ? NB200_MINBLOCKS_FUSED : NB200_MINBLOCKS)
    traverse_kernel(const Node* __restrict__ nodes, const int32_t* __restrict__ frontier, const float4* __restrict__ leaf_lo,
                    const float4* __restrict__ leaf_hi,
                    const float4* __restrict__ leaf_sub, const float4* __restrict__ pos, int n, int nL, float cutoff,
                    int32_t* __restrict__ tiles, unsigned long long tile_capacity, GroupHdr* __restrict__ groups,
                    unsigned int group_capacity, Counters* __restrict__ ctr, long long* __restrict__ dbg /* [nL][4] or null */,
                    int n_query /* MG: owned atoms (the pad slots behind them never query) */, const FusedArgs fa,
                    const int32_t* __restrict__ blist, const unsigned int* __restrict__ bcount) {
    using Smem = WarpSmemT<FUSED>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& S = reinterpret_cast<Smem*>(smem_raw)[threadIdx.x >> 5];
    const int w0 = blockIdx.x * TRAV_WARPS + (threadIdx.x >> 5);  // whole warps leave together; no block-wide barriers anywhere
    if (MG) {
        const int nb = blist ? (int)min(*bcount, (unsigned)nL) : nL;
        for (int k = w0; k < nb; k += gridDim.x * TRAV_WARPS) {
            const int A = blist ? blist[k] : k;
            traverse_leaf<HALF, MG, FUSED>(A, S, nodes, frontier, leaf_lo, leaf_hi, leaf_sub, pos, n, cutoff, tiles, tile_capacity, groups,
                                           group_capacity, ctr, dbg, n_query, fa);
            __syncwarp(0xffffffffu);
        }
    } else if (w0 < nL) {
        traverse_leaf<HALF, MG, FUSED>(w0, S, nodes, frontier, leaf_lo, leaf_hi, leaf_sub, pos, n, cutoff, tiles, tile_capacity, groups,
                                       group_capacity, ctr, dbg, n_query, fa);
    }
}

// ---- export: directed list -> the reference's unique (a, b, d) tuples -------------------------------------
// Keeps the entry of each pair that sits in the row of the Morton-earlier atom, converts sorted slots
// to original ids, orients the tuple like the reference (a = first atom in ITS sort order: 10-bit
// mortoncodes! key, then atom id; BVHTraverse.jl:259-284,570,1032,1050) and emits d = sqrt_rn(d2)
// (:1028,1049).
__device__ __forceinline__ int code10_ref(const float4& p) {
    const float binwidth = (float)(1.0 / 1023.0);  // const binwidth = Float32(1/1023)  (:184)
    int qx = __float2int_rd(__fdiv_rn(p.x, binwidth));
    int qy = __float2int_rd(__fdiv_rn(p.y, binwidth));
    int qz = __float2int_rd(__fdiv_rn(p.z, binwidth));
    return (qx & 0x09249249) | (qy & 0x12492492) | (qz & 0x24924924);  // magic_values (:241)
}

// One warp per group; lane <-> query atom of the group's leaf.  For every tile the 32 target columns are visited in
// turn: the lanes whose mask has the bit vote, one atomic reserves the output slots of that column.
__global__ void __launch_bounds__(256)
    export_kernel(const GroupHdr* __restrict__ groups, const int32_t* __restrict__ tiles, Counters* __restrict__ ctr,
                  unsigned int group_capacity, const float4* __restrict__ pos, const int32_t* __restrict__ id, int n,
                  int32_t* __restrict__ out_a, int32_t* __restrict__ out_b, float* __restrict__ out_d,
                  unsigned long long capacity, int index_base) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned ngrp = min(ctr->n_segments(), group_capacity);
    for (unsigned g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ngrp; g += nwarps) {
        const GroupHdr H = groups[g];
        const int nt = (int)(H.ntiles & ~GROUP_SELF);
        if (nt == 0) continue;
        const int ia = H.leaf * LEAF + lane;
        const bool valid = ia < n;
        const float4 pi = valid ? pos[ia] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int ida = valid ? id[ia] : 0;
        const int ca = code10_ref(pi);
        for (int k = 0; k < nt; ++k) {
            const int32_t* __restrict__ T = tiles + (H.base_tile + (unsigned)k) * TILE_WORDS;
            const int tj = T[lane] & 0x7fffffff;
            const unsigned mask = (unsigned)T[32 + lane];
            // my target's data, handed to the query lanes by shuffle
            const bool tv = tj < n;
            const float4 pt = tv ? pos[tj] : make_float4(0.f, 0.f, 0.f, 0.f);
            const int idt = tv ? id[tj] : 0;
            unsigned cols = __reduce_or_sync(full, mask);
            while (cols) {
                const int b = __ffs(cols) - 1;
                cols &= cols - 1;
                const int j = __shfl_sync(full, tj, b);
                const float4 pj = make_float4(__shfl_sync(full, pt.x, b), __shfl_sync(full, pt.y, b), __shfl_sync(full, pt.z, b), 0.f);
                const int idb = __shfl_sync(full, idt, b);
                // a directed list holds the pair in both rows: keep the entry of the curve-earlier atom
                const bool keep = ((mask >> b) & 1u) && (j > ia);
                const unsigned km = __ballot_sync(full, keep);
                if (km == 0) continue;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&ctr->n_export, (unsigned long long)__popc(km));
                base = __shfl_sync(full, base, 0);
                if (keep) {
                    const unsigned long long slot = base + __popc(km & lt_mask);
                    if (slot < capacity) {
                        const int cb = code10_ref(pj);
                        const bool a_first = (ca < cb) || (ca == cb && ida < idb);
                        out_a[slot] = (a_first ? ida : idb) + index_base;
                        out_b[slot] = (a_first ? idb : ida) + index_base;
                        out_d[slot] = __fsqrt_rn(dist2_exact(pi, pj));
                    }
                }
            }
        }
    }
}

// every list entry as (pre-sort index of the row atom, pre-sort index of the partner, d): the multi-GPU
// parity check unions these over the ranks
__global__ void __launch_bounds__(256)
    export_directed_kernel(const GroupHdr* __restrict__ groups, const int32_t* __restrict__ tiles, Counters* __restrict__ ctr,
                           unsigned int group_capacity, const float4* __restrict__ pos, const int32_t* __restrict__ id, int n,
                           int32_t* __restrict__ out_a, int32_t* __restrict__ out_b, float* __restrict__ out_d,
                           unsigned long long capacity) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned ngrp = min(ctr->n_segments(), group_capacity);
    for (unsigned g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ngrp; g += nwarps) {
        const GroupHdr H = groups[g];
        const int nt = (int)(H.ntiles & ~GROUP_SELF);
        if (nt == 0) continue;
        const int ia = H.leaf * LEAF + lane;
        const bool valid = ia < n;
        const float4 pi = valid ? pos[ia] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int ida = valid ? id[ia] : 0;
        for (int k = 0; k < nt; ++k) {
            const int32_t* __restrict__ T = tiles + (H.base_tile + (unsigned)k) * TILE_WORDS;
            const int tj = T[lane] & 0x7fffffff;
            const unsigned mask = (unsigned)T[32 + lane];
            const bool tv = tj < n;
            const float4 pt = tv ? pos[tj] : make_float4(0.f, 0.f, 0.f, 0.f);
            const int idt = tv ? id[tj] : 0;
            const int total = __reduce_add_sync(full, __popc(mask));
            if (total == 0) continue;
            unsigned long long obase = 0;
            if (lane == 0) obase = atomicAdd(&ctr->n_export, (unsigned long long)total);
            obase = __shfl_sync(full, obase, 0);
            unsigned cols = __reduce_or_sync(full, mask);
            while (cols) {
                const int b = __ffs(cols) - 1;
                cols &= cols - 1;
                const float4 pj = make_float4(__shfl_sync(full, pt.x, b), __shfl_sync(full, pt.y, b), __shfl_sync(full, pt.z, b), 0.f);
                const int idb = __shfl_sync(full, idt, b);
                const bool act = (mask >> b) & 1u;
                const unsigned m = __ballot_sync(full, act);
                const unsigned long long slot = obase + __popc(m & lt_mask);
                if (act && slot < capacity) {
                    out_a[slot] = ida;
                    out_b[slot] = idb;
                    out_d[slot] = __fsqrt_rn(dist2_exact(pi, pj));
                }
                obase += __popc(m);
            }
        }
    }
}

// neighbours per atom (original order): the row atom counts its mask bits; in a half list the pair also counts for
// the partner, whose lane gets its count from the transposed tile
__global__ void __launch_bounds__(256)
    neighbor_counts_kernel(const GroupHdr* __restrict__ groups, const int32_t* __restrict__ tiles, const Counters* __restrict__ ctr,
                           unsigned int group_capacity, const int32_t* __restrict__ id, int n, int32_t* __restrict__ counts, int half) {
    const int lane = threadIdx.x & 31;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned ngrp = min(ctr->n_segments(), group_capacity);
    for (unsigned g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ngrp; g += nwarps) {
        const GroupHdr H = groups[g];
        const int nt = (int)(H.ntiles & ~GROUP_SELF);
        if (nt == 0) continue;
        const int ia = H.leaf * LEAF + lane;
        int c = 0;
        for (int k = 0; k < nt; ++k) {
            const int32_t* __restrict__ T = tiles + (H.base_tile + (unsigned)k) * TILE_WORDS;
            const int tj = T[lane] & 0x7fffffff;
            const unsigned mask = (unsigned)T[32 + lane];
            c += __popc(mask);
            if (half) {
                const int ct = __popc(transpose32(mask, lane));
                if (ct && tj < n) atomicAdd(&counts[id[tj]], ct);
            }
        }
        if (ia < n && c) atomicAdd(&counts[id[ia]], c);
    }
}

}  // namespace

// n_leaves: QUERY leaves (all leaves; multi-GPU: the owned leaves).  mg != nullptr: the GHOST PASS over two-segment
// search arrays (see traverse_kernel): `frontier` is the ghost tree's, mg->n_query the owned atoms.
int launch_traverse(cudaStream_t s, int sm_count, const Node* nodes, const int32_t* frontier, const float4* leaf_lo, const float4* leaf_hi,
                    const float4* leaf_sub, const float4* pos, int n, int n_leaves, float cutoff, int32_t* entries, int64_t entry_capacity,
                    GroupHdr* segs, int64_t seg_capacity, Counters* counters, bool half, long long* dbg, const MgSearch* mg,
                    bool counters_clean, const ForceField* fused_ff, float4* fused_force) {
    if (!counters_clean) cudaMemsetAsync(counters, 0, COUNTERS_RESET_BYTES, s);  // alloc, n_valid, overflow
    int blocks = (n_leaves + TRAV_WARPS - 1) / TRAV_WARPS;
    if (mg && blocks > sm_count * NB200_GHOST_PASS_BLOCKS) blocks = sm_count * NB200_GHOST_PASS_BLOCKS;  // the ghost pass visits only the boundary leaves: it runs beside the owned pass
    const unsigned long long tile_capacity = (unsigned long long)(entry_capacity / TILE_WORDS);
    typedef void (*Kern)(const Node*, const int32_t*, const float4*, const float4*, const float4*, const float4*, int, int, float, int32_t*,
                         unsigned long long, GroupHdr*, unsigned int, Counters*, long long*, int, const FusedArgs, const int32_t*,
                         const unsigned int*);
    static const Kern table[8] = {traverse_kernel<false, false, false>, traverse_kernel<false, false, true>,
                                  traverse_kernel<false, true, false>,  traverse_kernel<false, true, true>,
                                  traverse_kernel<true, false, false>,  traverse_kernel<true, false, true>,
                                  traverse_kernel<true, true, false>,   traverse_kernel<true, true, true>};
    const bool fused = fused_ff != nullptr && fused_force != nullptr;
    const Kern kern = table[(half ? 4 : 0) | (mg ? 2 : 0) | (fused ? 1 : 0)];
    const size_t smem = (fused ? sizeof(WarpSmemT<true>) : sizeof(WarpSmemT<false>)) * TRAV_WARPS;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FusedArgs fa = {};
    if (fused) {
        fa.ff = make_ffdev(*fused_ff);
        fa.force = fused_force;
    }
    int launches = 1;
    if (mg && mg->blist && n_leaves > 0) {
        cudaMemsetAsync(mg->bcount, 0, sizeof(unsigned int), s);
        boundary_leaves_kernel<<<(n_leaves + 255) / 256, 256, 0, s>>>(leaf_lo, leaf_hi, n_leaves, frontier, cutoff, mg->blist, mg->bcount);
        ++launches;
    }
    if (blocks > 0)
        kern<<<blocks, TRAV_WARPS * 32, smem, s>>>(nodes, frontier, leaf_lo, leaf_hi, leaf_sub, pos, n, n_leaves, cutoff, entries, tile_capacity,
                                                   segs, (unsigned int)seg_capacity, counters, dbg, mg ? mg->n_query : n, fa,
                                                   mg ? mg->blist : nullptr, mg ? mg->bcount : nullptr);
    return launches;
}

int launch_export(cudaStream_t s, int sm_count, const GroupHdr* segs, const int32_t* entries, Counters* counters,
                  int64_t seg_capacity, const float4* pos, const int32_t* id, int n, int32_t* a, int32_t* b, float* d,
                  int64_t capacity, int index_base) {
    cudaMemsetAsync(&counters->n_export, 0, sizeof(unsigned long long), s);
    export_kernel<<<sm_count * 8, 256, 0, s>>>(segs, entries, counters, (unsigned int)seg_capacity, pos, id, n, a, b, d,
                                               (unsigned long long)capacity, index_base);
    return 1;
}

int launch_export_directed(cudaStream_t s, int sm_count, const GroupHdr* segs, const int32_t* entries, Counters* counters,
                           int64_t seg_capacity, const float4* pos, const int32_t* id, int n, int32_t* a, int32_t* b, float* d,
                           int64_t capacity) {
    cudaMemsetAsync(&counters->n_export, 0, sizeof(unsigned long long), s);
    export_directed_kernel<<<sm_count * 8, 256, 0, s>>>(segs, entries, counters, (unsigned int)seg_capacity, pos, id, n, a, b, d,
                                                        (unsigned long long)capacity);
    return 1;
}

int launch_neighbor_counts(cudaStream_t s, int sm_count, const GroupHdr* segs, const int32_t* entries, const Counters* counters,
                           int64_t seg_capacity, const int32_t* id, int n, int32_t* counts, bool half) {
    cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n, s);
    neighbor_counts_kernel<<<sm_count * 4, 256, 0, s>>>(segs, entries, counters, (unsigned int)seg_capacity, id, n, counts, half ? 1 : 0);
    return 1;
}

// Tuning (NB200_CARVEOUT, see atoms.cu)
void carveout_traverse(int pct) {
    cudaFuncSetAttribute(traverse_kernel<true, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(traverse_kernel<true, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(traverse_kernel<false, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(traverse_kernel<false, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(boundary_leaves_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(traverse_kernel<true, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(traverse_kernel<true, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(traverse_kernel<false, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(traverse_kernel<false, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

}  // namespace nb200
