// traverse.cu — warp-cooperative cutoff traversal of the LBVH, emitting the compacted neighbour list.
//
// Replaces leafneighbor_traverse + aabb_overlap_test + one/twocluster_proximitytest!
// (BVHTraverse.jl:1236-1323, 1094-1098, 1021-1055).  Like the reference's live path it is a
// LEAF-query traversal (its own finding: leaf-vs-tree beats atom-vs-tree, devdiary.md:1415), but
// re-shaped for a warp:
//   * one warp owns one query leaf A = 32 Morton-consecutive atoms, lane <-> atom;
//   * the tree walk is cooperative: up to 32 internal nodes are popped from a short shared-memory
//     stack per round, every lane tests BOTH child boxes of its node (they live in the 64-B parent),
//     hit internal children are pushed, hit leaf children become candidate tiles;
//   * every candidate leaf B is a 32x32 tile: B's atoms are staged in shared memory, lanes whose
//     atom is farther than the cutoff from A's box are dropped from the target mask, then each lane
//     tests its query atom against the surviving targets with the reference's exact predicate;
//   * hits are buffered per lane in shared memory ([round][lane], conflict free) and flushed as one
//     segment: a single atomicAdd reserves the space, rows are written interleaved-compact, so the
//     list write is fully coalesced and 4 B per directed entry.
// The box tests are conservative (cutoff^2 padded by 4e-6 relative, far above the 5-ulp worst case
// of the fp32 distance evaluation), so the emitted set is exactly the brute-force set of
//   fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)) < fl(r*r)              (BVHTraverse.jl:1026-1027,1248)
// — no FMA contraction: the distance uses __fmul_rn/__fadd_rn/__fsub_rn.
#include "nb200_internal.cuh"

namespace nb200 {

namespace {

#ifndef NB200_TRAV_WARPS
#define NB200_TRAV_WARPS 2
#endif
#ifndef NB200_KMAX
#define NB200_KMAX 72
#endif
#ifndef NB200_CHUNK
#define NB200_CHUNK 8
#endif
#ifndef NB200_MINBLOCKS
#define NB200_MINBLOCKS 10
#endif
constexpr int TRAV_WARPS = NB200_TRAV_WARPS;
constexpr int KMAX = NB200_KMAX;    // row buffer depth per lane (entries); rows are flushed when a lane may exceed it
constexpr int CHUNK = NB200_CHUNK;  // targets tested between two row-capacity checks
constexpr int STACK = 192;   // wide pops while sp <= 96, then one node per round: 96 + 32 + 64 (tree depth) = 192
constexpr int STACK_WIDE_LIMIT = 96;
constexpr int CAND = 64;     // a round pops <= 32 nodes -> <= 64 leaf candidates
constexpr int TGT_CAP = 256; // gathered target atoms per distance pass
constexpr int GATHER = 4;    // candidate leaves loaded per gather batch (4 x 16 B in flight per lane)
constexpr int CTAB = 256;    // candidate-leaf table: a buffered row entry is (table slot << 5 | lane), 16 bits

struct __align__(16) WarpSmem {
    float4 tgt[TGT_CAP + 4];    // x, y, z, (int bits) entry code; +4 sentinels for the unrolled loop
    uint16_t rows[KMAX * 32];   // [round][lane] entry codes
    int32_t stack[STACK];
    int32_t cand[CAND];
    int32_t ctab[CTAB];         // table slot -> leaf index
};

__device__ __forceinline__ float gap(float alo, float ahi, float blo, float bhi) {
    return fmaxf(0.f, fmaxf(alo - bhi, blo - ahi));
}

__device__ __forceinline__ bool box_near(const float3& alo, const float3& ahi, const float4& blo, const float4& bhi,
                                         float r2pad) {
    float gx = gap(alo.x, ahi.x, blo.x, bhi.x);
    float gy = gap(alo.y, ahi.y, blo.y, bhi.y);
    float gz = gap(alo.z, ahi.z, blo.z, bhi.z);
    return gx * gx + gy * gy + gz * gz <= r2pad;
}

__device__ __forceinline__ float dist2_exact(const float4& a, const float4& b) {
    float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Query region of a leaf: its AABB and the 4 sub-boxes of its Morton sub-runs.
struct Region {
    float3 alo, ahi;
    float3 slo[4], shi[4];
    float r2pad;
    bool wide;  // the AABB is much larger than the sub-boxes (the run crosses a coarse cell boundary)
    __device__ __forceinline__ bool near_aabb(const float4& blo, const float4& bhi) const { return box_near(alo, ahi, blo, bhi, r2pad); }
    __device__ __forceinline__ bool near_sub(const float4& blo, const float4& bhi) const {
        return box_near(slo[0], shi[0], blo, bhi, r2pad) || box_near(slo[1], shi[1], blo, bhi, r2pad) ||
               box_near(slo[2], shi[2], blo, bhi, r2pad) || box_near(slo[3], shi[3], blo, bhi, r2pad);
    }
    // tree walk: the sub-box tests only pay for themselves on wide leaves (measured: for ordinary leaves they
    // remove ~15 % of the candidates but cost more instructions than gathering those candidates)
    __device__ __forceinline__ bool near_node(const float4& blo, const float4& bhi) const {
        if (!near_aabb(blo, bhi)) return false;
        return wide ? near_sub(blo, bhi) : true;
    }
};

__global__ void __launch_bounds__(TRAV_WARPS * 32, NB200_MINBLOCKS)
    traverse_kernel(const Node* __restrict__ nodes, const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                    const float4* __restrict__ leaf_sub, const float4* __restrict__ pos, int n, int nL, float cutoff,
                    int32_t* __restrict__ entries, unsigned long long entry_capacity, SegHdr* __restrict__ segs,
                    unsigned int seg_capacity, Counters* __restrict__ ctr, long long* __restrict__ dbg /* [nL][4] or null */,
                    const int32_t* __restrict__ owner_id /* null, or pre-sort index per slot */, int n_own) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    WarpSmem& S = reinterpret_cast<WarpSmem*>(smem_raw)[warp];

    const int A = blockIdx.x * TRAV_WARPS + warp;
    if (A >= nL) return;  // whole warp leaves; no block-wide barriers below

    const int ia = A * LEAF + lane;
    // multi-GPU: rows are built for OWNED atoms only (pre-sort index < n_own); ghosts are targets only
    const bool valid_i = ia < n && (owner_id == nullptr || owner_id[ia] < n_own);
    if (__ballot_sync(full, valid_i) == 0u) return;  // a leaf of ghosts: nothing to query
    const float4 pi = valid_i ? pos[ia] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float r2 = __fmul_rn(cutoff, cutoff);  // squared_radius = neighbor_distance^2 in Float32
    Region R;
    {
        const float4 alo4 = leaf_lo[A], ahi4 = leaf_hi[A];
        R.alo = make_float3(alo4.x, alo4.y, alo4.z);
        R.ahi = make_float3(ahi4.x, ahi4.y, ahi4.z);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float4 lo = leaf_sub[(size_t)A * 8 + 2 * r], hi = leaf_sub[(size_t)A * 8 + 2 * r + 1];
            R.slo[r] = make_float3(lo.x, lo.y, lo.z);
            R.shi[r] = make_float3(hi.x, hi.y, hi.z);
        }
        R.r2pad = fmaf(r2, 4e-6f, r2) + 1e-37f;
        const float w = 3.0f * cutoff;
        R.wide = (R.ahi.x - R.alo.x > w) || (R.ahi.y - R.alo.y > w) || (R.ahi.z - R.alo.z > w);
    }

    int cnt = 0;       // entries buffered in my row
    int sp = 0, ncand = 0, ntgt = 0, ntab = 0;
    int self_code = -1;  // entry code of my own atom once leaf A sits in the candidate table
    long long dbg_t0 = dbg ? clock64() : 0, dbg_cand = 0, dbg_rounds = 0, dbg_targets = 0;
    if (nL == 1) {
        if (lane == 0) S.cand[0] = 0;
        ncand = 1;
    } else {
        if (lane == 0) S.stack[0] = 0;  // root
        sp = 1;
    }
    __syncwarp(full);

    // ---- flush the buffered rows as one segment ------------------------------------------------------
    auto flush = [&]() {
        int total = cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(full, total, o);
        if (total > 0) {
            unsigned long long base = 0;
            unsigned int seg = 0;
            if (lane == 0) {
                base = atomicAdd(&ctr->n_entries, (unsigned long long)total);
                seg = atomicAdd(&ctr->n_segments, 1u);
            }
            base = __shfl_sync(full, base, 0);
            seg = __shfl_sync(full, seg, 0);
            const bool fits = (base + (unsigned long long)total <= entry_capacity);
            if (seg < seg_capacity) {
                SegHdr* H = &segs[seg];
                if (lane == 0) {
                    H->leaf = A;
                    H->total = fits ? (uint32_t)total : 0u;
                    H->base = base;
                }
                H->cnt[lane] = fits ? (uint8_t)cnt : (uint8_t)0;
            }
            if (!fits || seg >= seg_capacity) {
                if (lane == 0) {
                    atomicExch(&ctr->overflow, 1u);
                    atomicExch(&ctr->overflow_sticky, 1u);
                }
            } else {
                int maxc = cnt;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(full, maxc, o));
                unsigned long long off = base;
                for (int k = 0; k < maxc; ++k) {
                    bool act = k < cnt;
                    unsigned m = __ballot_sync(full, act);
                    if (act) {
                        unsigned code = S.rows[k * 32 + lane];
                        entries[off + __popc(m & lt_mask)] = S.ctab[code >> 5] * LEAF + (int)(code & 31u);
                    }
                    off += __popc(m);
                }
            }
        }
        cnt = 0;
        __syncwarp(full);
    };

    // ---- distance pass ---------------------------------------------------------------------------------
    // (1) the buffered targets passed the AABB test only: filter them against the sub-boxes, 32 per
    //     instruction, compacting in place;  (2) every query atom (lane) against every surviving target.
    auto test_targets = [&]() {
        __syncwarp(full);
        int kept = 0;
        for (int t0 = 0; t0 < ntgt; t0 += 32) {
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            bool ok = false;
            if (t0 + lane < ntgt) {
                q = S.tgt[t0 + lane];
                ok = R.near_sub(q, q);
            }
            unsigned m = __ballot_sync(full, ok);  // all lanes have read before anyone writes (kept <= t0)
            if (ok) S.tgt[kept + __popc(m & lt_mask)] = q;
            kept += __popc(m);
            __syncwarp(full);
        }
        dbg_targets += kept;
        if (lane < 4) S.tgt[kept + lane] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, __int_as_float(-2));
        __syncwarp(full);
        for (int t0 = 0; t0 < kept; t0 += CHUNK) {
            if (__any_sync(full, cnt > KMAX - CHUNK)) flush();  // a chunk adds at most CHUNK entries per lane
            const int tend = min(t0 + CHUNK, kept);
            for (int t = t0; t < tend; t += 4) {
                float4 q0 = S.tgt[t], q1 = S.tgt[t + 1], q2 = S.tgt[t + 2], q3 = S.tgt[t + 3];
                float d0 = dist2_exact(pi, q0), d1 = dist2_exact(pi, q1), d2 = dist2_exact(pi, q2), d3 = dist2_exact(pi, q3);
                int j0 = __float_as_int(q0.w), j1 = __float_as_int(q1.w), j2 = __float_as_int(q2.w), j3 = __float_as_int(q3.w);
                if (valid_i && d0 < r2 && j0 != self_code) { S.rows[cnt * 32 + lane] = (uint16_t)j0; ++cnt; }
                if (valid_i && d1 < r2 && j1 != self_code) { S.rows[cnt * 32 + lane] = (uint16_t)j1; ++cnt; }
                if (valid_i && d2 < r2 && j2 != self_code) { S.rows[cnt * 32 + lane] = (uint16_t)j2; ++cnt; }
                if (valid_i && d3 < r2 && j3 != self_code) { S.rows[cnt * 32 + lane] = (uint16_t)j3; ++cnt; }
            }
        }
        ntgt = 0;
        __syncwarp(full);
    };

    // ---- gather: lanes load the atoms of up to GATHER candidate leaves ------------------------------------
    auto load_batch = [&](int c0, float4 (&p)[GATHER], bool (&v)[GATHER]) {
#pragma unroll
        for (int u = 0; u < GATHER; ++u) {
            v[u] = false;
            p[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + u < ncand) {
                int jb = S.cand[c0 + u] * LEAF + lane;
                v[u] = jb < n;
                if (v[u]) p[u] = __ldg(&pos[jb]);
            }
        }
    };

    while (sp > 0 || ncand > 0 || ntgt > 0) {
        if (sp > 0) {
            // ---- one cooperative round of the tree walk ------------------------------------------------
            const int m = (sp > STACK_WIDE_LIMIT) ? 1 : min(sp, 32);
            ++dbg_rounds;
            const bool have = lane < m;
            int nd = have ? S.stack[sp - 1 - lane] : 0;
            __syncwarp(full);
            bool pushL = false, pushR = false, candL = false, candR = false;
            int left_id = 0, right_id = 0;
            if (have) {
                const float4* np = reinterpret_cast<const float4*>(&nodes[nd]);
                float4 c0 = __ldg(np), c1 = __ldg(np + 1), c2 = __ldg(np + 2), c3 = __ldg(np + 3);
                left_id = __float_as_int(c0.w);
                right_id = __float_as_int(c1.w);
                bool hitL = R.near_node(c0, c1);
                bool hitR = R.near_node(c2, c3);
                pushL = hitL && left_id >= 0;
                candL = hitL && left_id < 0;
                pushR = hitR && right_id >= 0;
                candR = hitR && right_id < 0;
            }
            unsigned bL = __ballot_sync(full, pushL), bR = __ballot_sync(full, pushR);
            unsigned cL = __ballot_sync(full, candL), cR = __ballot_sync(full, candR);
            const int newsp = sp - m;
            if (pushL) S.stack[newsp + __popc(bL & lt_mask)] = left_id;
            if (pushR) S.stack[newsp + __popc(bL) + __popc(bR & lt_mask)] = right_id;
            sp = newsp + __popc(bL) + __popc(bR);
            if (candL) S.cand[ncand + __popc(cL & lt_mask)] = ~left_id;
            if (candR) S.cand[ncand + __popc(cL) + __popc(cR & lt_mask)] = ~right_id;
            ncand += __popc(cL) + __popc(cR);
            __syncwarp(full);
        }
        if (ncand > 0) {
            // ---- gather the candidates' atoms; the next batch's loads fly while this one is compacted ----
            float4 pn[GATHER];
            bool vn[GATHER];
            load_batch(0, pn, vn);
            for (int c = 0; c < ncand; c += GATHER) {
                float4 pc[GATHER];
                bool vc[GATHER];
#pragma unroll
                for (int u = 0; u < GATHER; ++u) { pc[u] = pn[u]; vc[u] = vn[u]; }
                if (c + GATHER < ncand) load_batch(c + GATHER, pn, vn);
                if (ntab + GATHER > CTAB) {  // candidate table full: drain everything that refers to it
                    test_targets();
                    flush();
                    ntab = 0;
                    self_code = -1;
                }
#pragma unroll
                for (int u = 0; u < GATHER; ++u) {
                    bool near = vc[u] && R.near_aabb(pc[u], pc[u]);
                    unsigned msk = __ballot_sync(full, near);
                    if (msk) {  // warp-uniform
                        const int B = S.cand[c + u];
                        const int code = (ntab << 5) | lane;
                        if (lane == 0) S.ctab[ntab] = B;
                        if (B == A) self_code = code;
                        if (near) {
                            float4 t = pc[u];
                            t.w = __int_as_float(code);
                            S.tgt[ntgt + __popc(msk & lt_mask)] = t;
                        }
                        ntgt += __popc(msk);
                        ++ntab;
                    }
                }
                if (ntgt > TGT_CAP - GATHER * 32) test_targets();
            }
            dbg_cand += ncand;
            ncand = 0;
            __syncwarp(full);
        }
        if (sp == 0 && ntgt > 0) test_targets();
    }
    flush();
    if (dbg && lane == 0) {
        dbg[4 * A + 0] = clock64() - dbg_t0;
        dbg[4 * A + 1] = dbg_cand;
        dbg[4 * A + 2] = dbg_rounds;
        dbg[4 * A + 3] = dbg_targets;
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

__global__ void __launch_bounds__(256)
    export_kernel(const SegHdr* __restrict__ segs, const int32_t* __restrict__ entries, Counters* __restrict__ ctr,
                  unsigned int seg_capacity, const float4* __restrict__ pos, const int32_t* __restrict__ id, int n,
                  int32_t* __restrict__ out_a, int32_t* __restrict__ out_b, float* __restrict__ out_d,
                  unsigned long long capacity, int index_base) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned nseg = min(ctr->n_segments, seg_capacity);
    for (unsigned seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; seg < nseg; seg += nwarps) {
        const SegHdr* H = &segs[seg];
        if (H->total == 0) continue;
        const int ia = H->leaf * LEAF + lane;
        const int c = H->cnt[lane];
        const bool valid = ia < n;
        const float4 pi = valid ? pos[ia] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int ida = valid ? id[ia] : 0;
        const int ca = code10_ref(pi);
        int maxc = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(full, maxc, o));
        unsigned long long off = H->base;
        for (int k = 0; k < maxc; ++k) {
            bool act = k < c;
            unsigned m = __ballot_sync(full, act);
            int j = act ? entries[off + __popc(m & lt_mask)] : 0;
            off += __popc(m);
            bool keep = act && (j > ia);
            unsigned km = __ballot_sync(full, keep);
            if (km == 0) continue;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&ctr->n_export, (unsigned long long)__popc(km));
            base = __shfl_sync(full, base, 0);
            if (keep) {
                unsigned long long slot = base + __popc(km & lt_mask);
                if (slot < capacity) {
                    float4 pj = pos[j];
                    int idb = id[j];
                    int cb = code10_ref(pj);
                    bool a_first = (ca < cb) || (ca == cb && ida < idb);
                    out_a[slot] = (a_first ? ida : idb) + index_base;
                    out_b[slot] = (a_first ? idb : ida) + index_base;
                    out_d[slot] = __fsqrt_rn(dist2_exact(pi, pj));
                }
            }
        }
    }
}

// directed entries as (pre-sort index of the row atom, pre-sort index of the partner, d): the multi-GPU
// parity check unions these over the ranks
__global__ void __launch_bounds__(256)
    export_directed_kernel(const SegHdr* __restrict__ segs, const int32_t* __restrict__ entries, Counters* __restrict__ ctr,
                           unsigned int seg_capacity, const float4* __restrict__ pos, const int32_t* __restrict__ id, int n,
                           int32_t* __restrict__ out_a, int32_t* __restrict__ out_b, float* __restrict__ out_d,
                           unsigned long long capacity) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned nseg = min(ctr->n_segments, seg_capacity);
    for (unsigned seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; seg < nseg; seg += nwarps) {
        const SegHdr* H = &segs[seg];
        if (H->total == 0) continue;
        const int ia = H->leaf * LEAF + lane;
        const int c = H->cnt[lane];
        const bool valid = ia < n;
        const float4 pi = valid ? pos[ia] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int ida = valid ? id[ia] : 0;
        int maxc = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(full, maxc, o));
        unsigned long long off = H->base;
        unsigned long long obase = 0;
        if (lane == 0) obase = atomicAdd(&ctr->n_export, (unsigned long long)H->total);
        obase = __shfl_sync(full, obase, 0);
        for (int k = 0; k < maxc; ++k) {
            bool act = k < c;
            unsigned m = __ballot_sync(full, act);
            unsigned long long slot = obase + (off - H->base) + __popc(m & lt_mask);
            if (act && slot < capacity) {
                int j = entries[off + __popc(m & lt_mask)];
                out_a[slot] = ida;
                out_b[slot] = id[j];
                out_d[slot] = __fsqrt_rn(dist2_exact(pi, pos[j]));
            }
            off += __popc(m);
        }
    }
}

__global__ void __launch_bounds__(256)
    neighbor_counts_kernel(const SegHdr* __restrict__ segs, const Counters* __restrict__ ctr, unsigned int seg_capacity,
                           const int32_t* __restrict__ id, int n, int32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned nseg = min(ctr->n_segments, seg_capacity);
    for (unsigned seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; seg < nseg; seg += nwarps) {
        const SegHdr* H = &segs[seg];
        if (H->total == 0) continue;
        const int ia = H->leaf * LEAF + lane;
        const int c = H->cnt[lane];
        if (ia < n && c) atomicAdd(&counts[id[ia]], c);
    }
}

}  // namespace

int launch_traverse(cudaStream_t s, int sm_count, const Node* nodes, const float4* leaf_lo, const float4* leaf_hi,
                    const float4* leaf_sub, const float4* pos, int n, int n_leaves, float cutoff, int32_t* entries, int64_t entry_capacity,
                    SegHdr* segs, int64_t seg_capacity, Counters* counters, long long* dbg, const int32_t* owner_id,
                    int n_own) {
    (void)sm_count;
    const size_t smem = sizeof(WarpSmem) * TRAV_WARPS;
    cudaFuncSetAttribute(traverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemsetAsync(counters, 0, 16, s);  // n_entries, n_segments, overflow
    int blocks = (n_leaves + TRAV_WARPS - 1) / TRAV_WARPS;
    traverse_kernel<<<blocks, TRAV_WARPS * 32, smem, s>>>(nodes, leaf_lo, leaf_hi, leaf_sub, pos, n, n_leaves, cutoff, entries,
                                                          (unsigned long long)entry_capacity, segs,
                                                          (unsigned int)seg_capacity, counters, dbg, owner_id, n_own);
    return 1;
}

int launch_export(cudaStream_t s, int sm_count, const SegHdr* segs, const int32_t* entries, Counters* counters,
                  int64_t seg_capacity, const float4* pos, const int32_t* id, int n, int32_t* a, int32_t* b, float* d,
                  int64_t capacity, int index_base) {
    cudaMemsetAsync(&counters->n_export, 0, sizeof(unsigned long long), s);
    export_kernel<<<sm_count * 8, 256, 0, s>>>(segs, entries, counters, (unsigned int)seg_capacity, pos, id, n, a, b, d,
                                               (unsigned long long)capacity, index_base);
    return 1;
}

int launch_export_directed(cudaStream_t s, int sm_count, const SegHdr* segs, const int32_t* entries, Counters* counters,
                           int64_t seg_capacity, const float4* pos, const int32_t* id, int n, int32_t* a, int32_t* b, float* d,
                           int64_t capacity) {
    cudaMemsetAsync(&counters->n_export, 0, sizeof(unsigned long long), s);
    export_directed_kernel<<<sm_count * 8, 256, 0, s>>>(segs, entries, counters, (unsigned int)seg_capacity, pos, id, n, a, b, d,
                                                        (unsigned long long)capacity);
    return 1;
}

int launch_neighbor_counts(cudaStream_t s, int sm_count, const SegHdr* segs, const Counters* counters,
                           int64_t seg_capacity, const int32_t* id, int n, int32_t* counts) {
    cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n, s);
    neighbor_counts_kernel<<<sm_count * 4, 256, 0, s>>>(segs, counters, (unsigned int)seg_capacity, id, n, counts);
    return 1;
}

}  // namespace nb200
