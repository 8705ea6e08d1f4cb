// lbvh_build.cu — bottom-up LBVH construction with atomic AABB refit (Apetrei 2014 / ArborX style).
//
// Replaces delta/bvh_interior!/bounding_volume_hierarchy! (BVHTraverse.jl:659-860): same agglomerative
// scheme — every leaf thread climbs; at each split one hand-off word decides who continues (the
// reference uses an Atomix CAS on `store[split]`, :750,777; here an atomicExch + __threadfence,
// which the reference gets away without on x86-TSO, :762,792).  Differences by design:
//   * leaves hold 32 Morton-consecutive atoms (one warp's worth), boxes are tight (no r/2 padding —
//     the traversal pads the query instead);
//   * the tree stores both CHILD boxes inside the parent (64-B node) for the warp-cooperative
//     stack traversal, instead of per-node boxes + skip ropes;
//   * similarity metric delta(i) = (key_i xor key_{i+1}) with (i xor i+1) as tie-break, compared as
//     one 64-bit word — the same augmentation as :675 without the signed-overflow trick.
// Internal node numbering (Karras): a node covering leaves [l,r] is node r if it is a left child,
// node l if it is a right child; the root is node 0.  So the children of the node that splits at p
// are left = (l==p ? leaf p : node p), right = (r==p+1 ? leaf p+1 : node p+1).
#include <cuda/atomic>

#include "nb200_internal.cuh"

namespace nb200 {

namespace {

__device__ __forceinline__ uint64_t delta(int i, const float4* __restrict__ leaf_hi, int nL) {
    if (i < 0 || i >= nL - 1) return ~0ull;
    uint32_t a = __float_as_uint(leaf_hi[i].w), b = __float_as_uint(leaf_hi[i + 1].w);
    return ((uint64_t)(a ^ b) << 32) | (uint32_t)(i ^ (i + 1));
}

__device__ __forceinline__ float4 ld_cg(const float4* p) { return __ldcg(p); }

#ifndef NB200_BUILD_TPB
#define NB200_BUILD_TPB 128
#endif
// `off`: a second tree in the same arrays (multi-GPU ghost tree): the array pointers are already shifted by `off`, and
// `off` is added to every leaf / node id the nodes store, so the traversal addresses both trees through one base.
__global__ void __launch_bounds__(NB200_BUILD_TPB) build_kernel(const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                                                    int nL, Node* __restrict__ nodes, float4* node_lo, float4* node_hi,
                                                    int32_t* node_flag, int off) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nL) return;
    int l = i, r = i;
    float4 lo4 = leaf_lo[i], hi4 = leaf_hi[i];
    float3 lo = make_float3(lo4.x, lo4.y, lo4.z), hi = make_float3(hi4.x, hi4.y, hi4.z);
    uint64_t dl = delta(l - 1, leaf_hi, nL), dr = delta(r, leaf_hi, nL);

    while (true) {
        const bool is_left = dr < dl;  // ties (only at the root: both +inf) go right -> parent l-1
        const int p = is_left ? r : l - 1;
        if (p < 0) return;  // [0, nL-1]: this was the root
        // hand-off: leave my far range end; whoever arrives second finds the sibling's.  ONE acq_rel exchange
        // (atom.acq_rel.gpu): the release half publishes my node's box (written on the previous lap) before the
        // flag, the acquire half lets the second arriver read the sibling's — two full __threadfence()s around a
        // relaxed atomicExch did the same at twice the cost (membar was the top stall of this kernel).
        cuda::atomic_ref<int32_t, cuda::thread_scope_device> flag(node_flag[p]);
        int other = flag.exchange(is_left ? l : r, cuda::memory_order_acq_rel);
        if (other == -1) return;  // first arriver stops (BVHTraverse.jl:753-755)
        // second arriver: merge with the sibling, write the parent
        float3 slo, shi;
        int left_id, right_id;
        float3 llo, lhi, rlo, rhi;
        if (is_left) {
            r = other;  // sibling covers [p+1, r]
            bool sib_leaf = (r == p + 1);
            float4 a = sib_leaf ? ld_cg(&leaf_lo[p + 1]) : ld_cg(&node_lo[p + 1]);
            float4 b = sib_leaf ? ld_cg(&leaf_hi[p + 1]) : ld_cg(&node_hi[p + 1]);
            slo = make_float3(a.x, a.y, a.z); shi = make_float3(b.x, b.y, b.z);
            left_id = (l == p) ? ~(p + off) : p + off;
            right_id = sib_leaf ? ~(p + 1 + off) : (p + 1 + off);
            llo = lo; lhi = hi; rlo = slo; rhi = shi;
        } else {
            l = other;  // sibling covers [l, p]
            bool sib_leaf = (l == p);
            float4 a = sib_leaf ? ld_cg(&leaf_lo[p]) : ld_cg(&node_lo[p]);
            float4 b = sib_leaf ? ld_cg(&leaf_hi[p]) : ld_cg(&node_hi[p]);
            slo = make_float3(a.x, a.y, a.z); shi = make_float3(b.x, b.y, b.z);
            left_id = sib_leaf ? ~(p + off) : p + off;
            right_id = (r == p + 1) ? ~(p + 1 + off) : (p + 1 + off);
            llo = slo; lhi = shi; rlo = lo; rhi = hi;
        }
        lo = make_float3(fminf(lo.x, slo.x), fminf(lo.y, slo.y), fminf(lo.z, slo.z));
        hi = make_float3(fmaxf(hi.x, shi.x), fmaxf(hi.y, shi.y), fmaxf(hi.z, shi.z));
        dl = delta(l - 1, leaf_hi, nL);
        dr = delta(r, leaf_hi, nL);
        const int me = (dr < dl) ? r : l;  // my node number (root: l == 0)
        Node nd;
        nd.c[0] = make_float4(llo.x, llo.y, llo.z, __int_as_float(left_id));
        nd.c[1] = make_float4(lhi.x, lhi.y, lhi.z, __int_as_float(right_id));
        nd.c[2] = make_float4(rlo.x, rlo.y, rlo.z, __int_as_float(l + off));
        nd.c[3] = make_float4(rhi.x, rhi.y, rhi.z, __int_as_float(r + off));
        nodes[me] = nd;
        node_lo[me] = make_float4(lo.x, lo.y, lo.z, 0.f);
        node_hi[me] = make_float4(hi.x, hi.y, hi.z, 0.f);
    }
}

// The first levels of the tree are the same for every query leaf and are walked one level per round with 1, 2, 4, ...
// busy lanes.  This one-warp kernel expands the root level by level while the frontier fits a warp (<= 32 entries,
// internal nodes >= 0 or leaves ~id) and stores it; the traversal starts from it: its first round already tests 64
// child boxes with every lane busy, four to five rounds (and dependent L2 round trips) fewer per query leaf.
// fbox[2 i], fbox[2 i + 1]: the frontier entry's OWN box, so a query leaf only pushes the entries it is near.
__global__ void frontier_kernel(const Node* __restrict__ nodes, int nL, int32_t* __restrict__ frontier /* [0] = count, [1..32] */, int off,
                                const float4* __restrict__ node_lo, const float4* __restrict__ node_hi, const float4* __restrict__ leaf_lo,
                                const float4* __restrict__ leaf_hi, float4* __restrict__ fbox) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    int id = nL >= 2 ? off : ~off; // lane 0 holds the root (node `off`; a one-leaf tree: that leaf)
    int count = nL >= 1 ? 1 : 0;
    if (off == 0 && nL == 1) count = 0;  // the primary tree's only leaf is the query's own leaf: nothing to walk
    for (int level = 0; level < 8 && count > 0; ++level) {
        const bool have = lane < count;
        int l = 0, r = 0;
        bool internal = false;
        if (have && id >= 0) {
            const float4* np = reinterpret_cast<const float4*>(&nodes[id]);
            l = __float_as_int(np[0].w);
            r = __float_as_int(np[1].w);
            internal = true;
        }
        const int mine = have ? (internal ? 2 : 1) : 0;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(full, incl, 31);
        if (total > 32 || __ballot_sync(full, internal) == 0u) break;  // would not fit a warp / only leaves left
        // scatter the children (or the leaf itself) to their new lanes through shared memory
        __shared__ int32_t next[32];
        const int at = incl - mine;
        if (have) {
            if (internal) { next[at] = l; next[at + 1] = r; }
            else next[at] = id;
        }
        __syncwarp(full);
        id = lane < total ? next[lane] : 0;
        count = total;
        __syncwarp(full);
    }
    if (lane == 0) frontier[0] = count;
    if (lane < count) {
        frontier[1 + lane] = id;
        fbox[2 * lane] = id >= 0 ? node_lo[id] : leaf_lo[~id];
        fbox[2 * lane + 1] = id >= 0 ? node_hi[id] : leaf_hi[~id];
    }
}

}  // namespace

int launch_frontier(cudaStream_t s, const Node* nodes, int n_leaves, int32_t* frontier, int off, const float4* node_lo, const float4* node_hi,
                    const float4* leaf_lo, const float4* leaf_hi) {
    // the entry boxes live behind the 64 ints of the frontier list (same allocation: 64 ints + 64 float4)
    frontier_kernel<<<1, 32, 0, s>>>(nodes, n_leaves, frontier, off, node_lo, node_hi, leaf_lo, leaf_hi, reinterpret_cast<float4*>(frontier + 64));
    return 1;
}

int launch_build(cudaStream_t s, const float4* leaf_lo, const float4* leaf_hi, int n_leaves, Node* nodes, float4* node_lo,
                 float4* node_hi, int32_t* node_flag, bool flags_clean, int off) {
    if (n_leaves < 2) return 0;
    if (!flags_clean) cudaMemsetAsync(node_flag + off, 0xff, sizeof(int32_t) * (size_t)(n_leaves - 1), s);
    build_kernel<<<(n_leaves + NB200_BUILD_TPB - 1) / NB200_BUILD_TPB, NB200_BUILD_TPB, 0, s>>>(leaf_lo + off, leaf_hi + off, n_leaves, nodes + off,
                                                                                         node_lo + off, node_hi + off, node_flag + off, off);
    return 1;
}

// Tuning (NB200_CARVEOUT, see atoms.cu)
void carveout_build(int pct) {
    cudaFuncSetAttribute(build_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(frontier_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

}  // namespace nb200
