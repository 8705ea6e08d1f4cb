// nd_oracle.cpp — CPU ORACLE for the NaiveDynamics.jl MD hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.  The
// product path (libnaiveb200.so) never links, loads or calls anything in here.
//
// It is a C++17 restatement (not a copy: the reference is Julia) of the reference's
// algorithm for the path  LBVH neighbour search -> pair forces -> velocity Verlet.
// Every function cites the reference file:line it follows (paths relative to the
// reference checkout, gwenbiophys/NaiveDynamics.jl v0.0.4).
//
// Pinning status (see DESIGN.md "Oracle"):
//   * neighbour search: PINNED against the only literal fixture the reference's tests hold
//     (test/BVHTraverse.jl:165-187, `position8`, r=10 -> 28 pairs, skip-rope structure)
//     and against the property the reference tests assert (BVH == all-to-all,
//     test/BVHTraverse.jl:194-289) on seeded inputs.
//   * forces / Verlet: PARITY UNPINNED — the reference has no test, fixture or expected
//     value for Forces.jl or Simulator.jl.  The restatement follows the source line by line.
//
// Float semantics: build with -ffp-contract=off (Julia does not contract a*b+c) on x86-64
// SSE so every `float` op is one IEEE-754 binary32 rounding, like Julia's Float32.
//
// Index conventions: arrays are 0-based in memory; values that the reference stores as
// 1-based links/ids (left, skip, atom index) are kept 1-based so dumps compare directly
// with the Julia structures (0 == sentinel, as in BVHTraverse.jl:1276,1291).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
struct V3 { float x, y, z; };

inline V3 ld3(const float* p, int64_t i) { return V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

// sum((a - b) .^ 2) on SVector{3,Float32}: left fold (dx*dx + dy*dy) + dz*dz, no FMA
// (BVHTraverse.jl:1026,1045; StaticArrays `sum` is a left fold).
inline float dist2(const V3& a, const V3& b) {
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    float s = xx + yy;
    return s + zz;
}

template <class F>
void parallel_static(int nthreads, F&& fn) {
    if (nthreads <= 1) { fn(0); return; }
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&fn, t] { fn(t); });
    for (auto& t : th) t.join();
}

// GridKey{T,K} (BVHTraverse.jl:102-109): min, max, left, skip.  1-based links, 0 sentinel.
struct GridKey {
    float mn[3];
    float mx[3];
    int32_t left;
    int32_t skip;
};

struct PairList {
    std::vector<int32_t> a, b;
    std::vector<float> d;
    void push(int32_t i, int32_t j, float dd) { a.push_back(i); b.push_back(j); d.push_back(dd); }
    size_t size() const { return a.size(); }
};

// ---------------------------------------------------------------------------------------
// N1  mortoncodes!  (BVHTraverse.jl:237-288, live branch :259-284; binwidth :184)
// ---------------------------------------------------------------------------------------
const float kBinwidth = (float)(1.0 / 1023.0);  // const binwidth = Float32(1/1023)

inline int32_t morton_code_ref(const V3& p) {
    // magic_values / magic_not_values, :241-242
    const int32_t magic[3] = {153391689, 306783378, 613566756};
    const int32_t magic_not[3] = {-153391690, -306783379, -613566757};
    const float c[3] = {p.x, p.y, p.z};
    int32_t input = 0;
    for (int dim = 0; dim < 3; ++dim) {
        // round(Int32, quantized[each][dim] / binwidth, RoundDown)   (:266-267)
        int32_t q = (int32_t)std::floor(c[dim] / kBinwidth);
        int32_t yin = q & magic[dim];
        int32_t yang = q | magic_not[dim];
        input |= (yin & yang);  // :268-269
    }
    return input;
}

// ---------------------------------------------------------------------------------------
// Sorted SoA "APointPrimitive" (BVHTraverse.jl:142-146) after sortperm+permute! (:570-573)
// ---------------------------------------------------------------------------------------
struct Sorted {
    std::vector<int32_t> index;   // 1-based original atom id
    std::vector<int32_t> morton;
    std::vector<V3> pos;
};

Sorted make_sorted(const float* xyz, int32_t n) {
    Sorted s;
    s.index.resize(n); s.morton.resize(n); s.pos.resize(n);
    std::vector<int32_t> code(n);
    for (int32_t i = 0; i < n; ++i) code[i] = morton_code_ref(ld3(xyz, i));  // serial loop, :261
    std::vector<int32_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    // sortperm(pos.morton_code): Julia's default is a STABLE sort (:570)
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return code[a] < code[b]; });
    for (int32_t k = 0; k < n; ++k) {
        s.index[k] = perm[k] + 1;
        s.morton[k] = code[perm[k]];
        s.pos[k] = ld3(xyz, perm[k]);
    }
    return s;
}

// ---------------------------------------------------------------------------------------
// Spec (BVHTraverse.jl:63-94)
// ---------------------------------------------------------------------------------------
struct Spec {
    float r;
    int32_t atom_count, leaves_count, branches_count, apl;
};

// returns 0 ok, 1 = "evenly divides" error (:74-76), 2 = "more than one leaf" error (:78-80)
int make_spec(float r, int32_t n, int32_t apl, Spec* out) {
    if (apl <= 0 || n % apl != 0) return 1;
    double leaves = (double)n / (double)apl;  // :77 (Float64 division)
    if (leaves < 2) return 2;
    out->r = r; out->atom_count = n; out->apl = apl;
    out->leaves_count = (int32_t)leaves;
    out->branches_count = out->leaves_count - 1;
    return 0;
}

// ---------------------------------------------------------------------------------------
// N3  leafcluster_primitives (:376-410)  pad r/2     /   cluster_primitives (:341-375) pad r
// ---------------------------------------------------------------------------------------
void cluster_boxes(const Sorted& s, const Spec& sp, float pad, std::vector<GridKey>& keys) {
    for (int32_t each = 1; each <= sp.leaves_count; ++each) {
        int32_t bird = (each - 1) * sp.apl + 1;
        GridKey k;
        const V3& p0 = s.pos[bird - 1];
        k.mn[0] = p0.x - pad; k.mn[1] = p0.y - pad; k.mn[2] = p0.z - pad;
        k.mx[0] = p0.x + pad; k.mx[1] = p0.y + pad; k.mx[2] = p0.z + pad;
        k.left = 0; k.skip = 0;
        for (int32_t i = 1; i <= sp.apl; ++i) {
            int32_t a = (each - 1) * sp.apl + i;
            const V3& p = s.pos[a - 1];
            const float c[3] = {p.x, p.y, p.z};
            for (int d = 0; d < 3; ++d) {
                k.mn[d] = std::min(k.mn[d], c[d] - pad);
                k.mx[d] = std::max(k.mx[d], c[d] + pad);
            }
        }
        keys[each - 1] = k;
    }
}

// ---------------------------------------------------------------------------------------
// N4  delta / exptdelta (:659-696), branch_index (:699-701), bvh_interior! (:704-846),
//     bounding_volume_hierarchy! (:849-860)
// ---------------------------------------------------------------------------------------
inline int32_t delta(int32_t i, const Sorted& s, const Spec& sp) {
    if (i >= sp.leaves_count || i < 1) return std::numeric_limits<int32_t>::max();  // :680-682
    int32_t atom_i = 1 + (i - 1) * sp.apl;
    int32_t atom_i_and1 = 1 + i * sp.apl;
    int32_t x = s.morton[atom_i - 1] ^ s.morton[atom_i_and1 - 1];
    // x + (x == 0) * (typemin(K) + (i xor (i+1))) - 1   with wrapping Int32 arithmetic (:693)
    uint32_t ux = (uint32_t)x;
    uint32_t aug = (x == 0) ? ((uint32_t)std::numeric_limits<int32_t>::min() + (uint32_t)(i ^ (i + 1))) : 0u;
    return (int32_t)(ux + aug - 1u);
}

inline int32_t branch_index(int32_t a, const Spec& sp) { return a + sp.leaves_count; }

void bvh_interior(std::vector<GridKey>& keys, std::vector<std::atomic<int32_t>>& store, int32_t i,
                  const Spec& sp, const Sorted& s) {
    const int32_t nL = sp.leaves_count;
    int32_t rangel = i, ranger = i;
    int32_t dell = delta(rangel - 1, s, sp);
    int32_t delr = delta(ranger, s, sp);
    float bmin[3], bmax[3];
    for (int d = 0; d < 3; ++d) { bmin[d] = keys[i - 1].mn[d]; bmax[d] = keys[i - 1].mx[d]; }

    // leaf skip rope (:717-741)
    if (i == nL) {
        keys[i - 1].skip = 0;
    } else {
        int32_t ir = i + 1;
        if (delr < delta(ir, s, sp)) keys[i - 1].skip = ir;
        else keys[i - 1].skip = branch_index(ir, sp);
    }

    while (true) {  // :743
        bool isLeftChild = delr < dell;
        int32_t leftChild;
        if (isLeftChild) {
            leftChild = i;
            int32_t split = ranger;
            int32_t expected = 0;  // Atomix.@atomicreplace store[split] 0 => rangel  (:750)
            bool ok = store[split - 1].compare_exchange_strong(expected, rangel);
            ranger = ok ? 0 : expected;
            if (ranger == 0) break;  // first arriver is culled (:753-755)
            delr = delta(ranger, s, sp);
            int32_t rightChild = split + 1;
            bool rightChildIsLeaf = (rightChild == ranger);
            if (!rightChildIsLeaf) rightChild = branch_index(rightChild, sp);
            for (int d = 0; d < 3; ++d) {
                bmax[d] = std::max(keys[rightChild - 1].mx[d], bmax[d]);
                bmin[d] = std::min(keys[rightChild - 1].mn[d], bmin[d]);
            }
        } else {
            int32_t split = rangel - 1;
            int32_t expected = 0;  // :777
            bool ok = store[split - 1].compare_exchange_strong(expected, ranger);
            rangel = ok ? 0 : expected;
            if (rangel == 0) break;
            dell = delta(rangel - 1, s, sp);
            leftChild = split;
            bool leftChildIsLeaf = (leftChild == rangel);
            if (!leftChildIsLeaf) leftChild = branch_index(leftChild, sp);
            for (int d = 0; d < 3; ++d) {
                bmax[d] = std::max(keys[leftChild - 1].mx[d], bmax[d]);
                bmin[d] = std::min(keys[leftChild - 1].mn[d], bmin[d]);
            }
        }
        int32_t q = (delr < dell) ? ranger : rangel;  // :806
        int32_t parentNode = branch_index(q, sp);
        GridKey& P = keys[parentNode - 1];
        P.left = leftChild;  // :812
        int32_t skip;
        if (ranger == nL) {
            skip = 0;  // :815-819
        } else {
            int32_t r = ranger + 1;
            if (delr < delta(r, s, sp)) skip = r;  // :822-825
            else skip = branch_index(r, sp);       // :827-830
        }
        for (int d = 0; d < 3; ++d) { P.mn[d] = bmin[d]; P.mx[d] = bmax[d]; }
        P.skip = skip;
        i = parentNode;  // :839
        if (i == branch_index(1, sp)) return;  // root (:842)
    }
}

void bounding_volume_hierarchy(std::vector<GridKey>& keys, const Spec& sp, const Sorted& s, int nthreads) {
    std::vector<std::atomic<int32_t>> store(sp.branches_count);
    for (auto& a : store) a.store(0, std::memory_order_relaxed);
    const int32_t nL = sp.leaves_count;
    // Polyester @batch: contiguous static chunks over 1:leaves_count (:851)
    parallel_static(nthreads, [&](int t) {
        int64_t lo = (int64_t)nL * t / nthreads, hi = (int64_t)nL * (t + 1) / nthreads;
        for (int64_t i = lo + 1; i <= hi; ++i) bvh_interior(keys, store, (int32_t)i, sp, s);
    });
    // root AABB forced to [0,1]^3 (:856-858)
    GridKey& R = keys[branch_index(1, sp) - 1];
    for (int d = 0; d < 3; ++d) { R.mn[d] = 0.0f; R.mx[d] = 1.0f; }
}

// ---------------------------------------------------------------------------------------
// leafTreeData (:544-597) — the LIVE path:  boxes padded by r/2, atom_count-1 nodes allocated
// TreeData (:500-543)     — atom-query path: boxes padded by r, leaves+branches nodes
// ---------------------------------------------------------------------------------------
struct Tree {
    Spec sp;
    Sorted s;
    std::vector<GridKey> keys;
};

int build_tree(const float* xyz, int32_t n, float r, int32_t apl, bool leaf_variant, int nthreads, Tree* T) {
    int rc = make_spec(r, n, apl, &T->sp);
    if (rc) return rc;
    const Spec& sp = T->sp;
    T->s = make_sorted(xyz, n);
    GridKey zero; std::memset(&zero, 0, sizeof(zero));
    if (leaf_variant) {
        // leaves = [GridKey(0..) for i in 1:spec.atom_count-1]  (:378) — needs 2*nL-1 <= N-1, i.e. apl >= 2
        if (2 * (int64_t)sp.leaves_count - 1 > (int64_t)n - 1) return 3;  // Julia would go out of bounds
        T->keys.assign(n - 1, zero);
        cluster_boxes(T->s, sp, sp.r / 2, T->keys);  // spec.neighbor_distance/2 in Float32 (:386)
    } else {
        T->keys.assign(sp.leaves_count + sp.branches_count, zero);  // L ++ I (:528-535)
        cluster_boxes(T->s, sp, sp.r, T->keys);
    }
    bounding_volume_hierarchy(T->keys, sp, T->s, nthreads);
    return 0;
}

// ---------------------------------------------------------------------------------------
// N5  aabb_overlap_test (:1094-1098), onecluster_/twocluster_proximitytest! (:1021-1055)
// ---------------------------------------------------------------------------------------
inline bool aabb_overlap(const GridKey& A, const GridKey& B) {
    bool all = true;
    for (int d = 0; d < 3; ++d) all = all & ((A.mn[d] < B.mx[d]) & (A.mx[d] > B.mn[d]));
    return all;
}

inline void onecluster(PairList& out, const Sorted& s, int32_t low, const Spec& sp, float r2) {
    for (int32_t i = 1; i <= sp.apl - 1; ++i)
        for (int32_t j = i + 1; j <= sp.apl; ++j) {
            float d2 = dist2(s.pos[low + i - 2], s.pos[low + j - 2]);
            if (d2 < r2) out.push(s.index[low + i - 2], s.index[low + j - 2], std::sqrt(d2));
        }
}

inline void twocluster(PairList& out, const Sorted& s, int32_t lowA, int32_t lowB, const Spec& sp, float r2) {
    for (int32_t i = 0; i < sp.apl; ++i)
        for (int32_t j = 0; j < sp.apl; ++j) {
            float d2 = dist2(s.pos[lowA - 1 + i], s.pos[lowB - 1 + j]);
            if (d2 < r2) out.push(s.index[lowA - 1 + i], s.index[lowB - 1 + j], std::sqrt(d2));
        }
}

// ---------------------------------------------------------------------------------------
// N6  leafneighbor_traverse (:1236-1323).  `stride`/`offset` let the bench traverse a
//     sample of the query leaves (offset, offset+stride, ...) — with stride=1 it is the
//     full reference loop.  Thread `chunk` takes query leaves chunk, chunk+T, ... (:1255-1256).
// ---------------------------------------------------------------------------------------
void leafneighbor_traverse(const Tree& T, int nthreads, int32_t qstride, PairList& result) {
    const Spec& sp = T.sp;
    const float r2 = sp.r * sp.r;  // squared_radius = neighbor_distance ^ 2 in Float32 (:1248)
    std::vector<PairList> vec(nthreads);  // parallel_neighbor_buffer (:1326-1329)
    parallel_static(nthreads, [&](int t) {
        PairList& out = vec[t];
        int32_t chunk = t + 1;
        for (int64_t qq = chunk; qq <= sp.leaves_count; qq += (int64_t)nthreads) {
            int32_t query_index = (int32_t)qq;
            if (qstride > 1 && (query_index - 1) % qstride != 0) continue;  // bench sampling only
            const GridKey& query_leaf = T.keys[query_index - 1];
            int32_t low = (query_index - 1) * sp.apl + 1;
            int32_t target_index = query_leaf.skip;  // :1267
            onecluster(out, T.s, low, sp, r2);       // :1270
            while (target_index != 0) {              // :1276
                const GridKey& target = T.keys[target_index - 1];
                bool overlap = aabb_overlap(query_leaf, target);
                if (overlap) {
                    if (target.left == 0) {  // leaf (:1291)
                        int32_t lowB = (target_index - 1) * sp.apl + 1;
                        twocluster(out, T.s, low, lowB, sp, r2);
                        target_index = target.skip;
                    } else {
                        target_index = target.left;
                    }
                } else {
                    target_index = target.skip;
                }
            }
        }
    });
    // reduce(vcat, neighbor_vec) (:1314) — serial concat in thread order
    size_t total = 0;
    for (auto& v : vec) total += v.size();
    result.a.reserve(total); result.b.reserve(total); result.d.reserve(total);
    for (auto& v : vec) {
        result.a.insert(result.a.end(), v.a.begin(), v.a.end());
        result.b.insert(result.b.end(), v.b.begin(), v.b.end());
        result.d.insert(result.d.end(), v.d.begin(), v.d.end());
    }
}

// ---------------------------------------------------------------------------------------
// N6'  neighbor_traverse (:1125-1183) + proximity_test! (:1057-1086) + overlap_test (:1089-1092).
//      Atom-query variant.  At HEAD it indexes an APointPrimitive struct directly (no getindex
//      exists) so it cannot run; restated as the indexing obviously intends
//      (positions[query_index] == the query_index-th sorted atom).
// ---------------------------------------------------------------------------------------
void neighbor_traverse(const Tree& T, int nthreads, PairList& result) {
    const Spec& sp = T.sp;
    const float r2 = sp.r * sp.r;  // :1136
    std::vector<PairList> vec(nthreads);
    const int32_t n = sp.atom_count;
    parallel_static(nthreads, [&](int t) {
        PairList& out = vec[t];
        for (int64_t qq = t + 1; qq <= n - 1; qq += nthreads) {  // chunk:threads:length-1 (:1143)
            int32_t query_index = (int32_t)qq;
            // round(K, query_index / atomsperleaf, RoundUp) (:1144)
            int32_t currentKey = (int32_t)std::ceil((double)query_index / (double)sp.apl);
            const V3& qp = T.s.pos[query_index - 1];
            const int32_t qid = T.s.index[query_index - 1];
            while (currentKey != 0) {
                const GridKey& k = T.keys[currentKey - 1];
                // all(myKey.min .< myPos.position .< myKey.max) (:1090)
                bool overlap = (k.mn[0] < qp.x) & (qp.x < k.mx[0]) & (k.mn[1] < qp.y) & (qp.y < k.mx[1]) &
                               (k.mn[2] < qp.z) & (qp.z < k.mx[2]);
                if (overlap) {
                    if (k.left == 0) {
                        int32_t low = (currentKey - 1) * sp.apl + 1;
                        for (int32_t each = 1; each <= sp.apl; ++each) {
                            int32_t subj = low + each - 1;  // sorted position of the subject
                            // query_index < each+low && query.index != subject.index (:1062)
                            if (query_index < each + low && qid != T.s.index[subj - 1]) {
                                float d2 = dist2(qp, T.s.pos[subj - 1]);
                                if (d2 < r2) out.push(qid, T.s.index[subj - 1], std::sqrt(d2));
                            }
                        }
                        currentKey = k.skip;
                    } else {
                        currentKey = k.left;
                    }
                } else {
                    currentKey = k.skip;
                }
            }
        }
    });
    for (auto& v : vec) {
        result.a.insert(result.a.end(), v.a.begin(), v.a.end());
        result.b.insert(result.b.end(), v.b.begin(), v.b.end());
        result.d.insert(result.d.end(), v.d.begin(), v.d.end());
    }
}

// ---------------------------------------------------------------------------------------
// O1  unique_pairs / update_pairslist! / threshold_pairs (AllToAll.jl:7-42,43-63,65-88)
//     predicate: sqrt(sum((a_i - a_j).^2)) < threshold   (NOT d2 < r2)
//     and the N5-predicate brute force (d2 < fl(r*r)) that is the authoritative pair set.
// ---------------------------------------------------------------------------------------
void brute_force(const float* xyz, int32_t n, float r, bool sqrt_predicate, int nthreads, PairList& result) {
    const float r2 = r * r;
    std::vector<PairList> vec(nthreads);
    parallel_static(nthreads, [&](int t) {
        PairList& out = vec[t];
        // contiguous i-chunks balanced by pair count so the concatenation stays (i,j)-ordered
        double tot = (double)n * (n - 1) / 2;
        auto bound = [&](int tt) {
            double target = tot * tt / nthreads;  // pairs before row i: i*n - i(i+1)/2
            int64_t lo = 0, hi = n;
            while (lo < hi) {
                int64_t mid = (lo + hi) / 2;
                double before = (double)mid * n - (double)mid * (mid + 1) / 2;
                if (before < target) lo = mid + 1; else hi = mid;
            }
            return (int32_t)lo;
        };
        int32_t i0 = t == 0 ? 0 : bound(t), i1 = t == nthreads - 1 ? n : bound(t + 1);
        for (int32_t i = i0; i < i1; ++i) {
            V3 a = ld3(xyz, i);
            for (int32_t j = i + 1; j < n; ++j) {
                float d2 = dist2(a, ld3(xyz, j));
                float d = std::sqrt(d2);
                bool keep = sqrt_predicate ? (d < r) : (d2 < r2);
                if (keep) out.push(i + 1, j + 1, d);
            }
        }
    });
    for (auto& v : vec) {
        result.a.insert(result.a.end(), v.a.begin(), v.a.end());
        result.b.insert(result.b.end(), v.b.begin(), v.b.end());
        result.d.insert(result.d.end(), v.d.begin(), v.d.end());
    }
}

// ---------------------------------------------------------------------------------------
// Independent O(N) exact search used ONLY to check full-size (1M) runs: a uniform cell grid
// with cell edge >= r * (1 + 1e-5), all 27 neighbour cells, exact N5 predicate.  It shares no
// code or structure with the BVH paths (oracle or CUDA).  Produces order-independent digests.
// ---------------------------------------------------------------------------------------
struct Digest {
    int64_t count;
    uint64_t xor_hash;   // xor over pairs of mix(min_id, max_id, bits(d))
    uint64_t sum_hash;   // wrapping sum of the same
    double sum_d;        // sum of d (double accumulate, order-dependent in last bits only)
};

inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
inline uint64_t pair_hash(int32_t a, int32_t b, float d) {
    uint32_t lo = (uint32_t)std::min(a, b), hi = (uint32_t)std::max(a, b);
    uint32_t db; std::memcpy(&db, &d, 4);
    return mix64(((uint64_t)lo << 32 | hi) * 0x9E3779B97F4A7C15ULL + db);
}

void cellgrid_digest(const float* xyz, int32_t n, float r, int nthreads, Digest* out, int32_t* per_atom_count) {
    const float r2 = r * r;
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int32_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], xyz[3 * i + d]); hi[d] = std::max(hi[d], xyz[3 * i + d]); }
    double edge = (double)r * (1.0 + 1e-5) + 1e-30;
    int64_t dim[3];
    for (int d = 0; d < 3; ++d) {
        double ext = (double)hi[d] - (double)lo[d];
        dim[d] = std::max<int64_t>(1, std::min<int64_t>(1024, (int64_t)std::floor(ext / edge)));
    }
    auto cell_of = [&](int32_t i, int d) {
        double ext = (double)hi[d] - (double)lo[d];
        if (ext <= 0) return (int64_t)0;
        int64_t c = (int64_t)std::floor(((double)xyz[3 * i + d] - (double)lo[d]) / ext * (double)dim[d]);
        return std::max<int64_t>(0, std::min<int64_t>(dim[d] - 1, c));
    };
    int64_t ncell = dim[0] * dim[1] * dim[2];
    std::vector<int32_t> start(ncell + 1, 0), order(n);
    std::vector<int64_t> cid(n);
    for (int32_t i = 0; i < n; ++i) { cid[i] = (cell_of(i, 2) * dim[1] + cell_of(i, 1)) * dim[0] + cell_of(i, 0); start[cid[i] + 1]++; }
    for (int64_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    { std::vector<int32_t> cur(start.begin(), start.end() - 1); for (int32_t i = 0; i < n; ++i) order[cur[cid[i]]++] = i; }
    std::vector<Digest> part(nthreads, Digest{0, 0, 0, 0.0});
    if (per_atom_count) std::memset(per_atom_count, 0, sizeof(int32_t) * (size_t)n);
    parallel_static(nthreads, [&](int t) {
        Digest& D = part[t];
        for (int64_t c = t; c < ncell; c += nthreads) {
            int64_t cx = c % dim[0], cy = (c / dim[0]) % dim[1], cz = c / (dim[0] * dim[1]);
            for (int32_t ii = start[c]; ii < start[c + 1]; ++ii) {
                int32_t i = order[ii];
                V3 a = ld3(xyz, i);
                int32_t cnt = 0;
                for (int64_t dz = -1; dz <= 1; ++dz) for (int64_t dy = -1; dy <= 1; ++dy) for (int64_t dx = -1; dx <= 1; ++dx) {
                    int64_t nx = cx + dx, ny = cy + dy, nz = cz + dz;
                    if (nx < 0 || ny < 0 || nz < 0 || nx >= dim[0] || ny >= dim[1] || nz >= dim[2]) continue;
                    int64_t nc = (nz * dim[1] + ny) * dim[0] + nx;
                    for (int32_t jj = start[nc]; jj < start[nc + 1]; ++jj) {
                        int32_t j = order[jj];
                        if (j == i) continue;
                        float d2 = dist2(a, ld3(xyz, j));
                        if (d2 < r2) {
                            ++cnt;
                            if (i < j) {  // each unordered pair digested once
                                float d = std::sqrt(d2);
                                uint64_t h = pair_hash(i + 1, j + 1, d);
                                D.count++; D.xor_hash ^= h; D.sum_hash += h; D.sum_d += d;
                            }
                        }
                    }
                }
                if (per_atom_count) per_atom_count[i] = cnt;  // full (two-sided) neighbour count
            }
        }
    });
    *out = Digest{0, 0, 0, 0.0};
    for (auto& p : part) { out->count += p.count; out->xor_hash ^= p.xor_hash; out->sum_hash += p.sum_hash; out->sum_d += p.sum_d; }
}

// ---------------------------------------------------------------------------------------
// F1  force_lennardjones! + lennardjones_interior (Forces.jl:6-45)   — literal
//     eps = -1f10 (Float32), sigma = 0.0001 (Float64) => the expression evaluates in Float64
//     and is rounded to Float32 by the in-place broadcast into force[i].
// ---------------------------------------------------------------------------------------
void force_lennardjones_literal(float* force, int32_t n, const int32_t* pa, const int32_t* pb, const float* pd, int64_t np) {
    (void)pb;
    const float eps = -1e10f;
    const double sigma = 0.0001;
    for (int64_t i = 0; i < 3 * (int64_t)n; ++i) force[i] = 0.0f;  // :23-25
    if (np < 1) return;                                             // :34-36
    for (int64_t e = 0; e < np; ++e) {
        int32_t i = pa[e];
        float dxyz = pd[e];
        // (24*eps ./ dxyz) .* ((2*sigma ./ dxyz) .^ T(12.0) .- (sigma ./ dxyz) .^ T(6.0))   (:12)
        float t24 = 24 * eps;                 // Int*Float32 -> Float32
        float a = t24 / dxyz;                 // Float32
        double b = std::pow((2 * sigma) / (double)dxyz, (double)12.0f) - std::pow(sigma / (double)dxyz, (double)6.0f);
        double v = (double)a * b;
        for (int d = 0; d < 3; ++d) {
            float* f = &force[3 * (int64_t)(i - 1) + d];
            *f = (float)((double)*f + v);     // force[i] .+= v  (Float32 + Float64 -> Float64 -> store Float32)
        }
    }
}

// ---------------------------------------------------------------------------------------
// F2  force_coulomb! + coulomb_interior (Forces.jl:46-66) — literal, ORDER DEPENDENT
//     force[i] .+= k*q_i*q_j ./ d^2 ;  force[j] .-= force[i]  (the whole accumulated force[i])
// ---------------------------------------------------------------------------------------
void force_coulomb_literal(float* force, int32_t n, const int32_t* pa, const int32_t* pb, const float* pd, int64_t np,
                           const float* charge) {
    for (int64_t i = 0; i < 3 * (int64_t)n; ++i) force[i] = 0.0f;
    for (int64_t e = 0; e < np; ++e) {
        int32_t i = pa[e], j = pb[e];
        float dxyz = pd[e];
        float num = 1 * charge[i - 1] * charge[j - 1];  // k = 1 (Int) * Float32 * Float32
        float den = dxyz * dxyz;                        // dxyz .^ 2 -> literal_pow -> x*x
        float v = num / den;
        for (int d = 0; d < 3; ++d) force[3 * (int64_t)(i - 1) + d] += v;
        for (int d = 0; d < 3; ++d) force[3 * (int64_t)(j - 1) + d] -= force[3 * (int64_t)(i - 1) + d];
    }
}

// F3  sum_forces! (Forces.jl:68-75)
void sum_forces(float* f, const float* f1, const float* f2, int64_t n3) {
    for (int64_t i = 0; i < n3; ++i) f[i] = f1[i] + f2[i];
}

// ---------------------------------------------------------------------------------------
// Physical LJ 12-6 + Coulomb over a half pair list, Float64, Newton's third law.
// This is the fp64 oracle the CUDA force kernel is checked against (<= 1e-5 relative).
//   F_ij = [24 eps (2 (s/r)^12 - (s/r)^6)/r^2 + kc q_i q_j / r^3] * (r_i - r_j)
//   U    = 4 eps ((s/r)^12 - (s/r)^6) - U_lj(rc)*shift  +  kc q_i q_j (1/r - shift/rc)
// `scale[i]` receives sum_j |F_ij| (conditioning scale for the tolerance).
// ---------------------------------------------------------------------------------------
void forces_physical_f64(const float* xyz, const float* charge, int32_t n, const int32_t* pa, const int32_t* pb, int64_t np,
                         double eps, double sigma, double kc, double rc, int shift, double* force, double* pe_atom,
                         double* scale) {
    for (int64_t i = 0; i < 3 * (int64_t)n; ++i) force[i] = 0.0;
    for (int32_t i = 0; i < n; ++i) { pe_atom[i] = 0.0; if (scale) scale[i] = 0.0; }
    double src6 = std::pow(sigma / rc, 6.0);
    double ulj_rc = shift ? 4.0 * eps * (src6 * src6 - src6) : 0.0;
    for (int64_t e = 0; e < np; ++e) {
        int32_t i = pa[e] - 1, j = pb[e] - 1;
        double dx = (double)xyz[3 * i] - (double)xyz[3 * j];
        double dy = (double)xyz[3 * i + 1] - (double)xyz[3 * j + 1];
        double dz = (double)xyz[3 * i + 2] - (double)xyz[3 * j + 2];
        double r2 = dx * dx + dy * dy + dz * dz;
        double r = std::sqrt(r2);
        double s2 = sigma * sigma / r2, s6 = s2 * s2 * s2;
        double fs = 24.0 * eps * (2.0 * s6 * s6 - s6) / r2;
        double u = 4.0 * eps * (s6 * s6 - s6) - ulj_rc;
        if (charge) {
            double qq = kc * (double)charge[i] * (double)charge[j];
            fs += qq / (r2 * r);
            u += qq * (1.0 / r - (shift ? 1.0 / rc : 0.0));
        }
        force[3 * i] += fs * dx; force[3 * i + 1] += fs * dy; force[3 * i + 2] += fs * dz;
        force[3 * j] -= fs * dx; force[3 * j + 1] -= fs * dy; force[3 * j + 2] -= fs * dz;
        pe_atom[i] += 0.5 * u; pe_atom[j] += 0.5 * u;
        if (scale) { double m = std::fabs(fs) * r; scale[i] += m; scale[j] += m; }
    }
}

// ---------------------------------------------------------------------------------------
// V1  velocity-Verlet body (Simulator.jl:198-222)  — literal, Float32, four separate loops
// V2  boundary_reflect! (Simulator.jl:81-111)
// ---------------------------------------------------------------------------------------
void verlet_literal(float* pos, float* vel, const float* force, const float* force_next, const float* mass, int32_t n,
                    float dt) {
    std::vector<float> a_t(3 * (size_t)n), a_tdt(3 * (size_t)n);
    for (int32_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) a_t[3 * i + d] = force[3 * i + d] / mass[i];  // :198-200
    float dt2 = dt * dt;  // spec.stepwidth ^ 2
    for (int32_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) {
        float t1 = vel[3 * i + d] * dt;
        float t2 = (a_t[3 * i + d] * dt2) / 2;
        float s = t1 + t2;
        pos[3 * i + d] = pos[3 * i + d] + s;  // :202-204
    }
    for (int32_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) a_tdt[3 * i + d] = force_next[3 * i + d] / mass[i];  // :206-208
    for (int32_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) {
        float s = a_t[3 * i + d] + a_tdt[3 * i + d];
        float t = (s * dt) / 2;  // (a_t .+ a_t_dt) .* stepwidth / 2   — left-assoc: ((..)*dt)/2
        vel[3 * i + d] = vel[3 * i + d] + t;  // :220-222
    }
}

void boundary_reflect(float* pos, float* vel, int32_t n, const float* mn, const float* mx) {
    for (int32_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            if (mn[d] > pos[3 * i + d]) { vel[3 * i + d] *= -1; pos[3 * i + d] = mn[d]; }
            if (mx[d] < pos[3 * i + d]) { vel[3 * i + d] *= -1; pos[3 * i + d] = mx[d]; }
        }
}

// ---------------------------------------------------------------------------------------
// Float64 kick-drift-kick velocity Verlet with reflective walls over the brute-force / cell
// pair set — the integrator oracle for the CUDA step loop (same formulation, fp64).
// One call = one step:  v += F/m dt/2 ; x += v dt ; reflect ; F = F(x) ; v += F/m dt/2.
// ---------------------------------------------------------------------------------------
void cell_pairs(const float* xyz, int32_t n, float r, std::vector<int32_t>& pa, std::vector<int32_t>& pb);

}  // namespace

// =========================================================================================
// C interface (ctypes)
// =========================================================================================
extern "C" {

struct nd_pairs {
    PairList list;
    double seconds_build;
    double seconds_traverse;
};

int32_t nd_hardware_threads() { return (int32_t)std::max(1u, std::thread::hardware_concurrency()); }

void nd_mortoncodes(const float* xyz, int32_t n, int32_t* codes) {
    for (int32_t i = 0; i < n; ++i) codes[i] = morton_code_ref(ld3(xyz, i));
}

// stable sortperm, returns 1-based permutation like Julia
void nd_sortperm(const int32_t* codes, int32_t n, int32_t* perm1) {
    std::vector<int32_t> p(n);
    std::iota(p.begin(), p.end(), 0);
    std::stable_sort(p.begin(), p.end(), [&](int32_t a, int32_t b) { return codes[a] < codes[b]; });
    for (int32_t i = 0; i < n; ++i) perm1[i] = p[i] + 1;
}

int32_t nd_spec(float r, int32_t n, int32_t apl, int32_t* leaves, int32_t* branches) {
    Spec sp;
    int rc = make_spec(r, n, apl, &sp);
    if (!rc) { *leaves = sp.leaves_count; *branches = sp.branches_count; }
    return rc;
}

// Dump of leafTreeData / TreeData: returns number of nodes written (or -rc on error).
// node arrays must hold max(n-1, 2*leaves-1) entries.
int32_t nd_tree(const float* xyz, int32_t n, float r, int32_t apl, int32_t leaf_variant, int32_t nthreads, float* node_min,
                float* node_max, int32_t* left, int32_t* skip, int32_t* sorted_index, int32_t* sorted_code,
                float* sorted_pos) {
    Tree T;
    int rc = build_tree(xyz, n, r, apl, leaf_variant != 0, nthreads, &T);
    if (rc) return -rc;
    int32_t nn = (int32_t)T.keys.size();
    for (int32_t k = 0; k < nn; ++k) {
        for (int d = 0; d < 3; ++d) { node_min[3 * k + d] = T.keys[k].mn[d]; node_max[3 * k + d] = T.keys[k].mx[d]; }
        left[k] = T.keys[k].left; skip[k] = T.keys[k].skip;
    }
    for (int32_t k = 0; k < n; ++k) {
        sorted_index[k] = T.s.index[k]; sorted_code[k] = T.s.morton[k];
        sorted_pos[3 * k] = T.s.pos[k].x; sorted_pos[3 * k + 1] = T.s.pos[k].y; sorted_pos[3 * k + 2] = T.s.pos[k].z;
    }
    return nn;
}

// leafbuild_traverse_bvh (:1423-1428) when leaf_variant, build_traverse_bvh (:1416-1421) otherwise.
// qstride > 1 traverses only every qstride-th query leaf (bench sampling; leaf variant only).
nd_pairs* nd_build_traverse(const float* xyz, int32_t n, float r, int32_t apl, int32_t leaf_variant, int32_t nthreads,
                            int32_t qstride, int32_t* rc_out) {
    auto* P = new nd_pairs();
    Tree T;
    auto t0 = std::chrono::steady_clock::now();
    int rc = build_tree(xyz, n, r, apl, leaf_variant != 0, nthreads, &T);
    auto t1 = std::chrono::steady_clock::now();
    if (rc_out) *rc_out = rc;
    if (rc) return P;
    if (leaf_variant) leafneighbor_traverse(T, nthreads, qstride, P->list);
    else neighbor_traverse(T, nthreads, P->list);
    auto t2 = std::chrono::steady_clock::now();
    P->seconds_build = std::chrono::duration<double>(t1 - t0).count();
    P->seconds_traverse = std::chrono::duration<double>(t2 - t1).count();
    return P;
}

// predicate 0: d2 < fl(r*r)  (N5, the BVH predicate);  1: sqrt(d2) < r  (O1, threshold_pairs(unique_pairs))
nd_pairs* nd_brute_force(const float* xyz, int32_t n, float r, int32_t predicate, int32_t nthreads) {
    auto* P = new nd_pairs();
    brute_force(xyz, n, r, predicate == 1, nthreads, P->list);
    return P;
}

int64_t nd_pairs_count(const nd_pairs* p) { return (int64_t)p->list.size(); }
double nd_pairs_seconds_build(const nd_pairs* p) { return p->seconds_build; }
double nd_pairs_seconds_traverse(const nd_pairs* p) { return p->seconds_traverse; }
void nd_pairs_copy(const nd_pairs* p, int32_t* a, int32_t* b, float* d) {
    size_t m = p->list.size();
    if (m == 0) return;
    std::memcpy(a, p->list.a.data(), m * 4); std::memcpy(b, p->list.b.data(), m * 4); std::memcpy(d, p->list.d.data(), m * 4);
}
void nd_pairs_free(nd_pairs* p) { delete p; }

// order-independent digest of an (a,b,d) list; canonicalises (min,max)
void nd_digest_pairs(const int32_t* a, const int32_t* b, const float* d, int64_t m, int64_t* count, uint64_t* xor_hash,
                     uint64_t* sum_hash, double* sum_d) {
    uint64_t x = 0, s = 0; double sd = 0;
    for (int64_t e = 0; e < m; ++e) { uint64_t h = pair_hash(a[e], b[e], d[e]); x ^= h; s += h; sd += d[e]; }
    *count = m; *xor_hash = x; *sum_hash = s; *sum_d = sd;
}

void nd_cellgrid_digest(const float* xyz, int32_t n, float r, int32_t nthreads, int64_t* count, uint64_t* xor_hash,
                        uint64_t* sum_hash, double* sum_d, int32_t* per_atom_count) {
    Digest D;
    cellgrid_digest(xyz, n, r, nthreads, &D, per_atom_count);
    *count = D.count; *xor_hash = D.xor_hash; *sum_hash = D.sum_hash; *sum_d = D.sum_d;
}

void nd_force_lennardjones(float* force, int32_t n, const int32_t* a, const int32_t* b, const float* d, int64_t np) {
    force_lennardjones_literal(force, n, a, b, d, np);
}
void nd_force_coulomb(float* force, int32_t n, const int32_t* a, const int32_t* b, const float* d, int64_t np, const float* q) {
    force_coulomb_literal(force, n, a, b, d, np, q);
}
void nd_sum_forces(float* f, const float* f1, const float* f2, int64_t n3) { sum_forces(f, f1, f2, n3); }

void nd_forces_physical_f64(const float* xyz, const float* charge, int32_t n, const int32_t* a, const int32_t* b, int64_t np,
                            double eps, double sigma, double kc, double rc, int32_t shift, double* force, double* pe_atom,
                            double* scale) {
    forces_physical_f64(xyz, charge, n, a, b, np, eps, sigma, kc, rc, shift, force, pe_atom, scale);
}

void nd_verlet(float* pos, float* vel, const float* force, const float* force_next, const float* mass, int32_t n, float dt) {
    verlet_literal(pos, vel, force, force_next, mass, n, dt);
}
void nd_boundary_reflect(float* pos, float* vel, int32_t n, const float* mn, const float* mx) {
    boundary_reflect(pos, vel, n, mn, mx);
}

}  // extern "C"

// =========================================================================================
// MD loops: (1) the fp64 integrator oracle for the CUDA step loop; (2) the timed CPU baseline
// =========================================================================================
namespace {

// exact-N5 half pair list via the cell grid (0-based ids here)
void cell_pairs(const float* xyz, int32_t n, float r, std::vector<int32_t>& pa, std::vector<int32_t>& pb) {
    // simple reuse of brute force for small n, cell grid otherwise
    pa.clear(); pb.clear();
    const float r2 = r * r;
    if (n <= 2048) {
        for (int32_t i = 0; i < n; ++i) { V3 a = ld3(xyz, i); for (int32_t j = i + 1; j < n; ++j) if (dist2(a, ld3(xyz, j)) < r2) { pa.push_back(i + 1); pb.push_back(j + 1); } }
        return;
    }
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int32_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], xyz[3 * i + d]); hi[d] = std::max(hi[d], xyz[3 * i + d]); }
    double edge = (double)r * (1.0 + 1e-5) + 1e-30;
    int64_t dim[3];
    for (int d = 0; d < 3; ++d) dim[d] = std::max<int64_t>(1, std::min<int64_t>(512, (int64_t)std::floor(((double)hi[d] - lo[d]) / edge)));
    auto cell_of = [&](int32_t i, int d) {
        double ext = (double)hi[d] - (double)lo[d];
        if (ext <= 0) return (int64_t)0;
        int64_t c = (int64_t)std::floor(((double)xyz[3 * i + d] - lo[d]) / ext * (double)dim[d]);
        return std::max<int64_t>(0, std::min<int64_t>(dim[d] - 1, c));
    };
    int64_t ncell = dim[0] * dim[1] * dim[2];
    std::vector<int32_t> start(ncell + 1, 0), order(n);
    std::vector<int64_t> cid(n);
    for (int32_t i = 0; i < n; ++i) { cid[i] = (cell_of(i, 2) * dim[1] + cell_of(i, 1)) * dim[0] + cell_of(i, 0); start[cid[i] + 1]++; }
    for (int64_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    { std::vector<int32_t> cur(start.begin(), start.end() - 1); for (int32_t i = 0; i < n; ++i) order[cur[cid[i]]++] = i; }
    for (int64_t c = 0; c < ncell; ++c) {
        int64_t cx = c % dim[0], cy = (c / dim[0]) % dim[1], cz = c / (dim[0] * dim[1]);
        for (int32_t ii = start[c]; ii < start[c + 1]; ++ii) {
            int32_t i = order[ii]; V3 a = ld3(xyz, i);
            for (int64_t dz = -1; dz <= 1; ++dz) for (int64_t dy = -1; dy <= 1; ++dy) for (int64_t dx = -1; dx <= 1; ++dx) {
                int64_t nx = cx + dx, ny = cy + dy, nz = cz + dz;
                if (nx < 0 || ny < 0 || nz < 0 || nx >= dim[0] || ny >= dim[1] || nz >= dim[2]) continue;
                int64_t nc = (nz * dim[1] + ny) * dim[0] + nx;
                for (int32_t jj = start[nc]; jj < start[nc + 1]; ++jj) {
                    int32_t j = order[jj];
                    if (j <= i) continue;
                    if (dist2(a, ld3(xyz, j)) < r2) { pa.push_back(i + 1); pb.push_back(j + 1); }
                }
            }
        }
    }
}

}  // namespace

extern "C" {

// fp64 KDK velocity-Verlet oracle.  State in/out as double (positions are rounded to float
// only for the neighbour predicate, exactly like the CUDA path which stores float positions;
// the pair set used is the exact-N5 set of the float-rounded positions).
// energies[0]=KE, [1]=PE after the last step.  `force` in/out (3n doubles) = F(x) at entry/exit.
void nd_md_steps_f64(double* pos, double* vel, double* force, const float* mass, const float* charge, int32_t n,
                     int32_t nsteps, double dt, float cutoff, double eps, double sigma, double kc, int32_t shift,
                     const float* bmin, const float* bmax, int32_t init_force, double* energies) {
    std::vector<float> xf(3 * (size_t)n);
    std::vector<int32_t> pa, pb;
    std::vector<double> pe(n);
    auto eval = [&]() {
        for (int64_t i = 0; i < 3 * (int64_t)n; ++i) xf[i] = (float)pos[i];
        cell_pairs(xf.data(), n, cutoff, pa, pb);
        // forces from the double positions over that pair set
        for (int64_t i = 0; i < 3 * (int64_t)n; ++i) force[i] = 0.0;
        for (int32_t i = 0; i < n; ++i) pe[i] = 0.0;
        double rc = (double)cutoff;
        double src6 = std::pow(sigma / rc, 6.0);
        double ulj_rc = shift ? 4.0 * eps * (src6 * src6 - src6) : 0.0;
        for (size_t e = 0; e < pa.size(); ++e) {
            int32_t i = pa[e] - 1, j = pb[e] - 1;
            double dx = pos[3 * i] - pos[3 * j], dy = pos[3 * i + 1] - pos[3 * j + 1], dz = pos[3 * i + 2] - pos[3 * j + 2];
            double r2 = dx * dx + dy * dy + dz * dz, r = std::sqrt(r2);
            double s2 = sigma * sigma / r2, s6 = s2 * s2 * s2;
            double fs = 24.0 * eps * (2.0 * s6 * s6 - s6) / r2;
            double u = 4.0 * eps * (s6 * s6 - s6) - ulj_rc;
            if (charge) { double qq = kc * (double)charge[i] * (double)charge[j]; fs += qq / (r2 * r); u += qq * (1.0 / r - (shift ? 1.0 / rc : 0.0)); }
            force[3 * i] += fs * dx; force[3 * i + 1] += fs * dy; force[3 * i + 2] += fs * dz;
            force[3 * j] -= fs * dx; force[3 * j + 1] -= fs * dy; force[3 * j + 2] -= fs * dz;
            pe[i] += 0.5 * u; pe[j] += 0.5 * u;
        }
    };
    if (init_force) eval();
    for (int32_t s = 0; s < nsteps; ++s) {
        for (int32_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) {
            double& v = vel[3 * i + d]; double& x = pos[3 * i + d];
            v += force[3 * i + d] / (double)mass[i] * (0.5 * dt);
            x += v * dt;
            if (x < (double)bmin[d]) { v = -v; x = (double)bmin[d]; }   // boundary_reflect! semantics (Simulator.jl:81-111)
            if (x > (double)bmax[d]) { v = -v; x = (double)bmax[d]; }
        }
        eval();
        for (int32_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) vel[3 * i + d] += force[3 * i + d] / (double)mass[i] * (0.5 * dt);
    }
    if (nsteps == 0 && !init_force) eval();
    double ke = 0, pes = 0;
    for (int32_t i = 0; i < n; ++i) {
        ke += 0.5 * (double)mass[i] * (vel[3 * i] * vel[3 * i] + vel[3 * i + 1] * vel[3 * i + 1] + vel[3 * i + 2] * vel[3 * i + 2]);
        pes += pe[i];
    }
    energies[0] = ke; energies[1] = pes;
}

// The timed CPU BASELINE: one "reference-shaped" MD step, Float32, exactly the composition
// BASELINE.md §2 describes: leafTreeData (serial keys + serial stable sort + threaded build) ->
// leafneighbor_traverse (threaded, per-thread vectors, serial concat) -> serial pair-force loop
// (Forces.jl-style single pass over the list; physical LJ(+Coulomb) so that it is the same
// physics the GPU runs) -> Verlet body (Simulator.jl:198-222) -> boundary_reflect!.
// qstride > 1: traverse every qstride-th query leaf only (bounded sample); the caller scales.
// timings[0..4] = tree, traverse(+concat), force, verlet, total  (seconds)
int64_t nd_cpu_step(float* pos, float* vel, float* force, const float* mass, const float* charge, int32_t n, float dt,
                    float cutoff, int32_t apl, int32_t nthreads, int32_t qstride, float eps, float sigma, float kc,
                    const float* bmin, const float* bmax, double* timings) {
    auto t0 = std::chrono::steady_clock::now();
    Tree T;
    int rc = build_tree(pos, n, cutoff, apl, true, nthreads, &T);
    if (rc) return -rc;
    auto t1 = std::chrono::steady_clock::now();
    PairList L;
    leafneighbor_traverse(T, nthreads, qstride, L);
    auto t2 = std::chrono::steady_clock::now();
    std::vector<float> fnext(3 * (size_t)n, 0.0f);
    const int64_t np = (int64_t)L.size();
    for (int64_t e = 0; e < np; ++e) {  // serial, like force_lennardjones!/force_coulomb! (Forces.jl:39-42,62-65)
        int32_t i = L.a[e] - 1, j = L.b[e] - 1;
        float dx = pos[3 * i] - pos[3 * j], dy = pos[3 * i + 1] - pos[3 * j + 1], dz = pos[3 * i + 2] - pos[3 * j + 2];
        float d = L.d[e];
        float r2 = d * d;
        float s2 = sigma * sigma / r2, s6 = s2 * s2 * s2;
        float fs = 24.0f * eps * (2.0f * s6 * s6 - s6) / r2;
        if (charge) fs += kc * charge[i] * charge[j] / (r2 * d);
        fnext[3 * i] += fs * dx; fnext[3 * i + 1] += fs * dy; fnext[3 * i + 2] += fs * dz;
        fnext[3 * j] -= fs * dx; fnext[3 * j + 1] -= fs * dy; fnext[3 * j + 2] -= fs * dz;
    }
    auto t3 = std::chrono::steady_clock::now();
    verlet_literal(pos, vel, force, fnext.data(), mass, n, dt);
    boundary_reflect(pos, vel, n, bmin, bmax);
    std::memcpy(force, fnext.data(), sizeof(float) * 3 * (size_t)n);
    auto t4 = std::chrono::steady_clock::now();
    auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
    if (timings) { timings[0] = sec(t0, t1); timings[1] = sec(t1, t2); timings[2] = sec(t2, t3); timings[3] = sec(t3, t4); timings[4] = sec(t0, t4); }
    return np;
}

}  // extern "C"
