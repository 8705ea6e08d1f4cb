// api.cu — the C ABI (include/naiveb200.h): handle lifecycle, host<->device staging, the stage
// pipeline, regrow protocol, profiling.  No torch types, no exceptions across the boundary.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <vector>

#include "nb200_internal.cuh"

using namespace nb200;

namespace {

thread_local char g_create_error[512] = "";

int32_t fail(nb200_handle* h, int32_t code, const char* fmt, ...) {
    char* dst = h ? h->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(h, expr)                                                                                  \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) return fail(h, NB200_ERR_CUDA, "%s -> %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

#define CHECK_LAUNCH(h, what)                                                                        \
    do {                                                                                             \
        cudaError_t e_ = cudaGetLastError();                                                         \
        if (e_ != cudaSuccess) return fail(h, NB200_ERR_CUDA, "launch %s -> %s", what, cudaGetErrorString(e_)); \
    } while (0)

// tile groups: every drain pass of a query leaf writes one group; all but a leaf's last pass hold >= 4 tiles — traverse.cu
int64_t seg_capacity_for(int64_t n_max, int64_t word_capacity) {
    return (n_max + LEAF - 1) / LEAF + word_capacity / (4 * TILE_WORDS) + 64;
}

template <class T>
cudaError_t dalloc(T** p, int64_t count) {
    *p = nullptr;
    if (count <= 0) count = 1;
    return cudaMalloc((void**)p, sizeof(T) * (size_t)count);
}

// ---- stage timing ------------------------------------------------------------------------------------
void timer_collect(nb200_handle* h) {
    StageTimer& t = h->timer;
    if (!t.created || t.n_ev == 0) return;
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i + 1 < t.n_ev; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.ev[i], t.ev[i + 1]) == cudaSuccess) t.ms[t.stage_of[i]] += ms;
    }
    t.n_ev = 0;
}

struct StageScope {
    nb200_handle* h;
    int stage;
    bool on;
    StageScope(nb200_handle* hh, int st) : h(hh), stage(st), on(hh->timer.enabled && (hh->timer.only_stage < 0 || hh->timer.only_stage == st)) {
        if (!on) return;
        StageTimer& t = h->timer;
        if (t.n_ev + 2 > StageTimer::MAX_EVENTS) timer_collect(h);
        t.stage_of[t.n_ev] = stage;
        cudaEventRecord(t.ev[t.n_ev], h->stream);
    }
    void add(int launches) {
        h->kernel_launches += launches;
        if (on) h->timer.launches[stage] += launches;
    }
    void count_as(int) {}  // (a fused launch is timed under the stage that launched it)
    ~StageScope() {
        if (!on) return;
        StageTimer& t = h->timer;
        cudaEventRecord(t.ev[t.n_ev + 1], h->stream);
        t.n_ev += 2;
    }
};

int32_t ensure_scratch(nb200_handle* h, int64_t bytes) {
    if (bytes <= h->scratch_bytes) return NB200_OK;
    if (h->scratch_dev) cudaFree(h->scratch_dev);
    h->scratch_dev = nullptr;
    h->scratch_bytes = 0;
    CU(h, cudaMalloc(&h->scratch_dev, (size_t)bytes));
    h->scratch_bytes = bytes;
    return NB200_OK;
}

int32_t ensure_entries(nb200_handle* h, int64_t entries_needed) {
    if (entries_needed <= h->entry_capacity && h->entries) return NB200_OK;
    int64_t cap = entries_needed + entries_needed / 4 + 4096;
    int64_t seg_cap = seg_capacity_for(h->n_max, cap);
    if (h->entries) cudaFree(h->entries);
    if (h->segs) cudaFree(h->segs);
    h->entries = nullptr;
    h->segs = nullptr;
    h->entry_capacity = 0;
    CU(h, dalloc(&h->entries, cap));
    CU(h, dalloc(&h->segs, seg_cap));
    h->entry_capacity = cap;
    h->seg_capacity = seg_cap;
    h->regrows++;
    return NB200_OK;
}

// unique pairs in the current list: a half list holds each pair once, a directed list twice
int64_t list_pairs(const nb200_handle* h) {
    return (int64_t)(h->list_half ? h->counters_h->n_valid : h->counters_h->n_valid / 2);
}

int32_t read_counters(nb200_handle* h) {
    CU(h, cudaMemcpyAsync(h->counters_h, h->counters, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->counters_h->stack_overflow) {  // sticky device flag of traverse_kernel: the list of that search is incomplete
        CU(h, cudaMemsetAsync(&h->counters->stack_overflow, 0, sizeof(unsigned int), h->stream));
        h->counters_h->stack_overflow = 0;
        h->list_valid = false;
        h->have_forces = false;
        return fail(h, NB200_ERR_STATE, "traversal stack overflow: the tree over these keys is deeper than the %d-entry warp stack "
                    "allows (degenerate key distribution); the neighbour list is incomplete", 192);
    }
    return NB200_OK;
}

// the current list as (a, b, d) arrays of original atom ids in exp_a / exp_b / exp_d (device); np = list_pairs(h)
int32_t export_device(nb200_handle* h, int64_t np, int32_t index_base) {
    if (h->exp_capacity < np) {
        cudaFree(h->exp_a); cudaFree(h->exp_b); cudaFree(h->exp_d);
        h->exp_a = h->exp_b = nullptr; h->exp_d = nullptr; h->exp_capacity = 0;
        CU(h, dalloc(&h->exp_a, np));
        CU(h, dalloc(&h->exp_b, np));
        CU(h, dalloc(&h->exp_d, np));
        h->exp_capacity = np;
    }
    {
        StageScope sc(h, NB200_STAGE_EXPORT);
        sc.add(launch_export(h->stream, h->sm_count, h->segs, h->entries, h->counters, h->seg_capacity, h->pos[h->cur],
                             h->id[h->cur], h->n, h->exp_a, h->exp_b, h->exp_d, np, index_base));
        CHECK_LAUNCH(h, "export");
    }
    return NB200_OK;
}

// keys[0]/vals[0] hold the Morton keys of pos[cur]: sort, gather into pos[cur^1], build, traverse.
// `with_vel`: carry velocities (MD state) or not (search-only entry point).
// `resort` = false (step loop with a re-sort interval > 1): the atoms keep the order of the last sort — after a few
// steps it is still a space-filling-curve order to within the atoms' displacement — and only the leaf boxes are
// recomputed from the current positions before the tree is rebuilt and traversed (the reference's TreeData!
// update path, BVHTraverse.jl:601-655); the neighbour list is rebuilt from scratch either way.
// `fused`: the traversal also evaluates the pair forces of every tile it emits (force[] was zeroed by the reorder
// kernel); the caller then skips enqueue_force.
int32_t enqueue_search(nb200_handle* h, bool with_vel, float cutoff, bool resort = true, bool fused = false, bool resident_order = false) {
    const int n = h->n;
    // the reorder kernel also initialises the scratch of the kernels after it (Housekeeping): build flags and traversal
    // counters for this step, the sort's scratch for the next one
    Housekeeping hk = {};
    hk.node_flag = h->node_flag;
    hk.n_flag = h->n_leaves > 1 ? h->n_leaves - 1 : 0;
    hk.counters = reinterpret_cast<uint32_t*>(h->counters);
    hk.n_counter_words = (int)(COUNTERS_RESET_BYTES / 4);
    if (!resort) {
        StageScope sc(h, NB200_STAGE_REORDER);
        sc.add(launch_reorder(h->stream, nullptr, h->keys[0], h->pos[h->cur], nullptr, nullptr, nullptr, nullptr, nullptr, h->force,
                              h->leaf_lo, h->leaf_hi, h->leaf_sub, n, cutoff, &hk));
        CHECK_LAUNCH(h, "leaf refresh");
        h->steps_since_sort++;
    } else {
    int buf = 0;
    // Key bits that matter.  A search on atoms in ARBITRARY order (first search of a system, one-call searches) sorts
    // log2(n) + 4 bits: cells 16x finer than one atom each (1M atoms -> bits [6,30), 3 passes).  The step loop re-sorts
    // atoms that are already in curve order from the step before (`resident_order`): ceil(log2(n)) - 4 bits are enough —
    // cells of 8-16 atoms, half a leaf; the stable sort keeps the previous order inside a cell and a leaf spans two
    // cells either way (tools/sort_bits_model.py: 17.1 -> 17.8 candidate leaves per query leaf at 1M atoms and 16 bits;
    // measured: traversal +2 %, one 37-us sort pass less).  Cells of ~30 atoms are too coarse: the order inside a cell
    // decays over a few hundred steps and the traversal slows down by 25 %.  Whole 8-bit passes from the top of the key.
    int lg = 0;
    while ((1ll << lg) < n && lg < 30) ++lg;
    int bits = resident_order ? lg - 4 : lg + 4;
    if (bits > 30) bits = 30;
    int passes = (bits + 7) / 8;
    if (h->sort_passes_override > 0) passes = h->sort_passes_override;  // tuning aid (NB200_SORT_PASSES)
    if (passes < 2) passes = 2;
    if (passes > 4) passes = 4;
    const int low_bit = passes == 4 ? 0 : 30 - 8 * passes;
    {
        StageScope sc(h, NB200_STAGE_SORT);
        const bool clean = h->hk_sort_clean && h->hk_n == n && h->hk_passes == passes;
        h->hk_sort_clean = false;
        sc.add(launch_sort(h->stream, h->keys, h->vals, n, h->sort_hist, h->sort_status, h->sort_ticket, &buf, low_bit, passes, clean));
        CHECK_LAUNCH(h, "sort");
    }
    const int src = h->cur, dst = h->cur ^ 1;
    {
        StageScope sc(h, NB200_STAGE_REORDER);
        const Housekeeping sk = sort_housekeeping(n, passes, h->sort_hist, h->sort_status, h->sort_ticket);
        hk.sort_hist = sk.sort_hist; hk.n_hist = sk.n_hist;
        hk.sort_ticket = sk.sort_ticket; hk.n_ticket = sk.n_ticket;
        hk.sort_status = sk.sort_status; hk.n_status = sk.n_status;
        sc.add(launch_reorder(h->stream, h->vals[buf], h->keys[buf], h->pos[src], with_vel ? h->vel[src] : nullptr, h->id[src],
                              h->pos[dst], h->vel[dst], h->id[dst], h->force, h->leaf_lo, h->leaf_hi, h->leaf_sub, n, cutoff, &hk));
        CHECK_LAUNCH(h, "reorder");
        h->hk_sort_clean = true;
        h->hk_n = n;
        h->hk_passes = passes;
    }
    h->cur = dst;
    h->steps_since_sort = 0;
    }
    {
        StageScope sc(h, NB200_STAGE_BUILD);
        sc.add(launch_build(h->stream, h->leaf_lo, h->leaf_hi, h->n_leaves, h->nodes, h->node_lo, h->node_hi, h->node_flag, true));
        sc.add(launch_frontier(h->stream, h->nodes, h->n_leaves, h->frontier, 0, h->node_lo, h->node_hi, h->leaf_lo, h->leaf_hi));
        CHECK_LAUNCH(h, "build");
    }
    {
        StageScope sc(h, NB200_STAGE_TRAVERSE);
        h->list_half = h->list_mode == NB200_LIST_HALF;
        sc.add(launch_traverse(h->stream, h->sm_count, h->nodes, h->frontier, h->leaf_lo, h->leaf_hi, h->leaf_sub, h->pos[h->cur], n, h->n_leaves, cutoff,
                               h->entries, h->entry_capacity, h->segs, h->seg_capacity, h->counters, h->list_half, nullptr,
                               nullptr, true, fused ? &h->ff : nullptr, fused ? h->force : nullptr));
        CHECK_LAUNCH(h, "traverse");
        if (fused) {
            sc.count_as(NB200_STAGE_FORCE);
            h->pe_valid = false;
        }
    }
    h->cutoff = cutoff;
    return NB200_OK;
}

// synchronous search with the regrow-and-retry protocol (tree stays valid, only the traversal reruns)
// `headroom`: the list is about to be rebuilt every step by an asynchronous loop that cannot regrow, so
// make sure the buffer holds the current list plus 20 % before returning.
int32_t search_sync(nb200_handle* h, bool with_vel, float cutoff, bool headroom = false) {
    // an overflow of THIS call is handled by the regrow-and-retry below and must not reach a later nb200_sync as the
    // overflow of a step loop — but a sticky flag that unreported asynchronous steps left must survive this call
    CU(h, cudaMemcpyAsync(&h->counters->sticky_saved, &h->counters->overflow_sticky, sizeof(unsigned int), cudaMemcpyDeviceToDevice, h->stream));
    int32_t rc = enqueue_search(h, with_vel, cutoff);
    if (rc) return rc;
    for (int attempt = 0; attempt < 4; ++attempt) {
        rc = read_counters(h);
        if (rc) return rc;
        const int64_t need = (int64_t)h->counters_h->n_entries();
        const bool tight = headroom && (need + need / 5 + 4096 > h->entry_capacity);
        if (!h->counters_h->overflow && !tight) {
            if (h->counters_h->overflow_sticky)
                CU(h, cudaMemcpyAsync(&h->counters->overflow_sticky, &h->counters->sticky_saved, sizeof(unsigned int), cudaMemcpyDeviceToDevice, h->stream));
            h->list_valid = true;
            return NB200_OK;
        }
        rc = ensure_entries(h, tight ? need + need / 4 : need);
        if (rc) return rc;
        StageScope sc(h, NB200_STAGE_TRAVERSE);
        sc.add(launch_traverse(h->stream, h->sm_count, h->nodes, h->frontier, h->leaf_lo, h->leaf_hi, h->leaf_sub, h->pos[h->cur], h->n, h->n_leaves,
                               cutoff, h->entries, h->entry_capacity, h->segs, h->seg_capacity, h->counters, h->list_half, nullptr,
                               nullptr));
        CHECK_LAUNCH(h, "traverse(retry)");
    }
    return fail(h, NB200_ERR_PAIR_OVERFLOW, "neighbour buffer still too small after regrowing");
}

int32_t enqueue_force(nb200_handle* h, bool with_pe) {
    // eps == 0 and kcoul == 0: the force-free loop of simulate_bvh! (Simulator.jl:327-379) — force[] stays
    // at the zeros the reorder kernel wrote
    if (h->ff.eps == 0.f && h->ff.kcoul == 0.f) {
        h->pe_valid = true;
        return NB200_OK;
    }
    StageScope sc(h, NB200_STAGE_FORCE);
    // a list built with a larger cutoff than the force field's (skin list of nb200_set_list_reuse): the force kernel
    // re-applies the exact pair predicate at the force cutoff
    sc.add(launch_force(h->stream, h->sm_count, h->segs, h->entries, h->counters, h->seg_capacity, h->pos[h->cur], h->force,
                        h->n, h->ff, with_pe, h->list_half, h->cutoff > h->ff.cutoff, h->mg_active ? h->mg_gbase : 0x7fffffff));
    CHECK_LAUNCH(h, "force");
    h->pe_valid = with_pe;
    return NB200_OK;
}

// The step loop evaluates the pair forces inside the traversal unless told otherwise (nb200_set_fused_force), a skin
// list is in use (its pairs need the cutoff check of the separate kernel) or there is no force field at all.
bool step_fuses_forces(const nb200_handle* h) {
    return h->fused_force && h->reuse_every <= 1 && !(h->ff.eps == 0.f && h->ff.kcoul == 0.f);
}

// cutoff the step loop's list is built with: the force cutoff, plus the skin when the list is reused across steps
float md_list_cutoff(const nb200_handle* h) { return h->reuse_every > 1 ? h->ff.cutoff + h->reuse_skin : h->ff.cutoff; }

// list reuse: remember where the atoms were when the list was built (vel[cur^1] is dead between two re-sorts)
int32_t mark_list_built(nb200_handle* h) {
    h->list_age = 0;
    if (h->reuse_every > 1) {
        CU(h, cudaMemcpyAsync(h->vel[h->cur ^ 1], h->pos[h->cur], sizeof(float4) * (size_t)h->n, cudaMemcpyDeviceToDevice, h->stream));
    }
    return NB200_OK;
}

// A new system replaces the state earlier asynchronous steps ran on: whatever they would still have reported at the
// next nb200_sync (neighbour-buffer overflow, an atom that outran the skin of a reused list) no longer concerns the caller.
void reset_async_reports(nb200_handle* h) {
    h->async_overflow_possible = false;
    cudaMemsetAsync(h->reuse_d2, 0, 2 * sizeof(unsigned int), h->stream);
}

void graph_invalidate(nb200_handle* h) {
    if (h->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->graph_exec);
    h->graph_exec = nullptr;
}

int32_t check_n(nb200_handle* h, int64_t n) {
    if (n < 1) return fail(h, NB200_ERR_BAD_ARG, "atom count must be >= 1 (got %lld)", (long long)n);
    if (n > h->n_max) return fail(h, NB200_ERR_BAD_ARG, "atom count %lld exceeds the handle's n_max %lld", (long long)n, (long long)h->n_max);
    return NB200_OK;
}

int32_t upload_system(nb200_handle* h, const float* xyz, const float* vel, int32_t stride, const float* mass,
                      const float* charge, int32_t n, bool with_vel) {
    float* sx = h->stage_dev;
    float* sv = sx + (size_t)n * 4;
    float* sm = sv + (size_t)n * 4;
    float* sq = sm + (size_t)n;
    CU(h, cudaMemcpyAsync(sx, xyz, sizeof(float) * (size_t)n * stride, cudaMemcpyHostToDevice, h->stream));
    if (vel) CU(h, cudaMemcpyAsync(sv, vel, sizeof(float) * (size_t)n * stride, cudaMemcpyHostToDevice, h->stream));
    if (mass) CU(h, cudaMemcpyAsync(sm, mass, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    if (charge) CU(h, cudaMemcpyAsync(sq, charge, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    h->n = n;
    h->n_leaves = (n + LEAF - 1) / LEAF;
    h->cur = 0;
    h->kernel_launches += launch_pack(h->stream, sx, stride, vel ? sv : nullptr, mass ? sm : nullptr, charge ? sq : nullptr, n,
                                      h->pos[0], with_vel ? h->vel[0] : nullptr, h->id[0]);
    CHECK_LAUNCH(h, "pack");
    return NB200_OK;
}

int32_t compute_forces_sync(nb200_handle* h);

int32_t download_vec(nb200_handle* h, float* out, int32_t stride, int mode) {
    if (!h->have_system && mode != 0) return fail(h, NB200_ERR_STATE, "no system loaded (call nb200_set_system first)");
    if (h->n <= 0) return fail(h, NB200_ERR_STATE, "no atoms loaded");
    if (stride != 3 && stride != 4) return fail(h, NB200_ERR_BAD_ARG, "stride must be 3 or 4");
    if (mode == 1 && h->vel_half && !h->have_forces) {  // closing half kick needs F at the current positions
        int32_t rc = compute_forces_sync(h);
        if (rc) return rc;
    }
    const float4* src = mode == 0 ? h->pos[h->cur] : (mode == 1 ? h->vel[h->cur] : h->force);
    const bool pending = (mode == 1) && h->vel_half;
    h->kernel_launches += launch_unpack(h->stream, src, h->id[h->cur], h->n, stride, h->stage_dev, mode,
                                        pending ? h->force : nullptr, 0.5f * h->last_dt);
    CHECK_LAUNCH(h, "unpack");
    CU(h, cudaMemcpyAsync(out, h->stage_dev, sizeof(float) * (size_t)h->n * stride, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

// Layout of a rank's published region.  It follows the rank's CAPACITY (its handle's n_max), not its current atom count,
// so that atoms can migrate between ranks without remapping:
//   [header 256 B | pos x2 | leaf boxes x2 | global ids x2 | outbox: pos, vel, gid, dest]
struct PubLayout {
    unsigned int* hdr;
    float4* pos[2];
    float4* box[2];
    int32_t* id[2];
    float4* out_pos;
    float4* out_vel;
    int32_t* out_gid;
    int32_t* out_dest;
    int64_t out_cap;
    int64_t bytes;
};
PubLayout pub_layout(void* base, int64_t cap) {
    PubLayout L;
    const int64_t nPL = (cap + LEAF - 1) / LEAF;
    L.out_cap = cap / 8 + 1024;
    char* b = (char*)base;
    L.hdr = (unsigned int*)b;
    L.pos[0] = (float4*)(b + 256);
    L.pos[1] = L.pos[0] + cap;
    L.box[0] = L.pos[1] + cap;
    L.box[1] = L.box[0] + 2 * nPL;
    L.out_pos = L.box[1] + 2 * nPL;
    L.out_vel = L.out_pos + L.out_cap;
    L.id[0] = (int32_t*)(L.out_vel + L.out_cap);
    L.id[1] = L.id[0] + cap;
    L.out_gid = L.id[1] + cap;
    L.out_dest = L.out_gid + L.out_cap;
    L.bytes = (char*)(L.out_dest + L.out_cap) - b;
    return L;
}

void mg_close_peers(nb200_handle* h) {
    for (int p = 0; p < 64; ++p)
        if (h->mg_ipc_opened[p]) {
            cudaIpcCloseMemHandle(h->mg_ipc_opened[p]);
            h->mg_ipc_opened[p] = nullptr;
        }
    h->mg_connected = false;
}

// ---- multi-GPU search over the two-segment arrays ---------------------------------------------
// Owned atoms: resident in curve order in pos/vel/id[cur], slots [0, n_own), their keys in keys[0]/vals[0].
// Ghosts of the step: pre-sort arrays mg_gpos / mg_gkeys[0] / mg_gvals[0], slots [0, n_g) (real ghosts first, inert NaN
// placeholders behind them).  Two sorts, two gathers, two trees; the ghost side runs on a second stream when
// `two_streams`, so waiting for the peers' publications and building the small ghost tree hide under the owned sort +
// tree build, and the traversal (owned leaves query both trees) starts when both are ready.
// pass 1: the owned segment against its own tree — the single-GPU kernel; pass 2 (ghost pass): the same owned leaves
// against the ghost tree, appended to the same tile list
int32_t mg_launch_traverse(nb200_handle* h, bool fused, bool counters_clean, int pass, cudaStream_t st) {
    StageScope sc(h, NB200_STAGE_TRAVERSE);
    h->list_half = h->list_mode == NB200_LIST_HALF;
    if (pass == 1) {
        sc.add(launch_traverse(st, h->sm_count, h->nodes, h->frontier, h->leaf_lo, h->leaf_hi, h->leaf_sub, h->pos[h->cur], h->mg_gbase,
                               h->mg_nLo, h->ff.cutoff, h->entries, h->entry_capacity, h->segs, h->seg_capacity, h->counters, h->list_half,
                               nullptr, nullptr, counters_clean, fused ? &h->ff : nullptr, fused ? h->force : nullptr));
    } else {
        MgSearch ms;
        ms.n_query = h->mg_n_own;
        ms.blist = h->mg_blist;
        ms.bcount = h->mg_bcount;
        sc.add(launch_traverse(st, h->sm_count, h->nodes, h->frontier2, h->leaf_lo, h->leaf_hi, h->leaf_sub, h->pos[h->cur], h->n,
                               h->mg_nLo, h->ff.cutoff, h->entries, h->entry_capacity, h->segs, h->seg_capacity, h->counters, h->list_half,
                               nullptr, &ms, true, fused ? &h->ff : nullptr, fused ? h->force : nullptr));
    }
    CHECK_LAUNCH(h, "traverse(slab)");
    if (fused) h->pe_valid = false;
    return NB200_OK;
}

void mg_trace(nb200_handle* h, int k, cudaStream_t st) {
    if (!h->mg_trace) return;
    if (!h->mg_trace_ev[0])
        for (int i = 0; i < 32 * 6; ++i) cudaEventCreate(&h->mg_trace_ev[i]);
    if (k == 0) ++h->mg_trace_step;
    cudaEventRecord(h->mg_trace_ev[(h->mg_trace_step % 32) * 6 + k], st);
}

// Ghost slots the handle can provide: what is left behind the owned segment, minus a reserve for the owned segment to grow
// into when atoms migrate here (an eighth of the owned atoms + 1024).
int64_t mg_ghost_alloc(const nb200_handle* h) {
    const int64_t reserve = h->mg_migrate_every > 0 ? h->mg_n_own / 8 + 1024 : 0;
    const int64_t g = h->n_max - h->mg_gbase - reserve;
    return g > 2 * LEAF ? g : 2 * LEAF;
}

// Is this slab ragged?  Compare the volume of its AABB with the volume its atoms would fill at the mean density of the
// whole system: a compact slab (uniform data, 2^k ranks) gets ~1 and the asynchronous step skips the occupancy grid (three
// stream operations per step); clustered data or odd rank counts get >> 1 and keep it.
bool mg_slab_is_ragged(const nb200_handle* h, const int* box6, int64_t n_own) {
    auto ord2f = [](int i) { i ^= ((i >> 31) & 0x7fffffff); float f; std::memcpy(&f, &i, 4); return f; };
    double vol = 1.0, boxvol = 1.0;
    for (int d = 0; d < 3; ++d) {
        vol *= std::max(0.0, (double)ord2f(box6[3 + d]) - (double)ord2f(box6[d]));
        boxvol *= (double)h->box_max[d] - (double)h->box_min[d];
    }
    const double need = boxvol * (double)n_own / (double)std::max<int64_t>(h->mg_n_total, n_own);
    return !(vol <= 1.5 * need);
}

// Second half of a migration step (the first is in nb200_mg_integrate): the atoms the peers sent here are appended behind
// the owned atoms; the counts come back to the host (the one host round trip per migration), which re-sizes the owned
// segment: n_own <- n_own - leavers + immigrants.
int32_t mg_take_immigrants(nb200_handle* h) {
    if (!h->mg_migration_pending) return NB200_OK;
    h->mg_migration_pending = false;
    const int n_old = h->mg_n_own;
    const int64_t room64 = h->n_max - n_old - h->mg_ghost_cap - 4 * LEAF;
    const int room = room64 > 0 ? (int)room64 : 0;
    h->kernel_launches += launch_mg_immigrate(h->stream, h->mg_peers_dev, h->mg_world, h->mg_rank, h->mg_max_peer_out, h->pos[h->cur], h->vel[h->cur],
                                              h->id[h->cur], h->keys[0], h->vals[0], n_old, room, h->mg_ghost_count + 1, h->box_min, h->box_max,
                                              h->curve, h->mg_err, 10000000000ll);
    CHECK_LAUNCH(h, "immigrate");
    unsigned int cnt[3] = {0, 0, 0};
    CU(h, cudaMemcpyAsync(&cnt[0], h->mg_flag + HDR_OUT, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(&cnt[1], h->mg_ghost_count + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(&cnt[2], h->mg_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (cnt[2]) {
        cudaMemsetAsync(h->mg_err, 0, sizeof(unsigned int), h->stream);
        return fail(h, NB200_ERR_STATE, "peer %u did not publish its migration outbox within the time limit", cnt[2] - 1u);
    }
    const int n_out = (int)cnt[0], n_in = (int)cnt[1];
    if (n_out > h->mg_out_cap) return fail(h, NB200_ERR_CAPACITY, "%d atoms left this rank in one migration, the outbox holds %lld", n_out, (long long)h->mg_out_cap);
    if (n_in > room) return fail(h, NB200_ERR_CAPACITY, "%d atoms migrated to this rank, the handle has room for %d more (n_max %lld)", n_in, room, (long long)h->n_max);
    const int n_new = n_old - n_out + n_in;
    if (n_new < 2) return fail(h, NB200_ERR_STATE, "this rank would own %d atoms after the migration", n_new);
    h->mg_n_pre = n_old + n_in;
    h->mg_n_own = n_new;
    h->mg_nLo = (n_new + LEAF - 1) / LEAF;
    h->mg_gbase = h->mg_nLo * LEAF;
    if (h->mg_gbase + h->mg_ghost_cap > h->n_max) h->mg_ghost_cap = h->n_max - h->mg_gbase;  // (the owned segment grew into the ghost slots)
    h->mg_sort_extra = true;
    h->mg_pub_current = false;  // the publication of this step predates the hand-over: a later synchronous search republishes
    h->mg_migrated = true;
    h->mg_last_out = n_out;
    h->mg_last_in = n_in;
    h->mg_keys_ready = true;  // keys[0] / vals[0] hold the step's keys: leavers marked, immigrants appended
    // the slab box the integrate accumulated does not know the immigrants: extend it (the leavers stay inside, harmless)
    h->kernel_launches += launch_slab_box(h->stream, h->pos[h->cur] + n_old, n_in, h->mg_box + 8 * h->mg_parity, false);
    CHECK_LAUNCH(h, "slab box(immigrants)");
    if (!h->mg_use_grid) {
        // Atoms that migrate into a far corner of this rank's key range stretch the slab box across the domain; from then
        // on the box alone would select a large part of the peers' atoms as ghosts, so the occupancy grid is switched on
        // (one more short host wait, on a step that has just waited for the counts anyway).
        int box6[6];
        CU(h, cudaMemcpyAsync(box6, h->mg_box + 8 * h->mg_parity, sizeof(box6), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        h->mg_use_grid = mg_slab_is_ragged(h, box6, n_new);
    }
    CU(h, cudaEventRecord(h->mg_ev_int, h->stream));  // what the ghost stream waits for
    graph_invalidate(h);
    return NB200_OK;
}

int32_t mg_search(nb200_handle* h, int n_g, bool two_streams, bool fused, bool coarse, bool owned_keys_ready,
                  const std::function<int32_t()>* before_ghost_side = nullptr) {
    const int n_own = h->mg_n_own, nLo = h->mg_nLo, gbase = h->mg_gbase;
    // migration step: the pre-sort arrays hold last step's owned atoms (the leavers with key 0xffffffff) + the immigrants;
    // one more pass over the top key bits puts the leavers strictly behind every real atom, and the gather stops before them
    const int n_pre = h->mg_n_pre > 0 ? h->mg_n_pre : n_own;
    const bool migrated_now = h->mg_sort_extra;
    const int nLg = (n_g + LEAF - 1) / LEAF;
    const float cutoff = h->ff.cutoff;
    const int src = h->cur, dst = h->cur ^ 1;
    cudaStream_t sA = h->stream, sB = two_streams ? h->mg_stream2 : h->stream;
    // ---- owned side ----
    if (!owned_keys_ready) {
        StageScope sc(h, NB200_STAGE_MORTON);
        sc.add(launch_morton(sA, h->pos[src], n_pre, h->box_min, h->box_max, h->keys[0], h->vals[0], h->curve));
        CHECK_LAUNCH(h, "morton(owned)");
    }
    // Same rule as enqueue_search, but the key space is the WHOLE box while this rank's atoms fill 1 / world of it: the
    // cell size follows the global atom count (cells of 16-32 atoms: tools/mg_decay_bench.py shows the tile count of the
    // list flat over 400 steps at that size).
    int lg = 0;
    const int64_t n_box = h->mg_n_total > n_own ? h->mg_n_total : n_own;
    while ((1ll << lg) < n_box && lg < 30) ++lg;
    int bits = coarse ? lg - 5 : lg + 4;
    if (bits > 30) bits = 30;
    int passes = (bits + 7) / 8;
    if (passes < 2) passes = 2;
    if (passes > 4) passes = 4;
    // With migration the step loop re-sorts finely on every migration step (24 key bits + the marker pass, every k-th
    // step); in between, atoms move a small fraction of a leaf, so two passes over the top 16 bits — which put an atom that
    // crossed a coarse cell border into its new cell and keep the order inside the cells — hold the order at any system
    // size (a third pass costs 0.05 ms per step at >= 4 M atoms in the box).
    if (coarse && h->mg_migrate_every > 0 && h->mg_world > 1) passes = migrated_now ? 3 : 2;
    const int low_bit = passes == 4 ? 0 : 30 - 8 * passes;
    if (migrated_now && passes < 4) ++passes;  // bits [30, 32): the leaver marker
    int buf = 0;
    {
        StageScope sc(h, NB200_STAGE_SORT);
        const bool clean = h->hk_sort_clean && h->hk_n == n_pre && h->hk_passes == passes;
        h->hk_sort_clean = false;
        sc.add(launch_sort(sA, h->keys, h->vals, n_pre, h->sort_hist, h->sort_status, h->sort_ticket, &buf, low_bit, passes, clean));
        CHECK_LAUNCH(h, "sort(owned)");
        h->mg_sorted_keys = h->keys[buf];  // slot s of the owned segment has this key (strays keep theirs, integrate_kernel<PUBLISH>)
    }
    {
        StageScope sc(h, NB200_STAGE_REORDER);
        Housekeeping hk = sort_housekeeping(n_own, passes, h->sort_hist, h->sort_status, h->sort_ticket);
        hk.node_flag = h->node_flag;
        hk.n_flag = nLo > 1 ? nLo - 1 : 0;
        hk.counters = reinterpret_cast<uint32_t*>(h->counters);
        hk.n_counter_words = (int)(COUNTERS_RESET_BYTES / 4);
        sc.add(launch_reorder(sA, h->vals[buf], h->keys[buf], h->pos[src], h->vel[src], h->id[src], h->pos[dst], h->vel[dst], h->id[dst],
                              h->force, h->leaf_lo, h->leaf_hi, h->leaf_sub, n_own, cutoff, &hk));
        CHECK_LAUNCH(h, "reorder(owned)");
        h->hk_sort_clean = !migrated_now;  // (the scratch was cleared for this step's pass count and size)
        h->hk_n = n_own;
        h->hk_passes = passes;
        if (migrated_now && gbase > n_own)  // the owned segment changed size: its pad slots must hold NaN in both buffers
            for (int b = 0; b < 2; ++b) CU(h, cudaMemsetAsync(h->pos[b] + n_own, 0xff, sizeof(float4) * (size_t)(gbase - n_own), sA));
        h->mg_sort_extra = false;
        h->mg_n_pre = n_own;
    }
    {
        StageScope sc(h, NB200_STAGE_BUILD);
        sc.add(launch_build(sA, h->leaf_lo, h->leaf_hi, nLo, h->nodes, h->node_lo, h->node_hi, h->node_flag, true, 0));
        sc.add(launch_frontier(sA, h->nodes, nLo, h->frontier, 0, h->node_lo, h->node_hi, h->leaf_lo, h->leaf_hi));
        CHECK_LAUNCH(h, "build(owned)");
    }
    mg_trace(h, 3, sA);
    h->cur = dst;
    h->n = gbase + n_g;
    h->n_leaves = nLo + nLg;
    h->mg_n_gslots = n_g;
    h->cutoff = cutoff;
    if (two_streams) CU(h, cudaEventRecord(h->mg_ev_owned, sA));  // owned positions, leaf boxes and the cleared counters are in place
    int32_t rc = mg_launch_traverse(h, fused, true, 1, sA);  // does not need the ghosts: the halo exchange hides under it
    if (rc) return rc;
    mg_trace(h, 4, sA);
    // The ghost side — sort, gather, tree of the ghosts the pull delivered, then the ghost pass — runs BESIDE the owned
    // pass on the ghost stream and only starts when the owned pass has been launched: its dozen small, latency-bound
    // kernels would otherwise slow down the (equally latency-bound) owned sort and tree build, which ARE on the
    // critical path.  Both passes append to the same tile list with atomic reservations and add forces with reductions.
    if (two_streams) CU(h, cudaStreamWaitEvent(sB, h->mg_ev_owned, 0));
    if (before_ghost_side) {
        const int32_t rcb = (*before_ghost_side)();
        if (rcb) return rcb;
    }
    {
        StageScope sc(h, NB200_STAGE_MORTON);
        if (n_g > 0) {
            const int passes = 3, low_bit = 6;  // the ghost shell is thin and arrives in arbitrary order: 24 key bits
            int gbuf = 0;
            const bool clean = h->hk2_clean && h->hk2_n == n_g;
            h->hk2_clean = false;
            sc.add(launch_sort(sB, h->mg_gkeys, h->mg_gvals, n_g, h->sort_hist2, h->sort_status2, h->sort_ticket2, &gbuf, low_bit, passes, clean));
            Housekeeping hk = sort_housekeeping(n_g, passes, h->sort_hist2, h->sort_status2, h->sort_ticket2);
            hk.node_flag = h->node_flag + nLo;
            hk.n_flag = nLg > 1 ? nLg - 1 : 0;
            sc.add(launch_reorder(sB, h->mg_gvals[gbuf], h->mg_gkeys[gbuf], h->mg_gpos, nullptr, nullptr, h->pos[dst] + gbase, nullptr,
                                  h->id[dst] + gbase, nullptr, h->leaf_lo + nLo, h->leaf_hi + nLo, h->leaf_sub + (size_t)nLo * 8, n_g, cutoff, &hk));
            h->hk2_clean = true;
            h->hk2_n = n_g;
            sc.add(launch_build(sB, h->leaf_lo, h->leaf_hi, nLg, h->nodes, h->node_lo, h->node_hi, h->node_flag, true, nLo));
        }
        sc.add(launch_frontier(sB, h->nodes, nLg, h->frontier2, nLo, h->node_lo, h->node_hi, h->leaf_lo, h->leaf_hi));
        CHECK_LAUNCH(h, "ghost tree");
        mg_trace(h, 2, sB);
    }
    rc = mg_launch_traverse(h, fused, true, 2, sB);
    if (rc) return rc;
    if (two_streams) {
        CU(h, cudaEventRecord(h->mg_ev_ghost, sB));
        CU(h, cudaStreamWaitEvent(sA, h->mg_ev_ghost, 0));
    }
    mg_trace(h, 5, sA);
    return rc;
}

}  // namespace

// ========================================================================================================
extern "C" {

int32_t nb200_version(void) { return 100; }

int32_t nb200_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return c;
}

const char* nb200_last_error(const nb200_handle* h) { return h ? h->err : g_create_error; }

int32_t nb200_create(int32_t device, int64_t n_max, int64_t pair_capacity_hint, nb200_handle** out) {
    if (!out) return fail(nullptr, NB200_ERR_BAD_ARG, "out pointer is NULL");
    *out = nullptr;
    if (n_max < 2) return fail(nullptr, NB200_ERR_BAD_ARG, "n_max must be >= 2");
    if (n_max >= (1ll << 30)) return fail(nullptr, NB200_ERR_BAD_ARG, "n_max must be < 2^30");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, NB200_ERR_CUDA, "no CUDA device available (%s); libnaiveb200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count) return fail(nullptr, NB200_ERR_BAD_ARG, "device %d out of range [0,%d)", device, count);
    nb200_handle* h = new (std::nothrow) nb200_handle();
    if (!h) return fail(nullptr, NB200_ERR_BAD_ARG, "out of host memory");
    std::memset(h, 0, sizeof(*h));
    h->device = device;
    h->n_max = n_max;
    h->owns_stream = true;
#define CUC(expr)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            fail(nullptr, NB200_ERR_CUDA, "%s -> %s", #expr, cudaGetErrorString(e_));                     \
            nb200_destroy(h);                                                                             \
            return NB200_ERR_CUDA;                                                                        \
        }                                                                                                 \
    } while (0)
    CUC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    CUC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const int64_t nLmax = (n_max + LEAF - 1) / LEAF;
    for (int b = 0; b < 2; ++b) {
        CUC(dalloc(&h->pos[b], n_max));
        CUC(dalloc(&h->vel[b], n_max));
        CUC(dalloc(&h->id[b], n_max));
        CUC(dalloc(&h->keys[b], n_max));
        CUC(dalloc(&h->vals[b], n_max));
    }
    CUC(dalloc(&h->force, n_max));
    CUC(cudaMemset(h->force, 0, sizeof(float4) * (size_t)n_max));
    h->sort_tiles_cap = sort_tiles(n_max);
    CUC(dalloc(&h->sort_hist, 4 * 256));
    CUC(dalloc(&h->sort_status, 4 * h->sort_tiles_cap * 256));
    CUC(dalloc(&h->sort_ticket, 4));
    CUC(dalloc(&h->leaf_lo, nLmax));
    CUC(dalloc(&h->leaf_hi, nLmax));
    CUC(dalloc(&h->leaf_sub, nLmax * 8));
    CUC(dalloc(&h->nodes, nLmax));
    CUC(dalloc(&h->node_lo, nLmax));
    CUC(dalloc(&h->node_hi, nLmax));
    CUC(dalloc(&h->node_flag, nLmax));
    CUC(dalloc(&h->frontier, FRONTIER_WORDS));
    CUC(dalloc(&h->reuse_d2, 2));
    CUC(cudaMemset(h->reuse_d2, 0, 2 * sizeof(unsigned int)));
    CUC(dalloc(&h->counters, 1));
    CUC(cudaMemset(h->counters, 0, sizeof(Counters)));
    CUC(cudaHostAlloc((void**)&h->counters_h, sizeof(Counters), cudaHostAllocDefault));
    h->stage_floats = n_max * 10;
    CUC(dalloc(&h->stage_dev, h->stage_floats));
    CUC(dalloc(&h->energy_dev, 4));
    // list slots: a half list holds one entry per pair, chunk padding adds ~50 %; a directed list twice that
    int64_t want = pair_capacity_hint > 0 ? 2 * pair_capacity_hint : 64 * n_max;
    {
        int64_t cap = want + want / 8 + 4096;
        int64_t seg_cap = seg_capacity_for(n_max, cap);
        CUC(dalloc(&h->entries, cap));
        CUC(dalloc(&h->segs, seg_cap));
        h->entry_capacity = cap;
        h->seg_capacity = seg_cap;
    }
    for (int d = 0; d < 3; ++d) { h->box_min[d] = 0.f; h->box_max[d] = 1.f; }
    h->curve = 1;
    h->list_mode = NB200_LIST_HALF;
    h->list_half = true;
    h->resort_interval = 1;
    h->reuse_every = 1;
    h->reuse_skin = 0.f;
    h->ff.eps = 1.f; h->ff.sigma = 1.f; h->ff.kcoul = 0.f; h->ff.cutoff = 2.5f; h->ff.shift = 1;
    h->fused_force = true;
    h->use_graph = std::getenv("NB200_NO_GRAPH") == nullptr;
    if (const char* sp = std::getenv("NB200_SORT_PASSES")) h->sort_passes_override = std::atoi(sp);
    h->mg_trace = std::getenv("NB200_MG_TRACE") != nullptr;
    h->mg_graph_multi = std::getenv("NB200_MG_GRAPH") != nullptr;
    h->mg_pull_late = std::getenv("NB200_MG_PULL_EARLY") == nullptr;
    h->mg_ghost_prio_normal = std::getenv("NB200_MG_PRIO_NORMAL") != nullptr;
    h->mg_ahead = 16;
    if (const char* sp = std::getenv("NB200_MG_AHEAD")) { h->mg_ahead = std::atoi(sp); if (h->mg_ahead < 1) h->mg_ahead = 1; if (h->mg_ahead > 16) h->mg_ahead = 16; }
#undef CUC
    *out = h;
    return NB200_OK;
}

int32_t nb200_destroy(nb200_handle* h) {
    if (!h) return NB200_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->graph_exec);
    for (int b = 0; b < 2; ++b) {
        cudaFree(h->pos[b]); cudaFree(h->vel[b]); cudaFree(h->id[b]); cudaFree(h->keys[b]); cudaFree(h->vals[b]);
    }
    cudaFree(h->force); cudaFree(h->sort_hist); cudaFree(h->sort_status); cudaFree(h->sort_ticket);
    cudaFree(h->leaf_lo); cudaFree(h->leaf_hi); cudaFree(h->leaf_sub); cudaFree(h->nodes); cudaFree(h->node_lo); cudaFree(h->node_hi);
    cudaFree(h->node_flag); cudaFree(h->frontier); cudaFree(h->reuse_d2); cudaFree(h->entries); cudaFree(h->segs); cudaFree(h->counters); cudaFree(h->stage_dev);
    cudaFree(h->scratch_dev); cudaFree(h->exp_a); cudaFree(h->exp_b); cudaFree(h->exp_d); cudaFree(h->energy_dev);
    if (h->counters_h) cudaFreeHost(h->counters_h);
    if (h->timer.created)
        for (int i = 0; i < StageTimer::MAX_EVENTS; ++i) cudaEventDestroy(h->timer.ev[i]);
    if (h->sw_created) { cudaEventDestroy(h->sw_start); cudaEventDestroy(h->sw_stop); }
    if (h->log_created) {
        for (int k = 0; k < 3; ++k) { cudaFree(h->log_stage[k]); cudaEventDestroy(h->log_ready[k]); cudaEventDestroy(h->log_copied[k]); }
        cudaStreamDestroy(h->log_stream);
    }
    if (h->stream && h->owns_stream) cudaStreamDestroy(h->stream);
    mg_close_peers(h);
    cudaFree(h->mg_pub); cudaFree(h->mg_box); cudaFree(h->mg_gpos); cudaFree(h->mg_ggidx); cudaFree(h->mg_sendbuf); cudaFree(h->mg_split);
    for (int b = 0; b < 2; ++b) { cudaFree(h->mg_gkeys[b]); cudaFree(h->mg_gvals[b]); }
    cudaFree(h->sort_hist2); cudaFree(h->sort_status2); cudaFree(h->sort_ticket2); cudaFree(h->frontier2); cudaFree(h->mg_blist); cudaFree(h->mg_bcount);
    if (h->mg_stream2) { cudaStreamDestroy(h->mg_stream2); cudaEventDestroy(h->mg_ev_int); cudaEventDestroy(h->mg_ev_ghost); cudaEventDestroy(h->mg_ev_owned); }
    cudaFree(h->mg_ghost_count); cudaFree(h->mg_err); cudaFree(h->mg_peers_dev); cudaFree(h->mg_ghost_stat); cudaFree(h->mg_grid);
    if (h->mg_ghost_count_h) cudaFreeHost(h->mg_ghost_count_h);
    if (h->mg_stat_h) cudaFreeHost(h->mg_stat_h);
    if (h->mg_ev_created)
        for (int k = 0; k < 16; ++k) cudaEventDestroy(h->mg_step_ev[k]);
    cudaGetLastError();
    delete h;
    return NB200_OK;
}

int32_t nb200_set_box(nb200_handle* h, const float box_min[3], const float box_max[3]) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!box_min || !box_max) return fail(h, NB200_ERR_BAD_ARG, "box pointers are NULL");
    for (int d = 0; d < 3; ++d)
        if (!(box_max[d] > box_min[d])) return fail(h, NB200_ERR_BAD_ARG, "box_max[%d] must exceed box_min[%d]", d, d);
    for (int d = 0; d < 3; ++d) { h->box_min[d] = box_min[d]; h->box_max[d] = box_max[d]; }
    return NB200_OK;
}

// ---- neighbour search --------------------------------------------------------------------------------------
int32_t nb200_neighbors(nb200_handle* h, const float* xyz, int32_t stride, int32_t n, float cutoff, int64_t* pair_count) {
    if (!h) return NB200_ERR_BAD_ARG;
    h->mg_active = false;
    if (!xyz) return fail(h, NB200_ERR_BAD_ARG, "xyz is NULL");
    if (stride != 3 && stride != 4) return fail(h, NB200_ERR_BAD_ARG, "stride must be 3 or 4");
    if (!(cutoff >= 0.f)) return fail(h, NB200_ERR_BAD_ARG, "cutoff must be >= 0");
    int32_t rc = check_n(h, n);
    if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    reset_async_reports(h);
    h->have_system = false;
    h->have_forces = false;
    h->list_valid = false;
    h->vel_half = false;
    rc = upload_system(h, xyz, nullptr, stride, nullptr, nullptr, n, false);
    if (rc) return rc;
    {
        StageScope sc(h, NB200_STAGE_MORTON);
        sc.add(launch_morton(h->stream, h->pos[h->cur], n, h->box_min, h->box_max, h->keys[0], h->vals[0], h->curve));
        CHECK_LAUNCH(h, "morton");
    }
    rc = search_sync(h, false, cutoff);
    if (rc) return rc;
    if (pair_count) *pair_count = list_pairs(h);
    return NB200_OK;
}

int32_t nb200_pair_count(nb200_handle* h, int64_t* pair_count) {
    if (!h || !pair_count) return NB200_ERR_BAD_ARG;
    if (!h->list_valid) return fail(h, NB200_ERR_STATE, "no neighbour list (call nb200_neighbors / nb200_set_system first)");
    CU(h, cudaSetDevice(h->device));
    int32_t rc = read_counters(h);
    if (rc) return rc;
    *pair_count = list_pairs(h);
    return NB200_OK;
}

int32_t nb200_get_pairs(nb200_handle* h, int32_t* a, int32_t* b, float* d, int64_t capacity, int32_t index_base,
                        int64_t* written) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->list_valid) return fail(h, NB200_ERR_STATE, "no neighbour list (call nb200_neighbors first)");
    CU(h, cudaSetDevice(h->device));
    int32_t rc = read_counters(h);
    if (rc) return rc;
    const int64_t np = list_pairs(h);
    if (written) *written = np;
    if (np == 0) return NB200_OK;
    if (capacity < np) return fail(h, NB200_ERR_CAPACITY, "pair buffers hold %lld, list has %lld", (long long)capacity, (long long)np);
    if (!a || !b || !d) return fail(h, NB200_ERR_BAD_ARG, "output pointers are NULL");
    rc = export_device(h, np, index_base);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(a, h->exp_a, sizeof(int32_t) * (size_t)np, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(b, h->exp_b, sizeof(int32_t) * (size_t)np, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(d, h->exp_d, sizeof(float) * (size_t)np, cudaMemcpyDeviceToHost, h->stream));
    rc = read_counters(h);
    if (rc) return rc;
    if ((int64_t)h->counters_h->n_export != np)
        return fail(h, NB200_ERR_STATE, "export produced %llu pairs, expected %lld", (unsigned long long)h->counters_h->n_export, (long long)np);
    return NB200_OK;
}

// ---- literal reference entry points ----------------------------------------------------------------------------
int32_t nb200_force_lennardjones(nb200_handle* h, float* force, int32_t n, const int32_t* a, const int32_t* b, const float* d,
                                 int64_t npairs, int32_t index_base) {
    (void)b;
    if (!h) return NB200_ERR_BAD_ARG;
    if (!force || n < 1 || npairs < 0 || (npairs > 0 && (!a || !d))) return fail(h, NB200_ERR_BAD_ARG, "bad arguments");
    CU(h, cudaSetDevice(h->device));
    int64_t bytes = npairs * 8 + (int64_t)n * 8 + (int64_t)n * 12 + 64;
    int32_t rc = ensure_scratch(h, bytes);
    if (rc) return rc;
    char* p = (char*)h->scratch_dev;
    double* acc = (double*)p; p += (size_t)n * 8;
    int32_t* da = (int32_t*)p; p += (size_t)npairs * 4;
    float* dd = (float*)p; p += (size_t)npairs * 4;
    float* df = (float*)p;
    if (npairs > 0) {
        CU(h, cudaMemcpyAsync(da, a, (size_t)npairs * 4, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemcpyAsync(dd, d, (size_t)npairs * 4, cudaMemcpyHostToDevice, h->stream));
    }
    h->kernel_launches += launch_lj_literal(h->stream, da, dd, npairs, index_base, n, acc, df);
    CHECK_LAUNCH(h, "lj_literal");
    CU(h, cudaMemcpyAsync(force, df, (size_t)n * 12, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

int32_t nb200_force_coulomb(nb200_handle* h, float* force, int32_t n, const int32_t* a, const int32_t* b, const float* d,
                            int64_t npairs, const float* charge, int32_t index_base) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!force || !charge || n < 1 || npairs < 0 || (npairs > 0 && (!a || !b || !d))) return fail(h, NB200_ERR_BAD_ARG, "bad arguments");
    CU(h, cudaSetDevice(h->device));
    int64_t bytes = npairs * 12 + (int64_t)n * 16 + 64;
    int32_t rc = ensure_scratch(h, bytes);
    if (rc) return rc;
    char* p = (char*)h->scratch_dev;
    int32_t* da = (int32_t*)p; p += (size_t)npairs * 4;
    int32_t* db = (int32_t*)p; p += (size_t)npairs * 4;
    float* dd = (float*)p; p += (size_t)npairs * 4;
    float* dq = (float*)p; p += (size_t)n * 4;
    float* df = (float*)p;
    if (npairs > 0) {
        CU(h, cudaMemcpyAsync(da, a, (size_t)npairs * 4, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemcpyAsync(db, b, (size_t)npairs * 4, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemcpyAsync(dd, d, (size_t)npairs * 4, cudaMemcpyHostToDevice, h->stream));
    }
    CU(h, cudaMemcpyAsync(dq, charge, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    h->kernel_launches += launch_coulomb_literal(h->stream, da, db, dd, npairs, index_base, dq, n, df);
    CHECK_LAUNCH(h, "coulomb_literal");
    CU(h, cudaMemcpyAsync(force, df, (size_t)n * 12, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

int32_t nb200_sum_forces(nb200_handle* h, float* force, const float* force1, const float* force2, int64_t n3) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!force || !force1 || !force2 || n3 < 1) return fail(h, NB200_ERR_BAD_ARG, "bad arguments");
    CU(h, cudaSetDevice(h->device));
    int32_t rc = ensure_scratch(h, n3 * 12 + 64);
    if (rc) return rc;
    float* f1 = (float*)h->scratch_dev;
    float* f2 = f1 + n3;
    float* fo = f2 + n3;
    CU(h, cudaMemcpyAsync(f1, force1, (size_t)n3 * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(f2, force2, (size_t)n3 * 4, cudaMemcpyHostToDevice, h->stream));
    h->kernel_launches += launch_sum_forces(h->stream, fo, f1, f2, n3);
    CHECK_LAUNCH(h, "sum_forces");
    CU(h, cudaMemcpyAsync(force, fo, (size_t)n3 * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

int32_t nb200_verlet_update(nb200_handle* h, float* pos, float* vel, const float* force, const float* force_next,
                            const float* mass, int32_t n, float dt, const float box_min[3], const float box_max[3]) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!pos || !vel || !force || !force_next || !mass || n < 1) return fail(h, NB200_ERR_BAD_ARG, "bad arguments");
    CU(h, cudaSetDevice(h->device));
    const size_t n3 = (size_t)n * 3;
    int32_t rc = ensure_scratch(h, (int64_t)(n3 * 4 * 4 + (size_t)n * 4 + 64));
    if (rc) return rc;
    float* dp = (float*)h->scratch_dev;
    float* dv = dp + n3;
    float* df = dv + n3;
    float* dn = df + n3;
    float* dm = dn + n3;
    CU(h, cudaMemcpyAsync(dp, pos, n3 * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(dv, vel, n3 * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(df, force, n3 * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(dn, force_next, n3 * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(dm, mass, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    const int reflect = (box_min && box_max) ? 1 : 0;
    h->kernel_launches += launch_verlet_literal(h->stream, dp, dv, df, dn, dm, n, dt, box_min, box_max, reflect);
    CHECK_LAUNCH(h, "verlet_literal");
    CU(h, cudaMemcpyAsync(pos, dp, n3 * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(vel, dv, n3 * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

// ---- MD system -----------------------------------------------------------------------------------------------------
int32_t nb200_set_forcefield(nb200_handle* h, float eps, float sigma, float kcoul, float cutoff, int32_t shift) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!(sigma > 0.f) || !(cutoff > 0.f)) return fail(h, NB200_ERR_BAD_ARG, "sigma and cutoff must be > 0");
    h->ff.eps = eps; h->ff.sigma = sigma; h->ff.kcoul = kcoul; h->ff.cutoff = cutoff; h->ff.shift = shift ? 1 : 0;
    h->have_forces = false;
    return NB200_OK;
}

}  // extern "C"
namespace {
int32_t compute_forces_sync(nb200_handle* h) {
    {
        StageScope sc(h, NB200_STAGE_MORTON);
        sc.add(launch_morton(h->stream, h->pos[h->cur], h->n, h->box_min, h->box_max, h->keys[0], h->vals[0], h->curve));
        CHECK_LAUNCH(h, "morton");
    }
    int32_t rc = search_sync(h, true, md_list_cutoff(h), true);
    if (rc) return rc;
    rc = mark_list_built(h);
    if (rc) return rc;
    rc = enqueue_force(h, true);
    if (rc) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    h->have_forces = true;
    return NB200_OK;
}
}  // namespace
extern "C" {

int32_t nb200_set_system(nb200_handle* h, const float* xyz, const float* vel, int32_t stride, const float* mass,
                         const float* charge, int32_t n) {
    if (!h) return NB200_ERR_BAD_ARG;
    h->mg_active = false;
    if (!xyz) return fail(h, NB200_ERR_BAD_ARG, "xyz is NULL");
    if (stride != 3 && stride != 4) return fail(h, NB200_ERR_BAD_ARG, "stride must be 3 or 4");
    int32_t rc = check_n(h, n);
    if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    reset_async_reports(h);
    h->list_valid = false;
    h->have_forces = false;
    h->vel_half = false;
    rc = upload_system(h, xyz, vel, stride, mass, charge, n, true);
    if (rc) return rc;
    h->have_system = true;
    return compute_forces_sync(h);
}

// ---- system setup on the device (MDInput.jl:175-190, 228-283, 305-369) ----------------------------------
int32_t nb200_collect_objects(nb200_handle* h, int32_t n, uint64_t seed, float minmass, float maxmass, float mincharge,
                              float maxcharge, float temperature, int32_t randomvelocity, float minimumdistance,
                              int32_t max_rounds, float* mass_out, float* charge_out, int32_t* rounds, int64_t* redrawn) {
    if (!h) return NB200_ERR_BAD_ARG;
    h->mg_active = false;
    int32_t rc = check_n(h, n);
    if (rc) return rc;
    if (!(maxmass >= minmass) || !(maxcharge >= mincharge)) return fail(h, NB200_ERR_BAD_ARG, "Uniform(a, b) needs a <= b");
    if (!(minimumdistance >= 0.f)) return fail(h, NB200_ERR_BAD_ARG, "minimumdistance must be >= 0");
    CU(h, cudaSetDevice(h->device));
    reset_async_reports(h);
    h->list_valid = false;
    h->have_system = false;
    h->have_forces = false;
    h->vel_half = false;
    // the staging layout of nb200_set_system, filled by kernels instead of copies
    float* sx = h->stage_dev;
    float* sv = sx + (size_t)n * 4;
    float* sm = sv + (size_t)n * 4;
    float* sq = sm + (size_t)n;
    rc = ensure_scratch(h, (int64_t)n * 4 + 64);
    if (rc) return rc;
    double* sums = (double*)h->scratch_dev;                               // 3 veldist sums, then the re-draw counter
    unsigned long long* counter = (unsigned long long*)h->scratch_dev + 4;
    int32_t* mark = (int32_t*)((char*)h->scratch_dev + 64);
    h->kernel_launches += launch_setup_draw(h->stream, n, seed, h->box_min, h->box_max, minmass, maxmass, mincharge, maxcharge,
                                            temperature, randomvelocity, sx, sv, sm, sq, sums);
    CHECK_LAUNCH(h, "setup_draw");
    h->n = n;
    h->n_leaves = (n + LEAF - 1) / LEAF;
    int32_t n_rounds = 0;
    int64_t n_redrawn = 0;
    if (minimumdistance > 0.f && n > 1) {
        const int64_t limit = max_rounds > 0 ? (int64_t)max_rounds : 10 * (int64_t)n;   // recursion_limit (MDInput.jl:266)
        CU(h, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), h->stream));
        CU(h, cudaMemsetAsync(mark, 0, sizeof(int32_t) * (size_t)n, h->stream));
        for (;;) {
            // the too-close pairs are the neighbour list at cutoff = minimumdistance
            h->cur = 0;
            h->kernel_launches += launch_pack(h->stream, sx, 4, nullptr, nullptr, nullptr, n, h->pos[0], nullptr, h->id[0]);
            CHECK_LAUNCH(h, "pack");
            h->kernel_launches += launch_morton(h->stream, h->pos[0], n, h->box_min, h->box_max, h->keys[0], h->vals[0], h->curve);
            CHECK_LAUNCH(h, "morton");
            rc = search_sync(h, false, minimumdistance);
            if (rc) return rc;
            h->list_valid = false;
            const int64_t np = list_pairs(h);
            if (np == 0) break;
            if (n_rounds >= limit)
                return fail(h, NB200_ERR_STATE,
                            "Objects could not be placed, increase box size, reduce object count, or decrease minimum spawning "
                            "distance (%lld pairs still too close after %d rounds)", (long long)np, n_rounds);
            rc = export_device(h, np, 0);
            if (rc) return rc;
            ++n_rounds;
            h->kernel_launches += launch_prune_redraw(h->stream, n, seed, (uint32_t)n_rounds, h->box_min, h->box_max, h->exp_a,
                                                      h->exp_b, np, mark, sx, counter);
            CHECK_LAUNCH(h, "prune_redraw");
        }
        unsigned long long total = 0;
        CU(h, cudaMemcpyAsync(&total, counter, sizeof(total), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        n_redrawn = (int64_t)total;
    }
    if (rounds) *rounds = n_rounds;
    if (redrawn) *redrawn = n_redrawn;
    if (mass_out) CU(h, cudaMemcpyAsync(mass_out, sm, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    if (charge_out) CU(h, cudaMemcpyAsync(charge_out, sq, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    h->cur = 0;
    h->kernel_launches += launch_pack(h->stream, sx, 4, sv, sm, sq, n, h->pos[0], h->vel[0], h->id[0]);
    CHECK_LAUNCH(h, "pack");
    h->have_system = true;
    return compute_forces_sync(h);
}

}  // extern "C"
namespace {

// one step of the device loop: kick-drift(+reflect, +keys) -> [sort -> reorder -> build -> traverse(+forces)] -> forces
int32_t step_once(nb200_handle* h, float dt) {
    const float kick_dt = h->vel_half ? 0.5f * (h->last_dt + dt) : 0.5f * dt;
    {
        StageScope sc(h, NB200_STAGE_INTEGRATE);
        sc.add(launch_integrate(h->stream, h->pos[h->cur], h->vel[h->cur], h->force, h->n, kick_dt, dt, h->box_min,
                                h->box_max, h->keys[0], h->vals[0], h->curve));
        CHECK_LAUNCH(h, "integrate");
    }
    h->vel_half = true;
    h->last_dt = dt;
    int32_t rc;
    bool fused = false;
    if (h->reuse_every > 1 && h->list_valid && h->list_age + 1 < h->reuse_every) {
        // LIST REUSE: the skin list of an earlier step is still good as long as no atom moved more than skin/2
        // since it was built (checked on the device, reported by nb200_sync); only the forces are recomputed
        ++h->list_age;
        h->kernel_launches += launch_displacement_check(h->stream, h->pos[h->cur], h->vel[h->cur ^ 1], h->n,
                                                        0.25f * h->reuse_skin * h->reuse_skin, h->reuse_d2);
        CU(h, cudaMemsetAsync(h->force, 0, sizeof(float4) * (size_t)h->n, h->stream));
    } else {
        const bool resort = h->resort_interval <= 1 || h->steps_since_sort + 1 >= h->resort_interval;
        fused = step_fuses_forces(h);
        rc = enqueue_search(h, true, md_list_cutoff(h), resort, fused, true);
        if (rc) return rc;
        h->list_valid = true;
        rc = mark_list_built(h);
        if (rc) return rc;
    }
    if (!fused) {
        rc = enqueue_force(h, false);  // energies are recomputed on demand (nb200_get_energies)
        if (rc) return rc;
    }
    h->steps_done++;
    h->async_overflow_possible = true;
    return NB200_OK;
}

// ---- the step loop as a CUDA graph -------------------------------------------------------------------------------
// A step is ~9 short launches; back to back on a stream each costs 1-2 us of GPU idle.  In steady state (same dt,
// list rebuilt every step, no per-stage events) two consecutive steps are an exact period of the loop — the state
// buffers alternate (h->cur) and every kernel argument repeats — so the pair is captured once and replayed.
struct GraphKey {
    int32_t n, cur, fused, list_mode, curve, passes_n;
    float dt, cutoff;
    ForceField ff;
    float box[6];
    const void *entries, *segs;
    int64_t entry_capacity, seg_capacity;
};

void graph_drop(nb200_handle* h) {
    if (h->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->graph_exec);
    h->graph_exec = nullptr;
}

bool graph_steady(const nb200_handle* h, float dt) {
    return h->use_graph && !h->timer.enabled && !h->mg_active && h->reuse_every <= 1 && h->resort_interval <= 1 && h->vel_half &&
           h->last_dt == dt && h->list_valid && h->hk_sort_clean && h->hk_n == h->n;
}

GraphKey graph_key(const nb200_handle* h, float dt) {
    GraphKey k;
    std::memset(&k, 0, sizeof(k));
    k.n = h->n; k.cur = h->cur; k.fused = step_fuses_forces(h) ? 1 : 0; k.list_mode = h->list_mode; k.curve = h->curve;
    k.passes_n = h->hk_passes;
    k.dt = dt; k.cutoff = md_list_cutoff(h);
    k.ff = h->ff;
    for (int d = 0; d < 3; ++d) { k.box[d] = h->box_min[d]; k.box[3 + d] = h->box_max[d]; }
    k.entries = h->entries; k.segs = h->segs;
    k.entry_capacity = h->entry_capacity; k.seg_capacity = h->seg_capacity;
    return k;
}

// captures two steps into h->graph_exec (nothing executes); host-side bookkeeping of the two steps is rolled back
int32_t graph_capture(nb200_handle* h, float dt) {
    graph_drop(h);
    const int64_t launches0 = h->kernel_launches, steps0 = h->steps_done;
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return NB200_ERR_CUDA; }
    int32_t rc = step_once(h, dt);
    if (!rc) rc = step_once(h, dt);
    const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    h->graph_launches = h->kernel_launches - launches0;
    h->kernel_launches = launches0;
    h->steps_done = steps0;
    if (rc || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return rc ? rc : NB200_ERR_CUDA;
    }
    cudaGraphExec_t ex = nullptr;
    const cudaError_t e2 = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) { cudaGetLastError(); return NB200_ERR_CUDA; }
    h->graph_exec = ex;
    return NB200_OK;
}

}  // namespace
extern "C" {

int32_t nb200_step_async(nb200_handle* h, int32_t nsteps, float dt) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->have_system) return fail(h, NB200_ERR_STATE, "no system loaded (call nb200_set_system first)");
    if (nsteps < 0) return fail(h, NB200_ERR_BAD_ARG, "nsteps must be >= 0");
    CU(h, cudaSetDevice(h->device));
    if (!h->have_forces) {
        int32_t rc = compute_forces_sync(h);
        if (rc) return rc;
    }
    int32_t s = 0;
    while (s < nsteps) {
        if (nsteps - s >= 4 && graph_steady(h, dt)) {
            const GraphKey key = graph_key(h, dt);
            static_assert(sizeof(GraphKey) <= sizeof(h->graph_key), "graph key storage");
            if (!h->graph_exec || h->graph_is_mg || std::memcmp(&key, h->graph_key, sizeof(key)) != 0) {
                if (graph_capture(h, dt) == NB200_OK) { std::memcpy(h->graph_key, &key, sizeof(key)); h->graph_is_mg = false; }
                else h->use_graph = false;  // capture unavailable: plain launches from here on
            }
            if (h->graph_exec) {
                const int32_t pairs = (nsteps - s) / 2;
                for (int32_t k = 0; k < pairs; ++k) CU(h, cudaGraphLaunch((cudaGraphExec_t)h->graph_exec, h->stream));
                h->kernel_launches += (int64_t)pairs * h->graph_launches;
                h->steps_done += 2 * (int64_t)pairs;
                h->async_overflow_possible = true;
                h->pe_valid = false;  // (the bookkeeping step_once does per step, for the replayed steps)
                h->list_valid = true;
                s += 2 * pairs;
                continue;
            }
        }
        int32_t rc = step_once(h, dt);
        if (rc) return rc;
        ++s;
    }
    return NB200_OK;
}

int32_t nb200_sync(nb200_handle* h) {
    if (!h) return NB200_ERR_BAD_ARG;
    CU(h, cudaSetDevice(h->device));
    int32_t rc = read_counters(h);
    if (rc) return rc;
    if (h->reuse_every > 1 && !h->mg_active) {
        unsigned int st[2] = {0, 0};
        CU(h, cudaMemcpy(st, h->reuse_d2, sizeof(st), cudaMemcpyDeviceToHost));
        if (st[1]) {
            CU(h, cudaMemset(h->reuse_d2, 0, sizeof(st)));
            float d2;
            std::memcpy(&d2, &st[0], 4);
            h->list_valid = false;
            h->have_forces = false;
            return fail(h, NB200_ERR_STATE, "list reuse: an atom moved %.3g since the list was built, more than skin/2 = %.3g — pairs may "
                        "have been missed; use a larger skin or a shorter interval (nb200_set_list_reuse)", sqrtf(d2), 0.5f * h->reuse_skin);
        }
    }
    if (h->async_overflow_possible) {
        h->async_overflow_possible = false;
        if (h->counters_h->overflow_sticky) {
            unsigned long long need = h->counters_h->n_entries();
            CU(h, cudaMemsetAsync(&h->counters->overflow_sticky, 0, sizeof(unsigned int), h->stream));
            h->have_forces = false;
            h->list_valid = false;
            ensure_entries(h, (int64_t)need * 2);
            return fail(h, NB200_ERR_PAIR_OVERFLOW,
                        "neighbour buffer overflowed during the step loop (needed >= %llu entries); buffer regrown — reload the system and retry",
                        need);
        }
    }
    return NB200_OK;
}

int32_t nb200_step(nb200_handle* h, int32_t nsteps, float dt) {
    int32_t rc = nb200_step_async(h, nsteps, dt);
    if (rc) return rc;
    return nb200_sync(h);
}

int32_t nb200_step_host(nb200_handle* h, float* xyz, float* vel, int32_t stride, int32_t n, int32_t nsteps, float dt) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->have_system || n != h->n) return fail(h, NB200_ERR_STATE, "nb200_step_host needs nb200_set_system with the same n first (mass/charge come from it)");
    if (!xyz) return fail(h, NB200_ERR_BAD_ARG, "xyz is NULL");
    if (stride != 3 && stride != 4) return fail(h, NB200_ERR_BAD_ARG, "stride must be 3 or 4");
    CU(h, cudaSetDevice(h->device));
    // positions (and velocities) come from the caller in ORIGINAL order; mass/charge stay resident:
    // scatter them back through id[] so the sorted state matches the caller's arrays.
    float* sx = h->stage_dev;
    float* sv = sx + (size_t)n * 4;
    CU(h, cudaMemcpyAsync(sx, xyz, sizeof(float) * (size_t)n * stride, cudaMemcpyHostToDevice, h->stream));
    if (vel) CU(h, cudaMemcpyAsync(sv, vel, sizeof(float) * (size_t)n * stride, cudaMemcpyHostToDevice, h->stream));
    h->kernel_launches += launch_refresh(h->stream, sx, vel ? sv : nullptr, stride, h->id[h->cur], n, h->pos[h->cur], h->vel[h->cur]);
    CHECK_LAUNCH(h, "refresh");
    h->vel_half = vel ? false : h->vel_half;
    int32_t rc = compute_forces_sync(h);
    if (rc) return rc;
    rc = nb200_step(h, nsteps, dt);
    if (rc) return rc;
    rc = download_vec(h, xyz, stride, 0);
    if (rc) return rc;
    if (vel) rc = download_vec(h, vel, stride, 1);
    return rc;
}

// Leapfrog-ordered host-buffer step, fully asynchronous: one neighbour search per call, no host
// synchronisation inside, so calls on DIFFERENT handles overlap their PCIe copies with each other's kernels.
int32_t nb200_leapfrog_host_async(nb200_handle* h, float* xyz, float* vel, int32_t stride, int32_t n, float dt,
                                  int32_t vel_is_half_step) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->have_system || n != h->n)
        return fail(h, NB200_ERR_STATE, "nb200_leapfrog_host_async needs nb200_set_system with the same n first (mass/charge come from it)");
    if (h->mg_active) return fail(h, NB200_ERR_STATE, "handle is in multi-GPU mode");
    if (!xyz) return fail(h, NB200_ERR_BAD_ARG, "xyz is NULL");
    if (stride != 3 && stride != 4) return fail(h, NB200_ERR_BAD_ARG, "stride must be 3 or 4");
    CU(h, cudaSetDevice(h->device));
    // vel == NULL: POSITIONS-ONLY exchange — the caller owns the positions (what simulate!'s poslog contract moves per
    // step, Simulator.jl:245), the velocities stay resident on the device at their half step; half the PCIe bytes.
    if (!vel) vel_is_half_step = h->vel_half ? 1 : 0;
    float* sx = h->stage_dev;
    float* sv = sx + (size_t)n * 4;
    const size_t bytes = sizeof(float) * (size_t)n * stride;
    CU(h, cudaMemcpyAsync(sx, xyz, bytes, cudaMemcpyHostToDevice, h->stream));
    if (vel) CU(h, cudaMemcpyAsync(sv, vel, bytes, cudaMemcpyHostToDevice, h->stream));
    {
        StageScope sc(h, NB200_STAGE_MORTON);  // scatter the uploaded state into the sorted slots + curve keys, one pass
        sc.add(launch_refresh(h->stream, sx, vel ? sv : nullptr, stride, h->id[h->cur], n, h->pos[h->cur], h->vel[h->cur], h->box_min,
                              h->box_max, h->curve, h->keys[0], h->vals[0]));
        CHECK_LAUNCH(h, "refresh");
    }
    const bool fused = h->fused_force && !(h->ff.eps == 0.f && h->ff.kcoul == 0.f);
    int32_t rc = enqueue_search(h, true, h->ff.cutoff, true, fused, true);
    if (rc) return rc;
    if (!fused) {
        rc = enqueue_force(h, false);
        if (rc) return rc;
    }
    {
        StageScope sc(h, NB200_STAGE_INTEGRATE);
        sc.add(launch_integrate(h->stream, h->pos[h->cur], h->vel[h->cur], h->force, n, vel_is_half_step ? dt : 0.5f * dt, dt,
                                h->box_min, h->box_max, h->keys[0], h->vals[0], h->curve));
        CHECK_LAUNCH(h, "integrate");
    }
    // the staged inputs were consumed by refresh_kernel (stream order): reuse the staging area for the outputs
    if (vel) h->kernel_launches += launch_unpack_state(h->stream, h->pos[h->cur], h->vel[h->cur], h->id[h->cur], n, stride, sx, sv);
    else h->kernel_launches += launch_unpack(h->stream, h->pos[h->cur], h->id[h->cur], n, stride, sx, 0, nullptr, 0.f);
    CHECK_LAUNCH(h, "unpack");
    CU(h, cudaMemcpyAsync(xyz, sx, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (vel) CU(h, cudaMemcpyAsync(vel, sv, bytes, cudaMemcpyDeviceToHost, h->stream));
    // positions moved after the force pass: list and forces now belong to x(t), velocities sit at t + dt/2
    h->vel_half = true;
    h->last_dt = dt;
    h->have_forces = false;
    h->list_valid = false;
    h->steps_done++;
    h->async_overflow_possible = true;
    return NB200_OK;
}

int32_t nb200_rescale_velocity(nb200_handle* h, float target_temperature, float gamma, int32_t physical) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->have_system || h->mg_active) return fail(h, NB200_ERR_STATE, "no system loaded (call nb200_set_system first)");
    CU(h, cudaSetDevice(h->device));
    const bool pending = h->vel_half;
    if (pending && !h->have_forces) {  // the closing half kick needs F at the current positions
        int32_t rc = compute_forces_sync(h);
        if (rc) return rc;
    }
    h->kernel_launches += launch_rescale_velocity(h->stream, h->vel[h->cur], h->force, h->n, pending ? 0.5f * h->last_dt : 0.f,
                                                  target_temperature, gamma, physical ? 1 : 0, h->energy_dev + 2);
    CHECK_LAUNCH(h, "rescale_velocity");
    h->vel_half = false;  // velocities are synchronised with the positions now
    return NB200_OK;
}

int32_t nb200_simulate(nb200_handle* h, int32_t nsteps, float dt, int32_t log_every, float* poslog, int32_t stride,
                       int64_t frame_capacity, int32_t rescale_every, float target_temperature, float gamma,
                       int64_t* frames_written) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->have_system || h->mg_active) return fail(h, NB200_ERR_STATE, "no system loaded (call nb200_set_system first)");
    if (nsteps < 0 || log_every < 0 || rescale_every < 0) return fail(h, NB200_ERR_BAD_ARG, "nsteps, log_every and rescale_every must be >= 0");
    if (stride != 3 && stride != 4) return fail(h, NB200_ERR_BAD_ARG, "stride must be 3 or 4");
    const int64_t frames = log_every > 0 ? nsteps / log_every : 0;
    if (frames > 0 && (!poslog || frame_capacity < frames))
        return fail(h, NB200_ERR_CAPACITY, "poslog holds %lld frames, the run logs %lld", (long long)frame_capacity, (long long)frames);
    CU(h, cudaSetDevice(h->device));
    const int64_t frame_floats = (int64_t)h->n * stride;
    if (frames > 0 && (!h->log_created || h->log_stage_floats < frame_floats)) {
        if (!h->log_created) {
            CU(h, cudaStreamCreateWithFlags(&h->log_stream, cudaStreamNonBlocking));
            for (int k = 0; k < 3; ++k) {
                CU(h, cudaEventCreateWithFlags(&h->log_ready[k], cudaEventDisableTiming));
                CU(h, cudaEventCreateWithFlags(&h->log_copied[k], cudaEventDisableTiming));
                h->log_stage[k] = nullptr;
            }
            h->log_created = true;
        }
        for (int k = 0; k < 3; ++k) {
            cudaFree(h->log_stage[k]);
            CU(h, dalloc(&h->log_stage[k], h->n_max * 4));
        }
        h->log_stage_floats = h->n_max * 4;
    }
    int64_t frame = 0;
    for (int32_t s = 1; s <= nsteps; ++s) {
        int32_t rc = nb200_step_async(h, 1, dt);
        if (rc) return rc;
        if (rescale_every > 0 && s % rescale_every == 0) {  // Simulator.jl:241-243
            rc = nb200_rescale_velocity(h, target_temperature, gamma, 0);
            if (rc) return rc;
        }
        if (log_every > 0 && s % log_every == 0) {  // push!(poslog, deepcopy(sys.position))  (Simulator.jl:245)
            const int slot = (int)(frame % 3);
            if (frame >= 3) CU(h, cudaStreamWaitEvent(h->stream, h->log_copied[slot], 0));  // staging slot free again
            h->kernel_launches += launch_unpack(h->stream, h->pos[h->cur], h->id[h->cur], h->n, stride, h->log_stage[slot], 0, nullptr, 0.f);
            CHECK_LAUNCH(h, "unpack(log)");
            CU(h, cudaEventRecord(h->log_ready[slot], h->stream));
            CU(h, cudaStreamWaitEvent(h->log_stream, h->log_ready[slot], 0));
            CU(h, cudaMemcpyAsync(poslog + frame * frame_floats, h->log_stage[slot], sizeof(float) * (size_t)frame_floats,
                                  cudaMemcpyDeviceToHost, h->log_stream));
            CU(h, cudaEventRecord(h->log_copied[slot], h->log_stream));
            ++frame;
        }
    }
    if (frames_written) *frames_written = frame;
    if (frame > 0) CU(h, cudaStreamSynchronize(h->log_stream));
    return nb200_sync(h);
}

int32_t nb200_get_positions(nb200_handle* h, float* xyz, int32_t stride) {
    if (!h || !xyz) return NB200_ERR_BAD_ARG;
    CU(h, cudaSetDevice(h->device));
    return download_vec(h, xyz, stride, 0);
}
int32_t nb200_get_velocities(nb200_handle* h, float* vel, int32_t stride) {
    if (!h || !vel) return NB200_ERR_BAD_ARG;
    CU(h, cudaSetDevice(h->device));
    return download_vec(h, vel, stride, 1);
}
int32_t nb200_get_forces(nb200_handle* h, float* force, int32_t stride) {
    if (!h || !force) return NB200_ERR_BAD_ARG;
    CU(h, cudaSetDevice(h->device));
    if (!h->have_forces) return fail(h, NB200_ERR_STATE, "forces not computed yet");
    return download_vec(h, force, stride, 2);
}

int32_t nb200_get_energies(nb200_handle* h, double* kinetic, double* potential) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->have_system || !h->have_forces) return fail(h, NB200_ERR_STATE, "no system / forces");
    CU(h, cudaSetDevice(h->device));
    if (!h->pe_valid) {  // the step loop skips the energy accumulation: redo the force pass with it on the same list
        CU(h, cudaMemsetAsync(h->force, 0, sizeof(float4) * (size_t)h->n, h->stream));
        int32_t rc = enqueue_force(h, true);
        if (rc) return rc;
    }
    h->kernel_launches += launch_energy(h->stream, h->vel[h->cur], h->force, h->n, h->vel_half ? 0.5f * h->last_dt : 0.f, h->energy_dev);
    CHECK_LAUNCH(h, "energy");
    double e[2];
    CU(h, cudaMemcpyAsync(e, h->energy_dev, sizeof(e), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (kinetic) *kinetic = e[0];
    if (potential) *potential = e[1];
    return NB200_OK;
}

// ---- stage-level ------------------------------------------------------------------------------------------------------
static int32_t keys_of(nb200_handle* h, const float* xyz, int32_t stride, int32_t n, uint32_t* keys, int curve) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!xyz || !keys) return fail(h, NB200_ERR_BAD_ARG, "NULL pointer");
    if (stride != 3 && stride != 4) return fail(h, NB200_ERR_BAD_ARG, "stride must be 3 or 4");
    int32_t rc = check_n(h, n);
    if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    h->have_system = false; h->have_forces = false; h->list_valid = false; h->mg_active = false;
    rc = upload_system(h, xyz, nullptr, stride, nullptr, nullptr, n, false);
    if (rc) return rc;
    {
        StageScope sc(h, NB200_STAGE_MORTON);
        sc.add(launch_morton(h->stream, h->pos[h->cur], n, h->box_min, h->box_max, h->keys[0], h->vals[0], curve));
        CHECK_LAUNCH(h, "morton");
    }
    CU(h, cudaMemcpyAsync(keys, h->keys[0], sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

int32_t nb200_morton30(nb200_handle* h, const float* xyz, int32_t stride, int32_t n, uint32_t* keys) {
    return keys_of(h, xyz, stride, n, keys, 0);
}

int32_t nb200_sort_keys(nb200_handle* h, const float* xyz, int32_t stride, int32_t n, uint32_t* keys) {
    return keys_of(h, xyz, stride, n, keys, h ? h->curve : 0);
}

int32_t nb200_set_curve(nb200_handle* h, int32_t curve) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (curve != 0 && curve != 1) return fail(h, NB200_ERR_BAD_ARG, "curve must be 0 (Morton) or 1 (Hilbert)");
    h->curve = curve;
    h->list_valid = false;
    h->have_forces = false;
    return NB200_OK;
}

int32_t nb200_set_list_mode(nb200_handle* h, int32_t mode) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (mode != NB200_LIST_HALF && mode != NB200_LIST_DIRECTED)
        return fail(h, NB200_ERR_BAD_ARG, "list mode must be NB200_LIST_HALF (1) or NB200_LIST_DIRECTED (0)");
    h->list_mode = mode;
    h->list_valid = false;
    h->have_forces = false;
    return NB200_OK;
}

int32_t nb200_set_fused_force(nb200_handle* h, int32_t enable) {
    if (!h) return NB200_ERR_BAD_ARG;
    h->fused_force = enable != 0;
    return NB200_OK;
}

int32_t nb200_set_list_reuse(nb200_handle* h, float skin, int32_t every) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (every < 1 || !(skin >= 0.f)) return fail(h, NB200_ERR_BAD_ARG, "list reuse needs every >= 1 and skin >= 0");
    if (every > 1 && !(skin > 0.f)) return fail(h, NB200_ERR_BAD_ARG, "reusing the list across steps needs a skin > 0");
    h->reuse_every = every;
    h->reuse_skin = every > 1 ? skin : 0.f;
    h->list_valid = false;   // the next force evaluation rebuilds the list with the new cutoff
    h->have_forces = false;
    return NB200_OK;
}

int32_t nb200_set_resort_interval(nb200_handle* h, int32_t every) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (every < 1) return fail(h, NB200_ERR_BAD_ARG, "re-sort interval must be >= 1");
    h->resort_interval = every;
    return NB200_OK;
}

int32_t nb200_sort_pairs(nb200_handle* h, uint32_t* keys, uint32_t* vals, int64_t n) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (n == 0) return NB200_OK;
    if (!keys || !vals) return fail(h, NB200_ERR_BAD_ARG, "NULL pointer");
    int32_t rc = check_n(h, n);
    if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    h->have_system = false; h->have_forces = false; h->list_valid = false;
    CU(h, cudaMemcpyAsync(h->keys[0], keys, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->vals[0], vals, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    int buf = 0;
    {
        StageScope sc(h, NB200_STAGE_SORT);
        h->hk_sort_clean = false;
        sc.add(launch_sort(h->stream, h->keys, h->vals, n, h->sort_hist, h->sort_status, h->sort_ticket, &buf));
        CHECK_LAUNCH(h, "sort");
    }
    CU(h, cudaMemcpyAsync(keys, h->keys[buf], sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(vals, h->vals[buf], sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

int32_t nb200_get_sorted_ids(nb200_handle* h, int32_t* ids) {
    if (!h || !ids) return NB200_ERR_BAD_ARG;
    if (!h->list_valid) return fail(h, NB200_ERR_STATE, "no search has run");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaMemcpyAsync(ids, h->id[h->cur], sizeof(int32_t) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

int32_t nb200_get_tree(nb200_handle* h, int32_t* n_leaves, int32_t* root, int32_t* node_child, float* node_box, float* leaf_box) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->list_valid) return fail(h, NB200_ERR_STATE, "no search has run");
    CU(h, cudaSetDevice(h->device));
    const int nL = h->n_leaves;
    if (n_leaves) *n_leaves = nL;
    if (root) *root = nL >= 2 ? 0 : ~0;
    std::vector<float4> lo(nL), hi(nL);
    if (leaf_box) {
        CU(h, cudaMemcpy(lo.data(), h->leaf_lo, sizeof(float4) * nL, cudaMemcpyDeviceToHost));
        CU(h, cudaMemcpy(hi.data(), h->leaf_hi, sizeof(float4) * nL, cudaMemcpyDeviceToHost));
        for (int i = 0; i < nL; ++i) {
            leaf_box[6 * i + 0] = lo[i].x; leaf_box[6 * i + 1] = lo[i].y; leaf_box[6 * i + 2] = lo[i].z;
            leaf_box[6 * i + 3] = hi[i].x; leaf_box[6 * i + 4] = hi[i].y; leaf_box[6 * i + 5] = hi[i].z;
        }
    }
    const int nI = nL - 1;
    if (nI > 0 && (node_child || node_box)) {
        std::vector<Node> nd(nI);
        CU(h, cudaMemcpy(nd.data(), h->nodes, sizeof(Node) * nI, cudaMemcpyDeviceToHost));
        CU(h, cudaMemcpy(lo.data(), h->node_lo, sizeof(float4) * nI, cudaMemcpyDeviceToHost));
        CU(h, cudaMemcpy(hi.data(), h->node_hi, sizeof(float4) * nI, cudaMemcpyDeviceToHost));
        for (int i = 0; i < nI; ++i) {
            if (node_child) {
                int32_t v[4];
                std::memcpy(&v[0], &nd[i].c[0].w, 4); std::memcpy(&v[1], &nd[i].c[1].w, 4);
                std::memcpy(&v[2], &nd[i].c[2].w, 4); std::memcpy(&v[3], &nd[i].c[3].w, 4);
                for (int k = 0; k < 4; ++k) node_child[4 * i + k] = v[k];
            }
            if (node_box) {
                node_box[6 * i + 0] = lo[i].x; node_box[6 * i + 1] = lo[i].y; node_box[6 * i + 2] = lo[i].z;
                node_box[6 * i + 3] = hi[i].x; node_box[6 * i + 4] = hi[i].y; node_box[6 * i + 5] = hi[i].z;
            }
        }
    }
    return NB200_OK;
}

int32_t nb200_get_neighbor_counts(nb200_handle* h, int32_t* counts) {
    if (!h || !counts) return NB200_ERR_BAD_ARG;
    if (!h->list_valid) return fail(h, NB200_ERR_STATE, "no neighbour list");
    CU(h, cudaSetDevice(h->device));
    int32_t rc = ensure_scratch(h, (int64_t)h->n * 4 + 64);
    if (rc) return rc;
    h->kernel_launches += launch_neighbor_counts(h->stream, h->sm_count, h->segs, h->entries, h->counters, h->seg_capacity,
                                                 h->id[h->cur], h->n, (int32_t*)h->scratch_dev, h->list_half);
    CHECK_LAUNCH(h, "neighbor_counts");
    CU(h, cudaMemcpyAsync(counts, h->scratch_dev, sizeof(int32_t) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

int32_t nb200_debug_traverse_profile(nb200_handle* h, int64_t* per_leaf4) {
    if (!h || !per_leaf4) return NB200_ERR_BAD_ARG;
    if (!h->list_valid) return fail(h, NB200_ERR_STATE, "no search has run");
    CU(h, cudaSetDevice(h->device));
    int32_t rc = ensure_scratch(h, (int64_t)h->n_leaves * 32 + 64);
    if (rc) return rc;
    h->kernel_launches += launch_traverse(h->stream, h->sm_count, h->nodes, h->frontier, h->leaf_lo, h->leaf_hi, h->leaf_sub, h->pos[h->cur], h->n,
                                          h->n_leaves, h->cutoff, h->entries, h->entry_capacity, h->segs, h->seg_capacity,
                                          h->counters, h->list_half, (long long*)h->scratch_dev);
    CHECK_LAUNCH(h, "traverse(debug)");
    CU(h, cudaMemcpyAsync(per_leaf4, h->scratch_dev, (size_t)h->n_leaves * 32, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

int32_t nb200_set_profiling(nb200_handle* h, int32_t enable) {
    if (!h) return NB200_ERR_BAD_ARG;
    CU(h, cudaSetDevice(h->device));
    StageTimer& t = h->timer;
    if (enable && !t.created) {
        for (int i = 0; i < StageTimer::MAX_EVENTS; ++i) CU(h, cudaEventCreate(&t.ev[i]));
        t.created = true;
    }
    CU(h, cudaStreamSynchronize(h->stream));
    t.n_ev = 0;
    for (int i = 0; i < NB200_STAGE_COUNT; ++i) { t.ms[i] = 0.0; t.launches[i] = 0; }
    t.enabled = enable != 0;
    t.only_stage = enable >= 2 ? enable - 2 : -1;
    return NB200_OK;
}

int32_t nb200_get_stage_times(nb200_handle* h, double* stage_ms, int64_t* stage_launches) {
    if (!h) return NB200_ERR_BAD_ARG;
    CU(h, cudaSetDevice(h->device));
    timer_collect(h);
    for (int i = 0; i < NB200_STAGE_COUNT; ++i) {
        if (stage_ms) stage_ms[i] = h->timer.ms[i];
        if (stage_launches) stage_launches[i] = h->timer.launches[i];
    }
    return NB200_OK;
}

// ---- multi-GPU: Morton-slab partition, one process per GPU -----------------------------
int32_t nb200_set_stream(nb200_handle* h, void* cuda_stream) {
    if (!h) return NB200_ERR_BAD_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->owns_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->owns_stream = false;
    return NB200_OK;
}

int32_t nb200_mg_set_owned(nb200_handle* h, const float* xyz, const float* vel, int32_t stride, const float* mass,
                           const float* charge, int32_t n_own) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!xyz) return fail(h, NB200_ERR_BAD_ARG, "xyz is NULL");
    if (stride != 3 && stride != 4) return fail(h, NB200_ERR_BAD_ARG, "stride must be 3 or 4");
    int32_t rc = check_n(h, n_own);
    if (rc) return rc;
    const int nLo = (n_own + LEAF - 1) / LEAF;
    const int gbase = nLo * LEAF;
    if ((int64_t)gbase + 2 * LEAF > h->n_max)
        return fail(h, NB200_ERR_BAD_ARG, "the handle's n_max %lld leaves no room for ghosts behind %d owned atoms", (long long)h->n_max, n_own);
    CU(h, cudaSetDevice(h->device));
    graph_invalidate(h);
    if (!h->mg_box) {
        CU(h, dalloc(&h->mg_box, 16));
        CU(h, dalloc(&h->mg_ghost_count, 2));
        CU(h, dalloc(&h->mg_err, 4));
        CU(h, dalloc(&h->mg_grid, 2 * 64 * 64));
        CU(h, dalloc(&h->mg_ghost_stat, 4));
        CU(h, cudaMemset(h->mg_ghost_stat, 0, 4 * sizeof(unsigned int)));
        CU(h, cudaHostAlloc((void**)&h->mg_ghost_count_h, 8, cudaHostAllocDefault));
        CU(h, cudaHostAlloc((void**)&h->mg_stat_h, 16, cudaHostAllocDefault));
        std::memset(h->mg_stat_h, 0, 16);
        // ghost pre-sort arrays and the second sort's scratch, sized for the largest ghost segment the handle can hold
        CU(h, dalloc(&h->mg_gpos, h->n_max));
        CU(h, dalloc(&h->mg_ggidx, h->n_max));
        for (int b = 0; b < 2; ++b) {
            CU(h, dalloc(&h->mg_gkeys[b], h->n_max));
            CU(h, dalloc(&h->mg_gvals[b], h->n_max));
        }
        CU(h, dalloc(&h->sort_hist2, 4 * 256));
        CU(h, dalloc(&h->sort_status2, 4 * h->sort_tiles_cap * 256));
        CU(h, dalloc(&h->sort_ticket2, 4));
        CU(h, dalloc(&h->frontier2, FRONTIER_WORDS));
        if (const char* cv = std::getenv("NB200_CARVEOUT")) {  // tuning: one shared-memory carve-out preference (percent) for every kernel of the slab step
            const int pct = std::atoi(cv);
            carveout_sort(pct); carveout_build(pct); carveout_peer(pct); carveout_atoms(pct); carveout_traverse(pct);
        }
        CU(h, dalloc(&h->mg_blist, h->n_max / LEAF + 2));
        CU(h, dalloc(&h->mg_bcount, 4));
        CU(h, dalloc(&h->mg_sendbuf, h->n_max));
        CU(h, dalloc(&h->mg_split, 66));
        {   // the ghost side's kernels are small and sit on the critical path of the ghost pass: highest priority
            int lo_p = 0, hi_p = 0;
            CU(h, cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
            CU(h, cudaStreamCreateWithPriority(&h->mg_stream2, cudaStreamNonBlocking, h->mg_ghost_prio_normal ? lo_p : hi_p));
        }
        CU(h, cudaEventCreateWithFlags(&h->mg_ev_int, cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&h->mg_ev_ghost, cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&h->mg_ev_owned, cudaEventDisableTiming));
    }
    // the published region (re)sized for this slab; peers must (re)connect afterwards
    mg_close_peers(h);
    if (h->mg_pub) cudaFree(h->mg_pub);
    h->mg_pub = nullptr;
    h->mg_pub_bytes = pub_layout(nullptr, h->n_max).bytes;
    CU(h, cudaMalloc(&h->mg_pub, (size_t)h->mg_pub_bytes));
    CU(h, cudaMemsetAsync(h->mg_pub, 0, (size_t)h->mg_pub_bytes, h->stream));
    CU(h, cudaMemsetAsync(h->mg_err, 0, 4 * sizeof(unsigned int), h->stream));
    {
        const PubLayout L = pub_layout(h->mg_pub, h->n_max);
        h->mg_flag = L.hdr;
        for (int b = 0; b < 2; ++b) { h->mg_pub_pos[b] = L.pos[b]; h->mg_pub_box[b] = L.box[b]; h->mg_pub_id[b] = L.id[b]; }
        h->mg_out_pos = L.out_pos; h->mg_out_vel = L.out_vel; h->mg_out_gid = L.out_gid; h->mg_out_dest = L.out_dest;
        h->mg_out_cap = L.out_cap;
        const unsigned int hdr[5] = {0u, (unsigned int)n_own, (unsigned int)n_own, 0u, (unsigned int)h->n_max};
        CU(h, cudaMemcpyAsync(L.hdr, hdr, sizeof(hdr), cudaMemcpyHostToDevice, h->stream));
    }
    h->mg_id_off = 0;
    h->mg_sorted_keys = nullptr;
    h->mg_migrated = false;
    h->mg_migrate_every = 0;
    h->mg_migration_pending = false;
    h->mg_steps_since_migration = 0;
    h->mg_n_pre = n_own;
    h->mg_parity = 0;
    h->mg_pub_step = 0;
    h->mg_world = 1;
    h->mg_rank = 0;
    h->mg_own_begin = 0;
    h->mg_max_peer_own = n_own;
    h->mg_ghost_cap = 0;
    h->mg_gfill = 0;
    h->mg_use_grid = true;
    h->mg_n_total = n_own;
    h->mg_n_own = n_own;
    h->mg_nLo = nLo;
    h->mg_gbase = gbase;
    h->mg_n_gslots = 0;
    h->hk2_clean = false;
    h->hk_sort_clean = false;
    rc = upload_system(h, xyz, vel, stride, mass, charge, n_own, true);  // packs into pos[0]/vel[0]/id[0] (id = hand-over index)
    if (rc) return rc;
    // the pad slots of the last owned leaf hold NaN in both buffers for good: they never pair and never widen a box
    if (gbase > n_own)
        for (int b = 0; b < 2; ++b) CU(h, cudaMemsetAsync(h->pos[b] + n_own, 0xff, sizeof(float4) * (size_t)(gbase - n_own), h->stream));
    // first publication (step 0): positions and ids in the hand-over order, leaf boxes, slab box
    CU(h, cudaMemcpyAsync(h->mg_pub_pos[0], h->pos[0], sizeof(float4) * (size_t)n_own, cudaMemcpyDeviceToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->mg_pub_id[0], h->id[0], sizeof(int32_t) * (size_t)n_own, cudaMemcpyDeviceToDevice, h->stream));
    h->kernel_launches += launch_mg_publish(h->stream, h->pos[0], n_own, h->mg_pub_box[0], h->mg_flag, ++h->mg_pub_step);
    h->kernel_launches += launch_slab_box(h->stream, h->pos[0], n_own, h->mg_box);  // slab box of parity 0
    h->kernel_launches += launch_slab_box_init(h->stream, h->mg_box + 8);           // parity 1: filled by the first integrate
    CHECK_LAUNCH(h, "mg_publish");
    CU(h, cudaStreamSynchronize(h->stream));
    h->mg_active = true;
    h->mg_pub_current = true;
    h->mg_keys_ready = false;
    h->have_system = false;
    h->have_forces = false;
    h->list_valid = false;
    h->vel_half = false;
    return NB200_OK;
}

// NCCL exchange: the owned positions (x, y, z, charge) in HAND-OVER order — the all-gather send buffer
int32_t nb200_mg_owned_pos_device(nb200_handle* h, void** ptr) {
    if (!h || !ptr) return NB200_ERR_BAD_ARG;
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    CU(h, cudaSetDevice(h->device));
    if (h->mg_migrated) return fail(h, NB200_ERR_STATE, "the all-gather send buffer needs the hand-over order: no migration with the NCCL exchange");
    h->kernel_launches += launch_unsort4(h->stream, h->pos[h->cur], h->id[h->cur], h->mg_n_own, h->mg_sendbuf, h->mg_id_off);
    CHECK_LAUNCH(h, "unsort4");
    *ptr = h->mg_sendbuf;
    return NB200_OK;
}

int32_t nb200_mg_publication(nb200_handle* h, void** device_base, int64_t* bytes, void* ipc_handle64) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    CU(h, cudaSetDevice(h->device));
    if (device_base) *device_base = h->mg_pub;
    if (bytes) *bytes = h->mg_pub_bytes;
    if (ipc_handle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        cudaIpcMemHandle_t hd;
        CU(h, cudaIpcGetMemHandle(&hd, h->mg_pub));
        std::memcpy(ipc_handle64, &hd, 64);
    }
    return NB200_OK;
}

int32_t nb200_mg_connect(nb200_handle* h, int32_t world, int32_t rank, const int64_t* own_begin, const int32_t* n_own,
                         const void* const* direct_base, const void* ipc_handles) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    if (world < 1 || world > 64 || rank < 0 || rank >= world || !own_begin || !n_own)
        return fail(h, NB200_ERR_BAD_ARG, "bad world / rank / slab tables (world <= 64)");
    if (n_own[rank] != h->mg_n_own) return fail(h, NB200_ERR_BAD_ARG, "n_own[rank] = %d but this handle owns %d atoms", n_own[rank], h->mg_n_own);
    CU(h, cudaSetDevice(h->device));
    mg_close_peers(h);
    std::vector<MgPeer> peers(world);
    int max_own = 0, max_out = 0, max_cap = 0;
    int64_t total = 0;
    for (int p = 0; p < world; ++p) total += n_own[p];
    if (total >= (1ll << 31)) return fail(h, NB200_ERR_BAD_ARG, "more than 2^31 atoms in all slabs");
    for (int p = 0; p < world; ++p) {
        void* base = nullptr;
        if (p == rank) {
            base = h->mg_pub;
        } else if (direct_base && direct_base[p]) {
            base = const_cast<void*>(direct_base[p]);  // same process (or otherwise already mapped)
        } else if (ipc_handles) {
            cudaIpcMemHandle_t hd;
            std::memcpy(&hd, (const char*)ipc_handles + 64 * (size_t)p, 64);
            CU(h, cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
            h->mg_ipc_opened[p] = base;
        } else {
            return fail(h, NB200_ERR_BAD_ARG, "no pointer and no IPC handle for peer %d", p);
        }
        unsigned int pcap = (unsigned int)h->n_max;
        if (p != rank) CU(h, cudaMemcpy(&pcap, (const unsigned int*)base + HDR_CAP, sizeof(pcap), cudaMemcpyDefault));  // the peer's capacity
        const PubLayout L = pub_layout(base, pcap);
        peers[p].pos[0] = L.pos[0]; peers[p].pos[1] = L.pos[1];
        peers[p].box[0] = L.box[0]; peers[p].box[1] = L.box[1];
        peers[p].id[0] = L.id[0]; peers[p].id[1] = L.id[1];
        peers[p].flag = L.hdr;
        peers[p].out_pos = L.out_pos; peers[p].out_vel = L.out_vel; peers[p].out_gid = L.out_gid; peers[p].out_dest = L.out_dest;
        peers[p].cap = (int)pcap;
        peers[p].out_cap = (int)L.out_cap;
        if (p != rank && (int)L.out_cap > max_out) max_out = (int)L.out_cap;
        if (p != rank && (int)pcap > max_cap) max_cap = (int)pcap;
        if (p != rank && n_own[p] > max_own) max_own = n_own[p];
    }
    if (h->mg_peers_dev) cudaFree(h->mg_peers_dev);
    h->mg_peers_dev = nullptr;
    CU(h, dalloc(&h->mg_peers_dev, world));
    CU(h, cudaMemcpy(h->mg_peers_dev, peers.data(), sizeof(MgPeer) * (size_t)world, cudaMemcpyHostToDevice));
    h->mg_world = world;
    h->mg_rank = rank;
    // the atom ids become GLOBAL (gathered) indices: own_begin + hand-over index — what the peers pull with the positions and
    // what follows an atom when it migrates
    const int delta = (int)own_begin[rank] - h->mg_id_off;
    if (delta != 0) {
        h->kernel_launches += launch_add_offset(h->stream, h->id[h->cur], h->mg_n_own, delta);
        CU(h, cudaMemcpyAsync(h->mg_pub_id[h->mg_parity], h->id[h->cur], sizeof(int32_t) * (size_t)h->mg_n_own, cudaMemcpyDeviceToDevice, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        h->mg_id_off = (int)own_begin[rank];
    }
    h->mg_own_begin = own_begin[rank];
    h->mg_max_peer_own = max_cap > 0 ? max_cap : max_own;  // the pull's grid covers the peers' CAPACITY: their atom counts change
    h->mg_max_peer_out = max_out;
    h->mg_n_total = total;
    h->mg_connected = true;
    return NB200_OK;
}

// Kick-drift(+reflect) of the owned atoms IN PLACE in their resident curve order; the same kernel publishes the step
// (positions, hand-over ids, leaf boxes, slab box), pre-fills the ghost slots and releases the flag (atoms.cu).
int32_t nb200_mg_integrate(nb200_handle* h, float dt) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->mg_active || !h->have_forces) return fail(h, NB200_ERR_STATE, "multi-GPU state needs nb200_mg_set_owned + nb200_mg_search_force first");
    CU(h, cudaSetDevice(h->device));
    const float kick_dt = h->vel_half ? 0.5f * (h->last_dt + dt) : 0.5f * dt;
    const int np = h->mg_parity ^ 1;
    mg_trace(h, 0, h->stream);
    {
        StageScope sc(h, NB200_STAGE_INTEGRATE);
        MgPublish pub = {};
        pub.pub_pos = h->mg_pub_pos[np]; pub.pub_id = h->mg_pub_id[np]; pub.id_in = h->id[h->cur]; pub.pub_box = h->mg_pub_box[np];
        pub.n_pub = h->mg_flag + HDR_NPUB + np;
        pub.slab_box6 = h->mg_box + 8 * np; pub.slab_box6_next = h->mg_box + 8 * (np ^ 1);
        pub.g_pos = h->mg_gpos; pub.g_keys = h->mg_gkeys[0]; pub.g_vals = h->mg_gvals[0];
        pub.g_fill = h->mg_world > 1 ? (int)h->mg_ghost_cap : 0;
        pub.flag = h->mg_flag; pub.done = h->mg_err + 2;
        if (h->mg_migrate_every > 0 && h->mg_world > 1 && h->mg_sorted_keys) {
            pub.prev_keys = h->mg_sorted_keys; pub.split = h->mg_split; pub.world = h->mg_world; pub.rank = h->mg_rank;
        }
        ++h->mg_pub_step;
        // MIGRATION step (every mg_migrate_every-th): the flag is released only after the leavers were classified and written to the outbox
        const bool migrate = h->mg_migrate_every > 0 && h->mg_world > 1 && h->mg_ghost_cap > 0 &&
                             ++h->mg_steps_since_migration >= h->mg_migrate_every;
        if (migrate) pub.flag = nullptr;
        sc.add(launch_integrate(h->stream, h->pos[h->cur], h->vel[h->cur], h->force, h->mg_n_own, kick_dt, dt, h->box_min, h->box_max,
                                h->keys[0], h->vals[0], h->curve, nullptr, &pub));
        CHECK_LAUNCH(h, "integrate(owned)");
        if (migrate) {
            sc.add(launch_mg_classify(h->stream, h->pos[h->cur], h->vel[h->cur], h->id[h->cur], h->keys[0], h->mg_n_own, h->box_min, h->box_max,
                                      h->curve, h->mg_split, h->mg_world, h->mg_rank, h->mg_out_pos, h->mg_out_vel, h->mg_out_gid, h->mg_out_dest,
                                      h->mg_flag + HDR_OUT, (int)h->mg_out_cap, h->mg_gpos, h->mg_ggidx, h->mg_gkeys[0], h->mg_gvals[0],
                                      h->mg_ghost_count, (unsigned int)h->mg_ghost_cap));
            sc.add(launch_mg_release_flag(h->stream, h->mg_flag, h->mg_pub_step));
            CHECK_LAUNCH(h, "classify");
            h->mg_migration_pending = true;
            h->mg_steps_since_migration = 0;
        }
        h->mg_parity = np;
        h->mg_gfill = pub.g_fill;
        h->mg_keys_ready = true;
        h->mg_pub_current = true;
        CU(h, cudaEventRecord(h->mg_ev_int, h->stream));
        mg_trace(h, 1, h->stream);
    }
    h->vel_half = true;
    h->last_dt = dt;
    h->have_forces = false;
    h->list_valid = false;
    h->steps_done++;
    return NB200_OK;
}

}  // extern "C"
namespace {
// (re)publish the owned atoms where they are now: the publishing integrate kernel with a zero kick and a zero drift
int32_t mg_publish_now(nb200_handle* h) {
    const int np = h->mg_parity ^ 1;
    MgPublish pub = {};
    pub.pub_pos = h->mg_pub_pos[np]; pub.pub_id = h->mg_pub_id[np]; pub.id_in = h->id[h->cur]; pub.pub_box = h->mg_pub_box[np];
    pub.n_pub = h->mg_flag + HDR_NPUB + np;
    pub.slab_box6 = h->mg_box + 8 * np; pub.slab_box6_next = h->mg_box + 8 * (np ^ 1);
    pub.g_pos = h->mg_gpos; pub.g_keys = h->mg_gkeys[0]; pub.g_vals = h->mg_gvals[0];
    pub.g_fill = h->mg_world > 1 ? (int)h->mg_ghost_cap : 0;
    pub.flag = h->mg_flag; pub.done = h->mg_err + 2;
    if (h->mg_migrate_every > 0 && h->mg_world > 1 && h->mg_sorted_keys) {
        pub.prev_keys = h->mg_sorted_keys; pub.split = h->mg_split; pub.world = h->mg_world; pub.rank = h->mg_rank;
    }
    ++h->mg_pub_step;
    h->kernel_launches += launch_integrate(h->stream, h->pos[h->cur], h->vel[h->cur], h->force, h->mg_n_own, 0.f, 0.f, h->box_min, h->box_max,
                                           h->keys[0], h->vals[0], h->curve, nullptr, &pub);
    CHECK_LAUNCH(h, "publish(owned)");
    h->mg_parity = np;
    h->mg_gfill = pub.g_fill;
    h->mg_keys_ready = true;
    h->mg_pub_current = true;
    CU(h, cudaEventRecord(h->mg_ev_int, h->stream));
    return NB200_OK;
}
}  // namespace
extern "C" {

// Ghosts: from the peers' published memory (all_pos_device == NULL; NVLink loads in mg_pull_kernel, no collective),
// or from an all-gathered float4[n_all] array in hand-over order (NCCL path; this rank's atoms sit at
// [own_begin, own_begin + n_own)).  Then: two sorts, two trees, traversal, forces on the owned atoms.
// Synchronous: the ghost count is read back, so the ghost segment is exact and the neighbour buffer can regrow.
int32_t nb200_mg_search_force(nb200_handle* h, const void* all_pos_device, int64_t n_all, int64_t own_begin, int64_t* n_ghost,
                              int64_t* n_entries) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    const bool mig = h->mg_migration_pending && !all_pos_device;
    if (mig) {
        CU(h, cudaSetDevice(h->device));
        int32_t rcm = mg_take_immigrants(h);
        if (rcm) return rcm;
    }
    const int n_own = h->mg_n_own;
    if (all_pos_device) {
        if (n_all < n_own || own_begin < 0 || own_begin + n_own > n_all) return fail(h, NB200_ERR_BAD_ARG, "bad gathered array / own range");
        if (n_all >= (1ll << 31)) return fail(h, NB200_ERR_BAD_ARG, "gathered array too large for int32 handles");
        h->mg_own_begin = own_begin;
    } else if (h->mg_world > 1 && !h->mg_connected) {
        return fail(h, NB200_ERR_STATE, "peer exchange needs nb200_mg_connect first");
    }
    CU(h, cudaSetDevice(h->device));
    if (!mig && !h->mg_pub_current && !all_pos_device) {  // atoms moved or changed owner since the last publication: every rank republishes
        int32_t rcp = mg_publish_now(h);
        if (rcp) return rcp;
    }
    const float cutoff = h->ff.cutoff;
    const int64_t galloc = mg_ghost_alloc(h);
    {
        StageScope sc(h, NB200_STAGE_MORTON);
        int* slab_box = h->mg_box + 8 * h->mg_parity;  // already filled by the publication; recomputing is idempotent
        if (!mig) sc.add(launch_slab_box(h->stream, h->pos[h->cur], n_own, slab_box));
        if (!all_pos_device) {
            // The ghost filter of this search must be the one the asynchronous steps will use (slab box alone, or box +
            // occupancy grid): this search sizes their ghost slots from its own count.
            int box6[6];
            CU(h, cudaMemcpyAsync(box6, slab_box, sizeof(box6), cudaMemcpyDeviceToHost, h->stream));
            CU(h, cudaStreamSynchronize(h->stream));
            h->mg_use_grid = mg_slab_is_ragged(h, box6, n_own);
        }
        if (all_pos_device) {
            sc.add(launch_mg_grid(h->stream, h->pos[h->cur], n_own, h->box_min, h->box_max, cutoff, h->mg_grid));
            sc.add(launch_ghost_select(h->stream, (const float4*)all_pos_device, n_all, own_begin, n_own, slab_box, cutoff, h->mg_gpos,
                                       h->mg_ggidx, h->mg_ghost_count, galloc, h->mg_grid + 64 * 64, h->box_min, h->box_max, h->curve,
                                       h->mg_gkeys[0], h->mg_gvals[0]));
            CHECK_LAUNCH(h, "ghost_select");
        } else {
            // a peer that never publishes is reported after ~5 s instead of hanging the GPU
            sc.add(launch_mg_pull(h->stream, h->mg_peers_dev, h->mg_world, h->mg_rank, h->mg_max_peer_own, h->mg_parity,
                                  h->pos[h->cur], mig ? h->mg_n_pre : n_own, slab_box, cutoff, h->mg_gpos, h->mg_ggidx, h->mg_ghost_count, galloc, h->mg_err,
                                  10000000000ll, nullptr, h->box_min, h->box_max, h->curve, h->mg_gkeys[0], h->mg_gvals[0],
                                  h->mg_use_grid ? h->mg_grid : nullptr, h->mg_err + 3, mig ? h->mg_split : nullptr, mig, true));
            CHECK_LAUNCH(h, "mg_pull");
        }
    }
    if (h->mg_world == 1 && !all_pos_device) CU(h, cudaMemsetAsync(h->mg_ghost_count, 0, sizeof(unsigned int), h->stream));
    CU(h, cudaMemcpyAsync(h->mg_ghost_count_h, h->mg_ghost_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(h->mg_ghost_count_h + 1, h->mg_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (all_pos_device) h->mg_use_grid = true;
    if (h->mg_ghost_count_h[1]) {
        const unsigned peer = h->mg_ghost_count_h[1] - 1u;
        cudaMemsetAsync(h->mg_err, 0, sizeof(unsigned int), h->stream);
        return fail(h, NB200_ERR_STATE, "peer %u did not publish step %u within the time limit", peer, h->mg_pub_step);
    }
    const int64_t ng = *h->mg_ghost_count_h;
    if (ng > galloc)
        return fail(h, NB200_ERR_BAD_ARG, "owned %d + ghosts %lld exceed the handle's n_max %lld", n_own, (long long)ng, (long long)h->n_max);
    h->mg_n_ghost = (int32_t)ng;
    {   // ghost slots of the asynchronous step: 50 % above what this slab needs now, within the handle's n_max
        int64_t cap = ng + ng / 2 + 8192;
        if (cap > galloc) cap = galloc;
        if (cap > h->mg_ghost_cap) h->mg_ghost_cap = cap;
    }
    CU(h, cudaMemcpyAsync(&h->counters->sticky_saved, &h->counters->overflow_sticky, sizeof(unsigned int), cudaMemcpyDeviceToDevice, h->stream));
    int32_t rc = mg_search(h, (int)ng, false, false, false, h->mg_keys_ready);
    if (rc) return rc;
    for (int attempt = 0;; ++attempt) {  // regrow-and-retry like search_sync (both trees stay valid, only the traversal reruns)
        rc = read_counters(h);
        if (rc) return rc;
        const int64_t need = (int64_t)h->counters_h->n_entries();
        const bool tight = need + need / 5 + 4096 > h->entry_capacity;
        if (!h->counters_h->overflow && !tight) break;
        if (attempt == 4) return fail(h, NB200_ERR_PAIR_OVERFLOW, "neighbour buffer still too small after regrowing");
        rc = ensure_entries(h, need + need / 4);
        if (rc) return rc;
        rc = mg_launch_traverse(h, false, false, 1, h->stream);
        if (!rc) rc = mg_launch_traverse(h, false, true, 2, h->stream);
        if (rc) return rc;
    }
    if (h->counters_h->overflow_sticky)
        CU(h, cudaMemcpyAsync(&h->counters->overflow_sticky, &h->counters->sticky_saved, sizeof(unsigned int), cudaMemcpyDeviceToDevice, h->stream));
    h->list_valid = true;
    h->mg_keys_ready = false;
    rc = enqueue_force(h, false);  // energies are accumulated on demand (nb200_mg_get_energies)
    if (rc) return rc;
    h->have_forces = true;
    if (n_ghost) *n_ghost = ng;
    if (n_entries) *n_entries = (int64_t)h->counters_h->n_valid;
    return NB200_OK;
}

// The same search without a host round trip: the ghost segment is sized for the ghost capacity, its unused slots hold
// inert NaN placeholders, the neighbour buffer cannot regrow, and the ghost side (waiting for the peers, pull, sort,
// ghost tree) runs on a second stream under the owned sort and tree build.  Problems (more ghosts than slots, list
// overflow, a peer that never published) are sticky on the device and reported by nb200_mg_sync.
}  // extern "C"
namespace {

// Host side of an asynchronous slab step, outside any graph capture: bounds how far the host runs ahead and lets the
// ghost capacity follow the ghost count.  The host may run at most 16 submissions ahead of the GPU (it waits for the one
// enqueued 16 calls ago — the GPU stays busy), so the ghost statistics the steps copy to pinned memory are at most that
// old when read here: the capacity follows the count with 50 % headroom, without a round trip.
int32_t mg_async_prologue(nb200_handle* h) {
    if (!h->mg_ev_created) {
        for (int k = 0; k < 16; ++k) CU(h, cudaEventCreateWithFlags(&h->mg_step_ev[k], cudaEventDisableTiming));
        h->mg_ev_created = true;
        h->mg_async_subs = 0;
    }
    if (h->mg_async_subs >= h->mg_ahead) CU(h, cudaEventSynchronize(h->mg_step_ev[(h->mg_async_subs - h->mg_ahead) % 16]));
    if (h->mg_world > 1) {
        const int64_t galloc = mg_ghost_alloc(h);
        const int64_t seen = h->mg_stat_h[0];  // largest count of the steps that have completed
        if (seen + seen / 4 > h->mg_ghost_cap) {
            int64_t want = seen + seen / 2 + 8192;
            if (want > galloc) want = galloc;
            if (want > h->mg_ghost_cap) h->mg_ghost_cap = want;
        }
    }
    return NB200_OK;
}

int32_t mg_async_epilogue(nb200_handle* h) {
    CU(h, cudaEventRecord(h->mg_step_ev[h->mg_async_subs % 16], h->stream));
    ++h->mg_async_subs;
    return NB200_OK;
}

// Device side of the asynchronous search: pure stream work (capturable into a CUDA graph).
int32_t mg_async_enqueue(nb200_handle* h, bool copy_stats) {
    const bool mig = h->mg_migration_pending;  // the integrate of this step classified the atoms: take the immigrants in first
    if (mig) {
        const int32_t rcm = mg_take_immigrants(h);
        if (rcm) return rcm;
    }
    const int n_own = h->mg_n_own;
    const int cap = h->mg_world > 1 ? (int)h->mg_ghost_cap : 0;
    const float cutoff = h->ff.cutoff;
    const bool keys_ready = h->mg_keys_ready;  // the publishing integrate kernel wrote the owned keys (and pre-filled the ghost slots)
    const float4* own_pos_now = h->pos[h->cur];
    const int n_pre_pull = mig ? h->mg_n_pre : n_own;
    auto enqueue_pull = [&]() -> int32_t {
        cudaStream_t sB = h->mg_stream2;
        StageScope sc(h, NB200_STAGE_MORTON);
        if (keys_ready) CU(h, cudaStreamWaitEvent(sB, h->mg_ev_int, 0));
        else {  // (first step after a synchronous search: nothing was published by an integrate yet)
            CU(h, cudaEventRecord(h->mg_ev_int, h->stream));
            CU(h, cudaStreamWaitEvent(sB, h->mg_ev_int, 0));
            h->mg_gfill = 0;
        }
        if (h->mg_gfill < cap)  // the capacity grew since the integrate pre-filled the ghost slots
            sc.add(launch_mg_ghost_fill(sB, h->mg_gpos, h->mg_gkeys[0], h->mg_gvals[0], (int)h->mg_gfill, cap, h->box_min, h->box_max, h->curve));
        sc.add(launch_mg_pull(sB, h->mg_peers_dev, h->mg_world, h->mg_rank, h->mg_max_peer_own, h->mg_parity,
                              own_pos_now, n_pre_pull, h->mg_box + 8 * h->mg_parity, cutoff, h->mg_gpos, h->mg_ggidx, h->mg_ghost_count, cap,
                              h->mg_err, 10000000000ll, h->mg_ghost_stat, h->box_min, h->box_max, h->curve, h->mg_gkeys[0], h->mg_gvals[0],
                              h->mg_use_grid ? h->mg_grid : nullptr, h->mg_err + 3, mig ? h->mg_split : nullptr, mig, true));
        CHECK_LAUNCH(h, "mg_pull");
        return NB200_OK;
    };
    // The pull is enqueued together with the rest of the ghost side once the owned pass has been launched (mg_search):
    // right behind the integrate its ~2000 high-priority blocks ran beside the owned sort, whose decoupled look-back chains
    // stall when tiles are descheduled (2 GPUs: sort 0.062 -> 0.042 ms, reorder 0.027 -> 0.020 ms, step 0.578 -> 0.558 ms).
    // The whole ghost side still fits under the owned pass.  (NB200_MG_PULL_EARLY restores the early pull for comparison.)
    std::function<int32_t()> pull_fn = enqueue_pull;
    if (cap > 0 && !h->mg_pull_late) {
        const int32_t rcp = enqueue_pull();
        if (rcp) return rcp;
    }
    const bool fused = h->fused_force && !(h->ff.eps == 0.f && h->ff.kcoul == 0.f);
    int32_t rc = mg_search(h, cap, cap > 0, fused, true, keys_ready, (cap > 0 && h->mg_pull_late) ? &pull_fn : nullptr);  // coarse owned sort: the atoms are resident in curve order
    if (rc) return rc;
    h->mg_keys_ready = false;
    if (!fused) {
        rc = enqueue_force(h, false);
        if (rc) return rc;
    }
    if (copy_stats)  // the statistics the capacity follows
        CU(h, cudaMemcpyAsync(h->mg_stat_h, h->mg_ghost_stat, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    ++h->mg_async_steps;
    h->have_forces = true;
    h->list_valid = true;
    h->async_overflow_possible = true;
    return NB200_OK;
}

int32_t mg_async_checks(nb200_handle* h) {
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    if (h->mg_world > 1 && !h->mg_connected) return fail(h, NB200_ERR_STATE, "peer exchange needs nb200_mg_connect first");
    if (h->mg_ghost_cap <= 0 && h->mg_world > 1)
        return fail(h, NB200_ERR_STATE, "run nb200_mg_search_force once first: it sizes the ghost region and the neighbour buffer");
    return NB200_OK;
}

// ---- the slab step loop as a CUDA graph: two steps (the buffers and the publication parity alternate), both streams ----
struct MgGraphKey {
    int32_t n_own, cap, cur, parity, fused, list_mode, curve, world, use_grid;
    float dt;
    ForceField ff;
    float box[6];
    const void *entries, *segs, *pub, *peers;
    int64_t entry_capacity, seg_capacity;
};

MgGraphKey mg_graph_key(const nb200_handle* h, float dt) {
    MgGraphKey k;
    std::memset(&k, 0, sizeof(k));
    k.n_own = h->mg_n_own; k.cap = (int32_t)h->mg_ghost_cap; k.cur = h->cur; k.parity = h->mg_parity;
    k.fused = (h->fused_force && !(h->ff.eps == 0.f && h->ff.kcoul == 0.f)) ? 1 : 0;
    k.list_mode = h->list_mode; k.curve = h->curve; k.world = h->mg_world; k.use_grid = h->mg_use_grid ? 1 : 0;
    k.dt = dt; k.ff = h->ff;
    for (int d = 0; d < 3; ++d) { k.box[d] = h->box_min[d]; k.box[3 + d] = h->box_max[d]; }
    k.entries = h->entries; k.segs = h->segs; k.pub = h->mg_pub; k.peers = h->mg_peers_dev;
    k.entry_capacity = h->entry_capacity; k.seg_capacity = h->seg_capacity;
    return k;
}

int32_t mg_graph_capture(nb200_handle* h, float dt) {
    graph_invalidate(h);
    const int64_t launches0 = h->kernel_launches, steps0 = h->steps_done, asteps0 = h->mg_async_steps;
    const unsigned int pub0 = h->mg_pub_step;
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return NB200_ERR_CUDA; }
    int32_t rc = NB200_OK;
    for (int k = 0; k < 2 && !rc; ++k) {
        rc = nb200_mg_integrate(h, dt);
        if (!rc) rc = mg_async_enqueue(h, k == 0);
    }
    const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    h->graph_launches = h->kernel_launches - launches0;
    h->kernel_launches = launches0;
    h->steps_done = steps0;
    h->mg_async_steps = asteps0;
    h->mg_pub_step = pub0;
    if (rc || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return rc ? rc : NB200_ERR_CUDA;
    }
    cudaGraphExec_t ex = nullptr;
    const cudaError_t e2 = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) { cudaGetLastError(); return NB200_ERR_CUDA; }
    h->graph_exec = ex;
    return NB200_OK;
}

}  // namespace
extern "C" {

// The same search without a host round trip: the ghost segment is sized for the ghost capacity, its unused slots hold
// inert NaN placeholders, the neighbour buffer cannot regrow, and the ghost side (waiting for the peers, pull, sort,
// ghost tree, ghost pass) runs on a second stream beside the owned pass.  Problems (more ghosts than slots, list
// overflow, a peer that never published) are sticky on the device and reported by nb200_mg_sync.
int32_t nb200_mg_search_force_async(nb200_handle* h) {
    if (!h) return NB200_ERR_BAD_ARG;
    int32_t rc = mg_async_checks(h);
    if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    rc = mg_async_prologue(h);
    if (rc) return rc;
    rc = mg_async_enqueue(h, h->mg_async_steps % 8 == 0);
    if (rc) return rc;
    return mg_async_epilogue(h);
}

// nsteps x (nb200_mg_integrate + nb200_mg_search_force_async).  In steady state two consecutive steps — both streams, the
// waits on the peers' flags included — are captured once as a CUDA graph and replayed: the ~25 launches of a slab step
// otherwise leave the GPU idle for 1-2 us each.  (Every rank must call it with the same nsteps: the ranks move in lockstep.)
int32_t nb200_mg_step_async(nb200_handle* h, int32_t nsteps, float dt) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (nsteps < 0) return fail(h, NB200_ERR_BAD_ARG, "nsteps must be >= 0");
    int32_t rc = mg_async_checks(h);
    if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    int32_t s = 0;
    while (s < nsteps) {
        // (world > 1: replay measured SLOWER than stream launches, 0.74 vs 0.62 ms per step on 2 B200 — captured kernel nodes do
        //  not carry the ghost stream's priority, so the ghost side starves behind the 15 k blocks of the owned pass instead
        //  of running beside it; NB200_MG_GRAPH=1 turns it on for experiments)
        const bool steady = h->use_graph && (h->mg_world == 1 || h->mg_graph_multi) && !h->timer.enabled && !h->mg_trace &&
                            h->mg_async_steps >= 2 && h->have_forces && h->vel_half && h->last_dt == dt && h->hk_sort_clean &&
                            (h->mg_world == 1 || h->hk2_clean);
        if (nsteps - s >= 4 && steady) {
            rc = mg_async_prologue(h);  // (may raise the ghost capacity: part of the key)
            if (rc) return rc;
            const MgGraphKey key = mg_graph_key(h, dt);
            static_assert(sizeof(MgGraphKey) <= sizeof(h->graph_key), "graph key storage");
            if (!h->graph_exec || !h->graph_is_mg || std::memcmp(&key, h->graph_key, sizeof(key)) != 0) {
                if (mg_graph_capture(h, dt) == NB200_OK) { std::memcpy(h->graph_key, &key, sizeof(key)); h->graph_is_mg = true; }
                else h->use_graph = false;
            }
            if (h->graph_exec) {
                CU(h, cudaGraphLaunch((cudaGraphExec_t)h->graph_exec, h->stream));
                h->kernel_launches += h->graph_launches;
                h->steps_done += 2;
                h->mg_async_steps += 2;
                h->mg_pub_step += 2;
                h->async_overflow_possible = true;
                h->pe_valid = false;
                rc = mg_async_epilogue(h);
                if (rc) return rc;
                s += 2;
                continue;
            }
        }
        rc = nb200_mg_integrate(h, dt);
        if (!rc) rc = nb200_mg_search_force_async(h);
        if (rc) return rc;
        ++s;
    }
    return NB200_OK;
}

// The slab step with HOST buffers, leapfrog order like nb200_leapfrog_host_async (positions-only exchange): the owned
// positions x(t) come from the caller (hand-over order, pinned memory for true asynchrony), are published to the peers,
// the halo is pulled, list and forces are rebuilt at x(t), the owned atoms are kicked and drifted, and x(t + dt) goes back
// into the same buffer.  Velocities stay resident.  Fully asynchronous; finish with nb200_mg_sync.
int32_t nb200_mg_leapfrog_host_async(nb200_handle* h, float* xyz, int32_t stride, float dt) {
    if (!h) return NB200_ERR_BAD_ARG;
    int32_t rc = mg_async_checks(h);
    if (rc) return rc;
    if (!xyz) return fail(h, NB200_ERR_BAD_ARG, "xyz is NULL");
    if (stride != 3 && stride != 4) return fail(h, NB200_ERR_BAD_ARG, "stride must be 3 or 4");
    CU(h, cudaSetDevice(h->device));
    const int n = h->mg_n_own;
    float* sx = h->stage_dev;
    const size_t bytes = sizeof(float) * (size_t)n * stride;
    rc = mg_async_prologue(h);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(sx, xyz, bytes, cudaMemcpyHostToDevice, h->stream));
    if (h->mg_migrate_every > 0 || h->mg_migrated)
        return fail(h, NB200_ERR_STATE, "the host-buffer slab step keeps the hand-over order of the caller's array: not with migration");
    h->kernel_launches += launch_refresh(h->stream, sx, nullptr, stride, h->id[h->cur], n, h->pos[h->cur], h->vel[h->cur], nullptr, nullptr, 0,
                                         nullptr, nullptr, h->mg_id_off);
    CHECK_LAUNCH(h, "refresh(owned)");
    rc = mg_publish_now(h);  // x(t) to the peers
    if (rc) return rc;
    rc = mg_async_enqueue(h, h->mg_async_steps % 8 == 0);
    if (rc) return rc;
    {
        StageScope sc(h, NB200_STAGE_INTEGRATE);
        sc.add(launch_integrate(h->stream, h->pos[h->cur], h->vel[h->cur], h->force, n, h->vel_half ? dt : 0.5f * dt, dt, h->box_min,
                                h->box_max, h->keys[0], h->vals[0], h->curve));
        CHECK_LAUNCH(h, "integrate(owned)");
    }
    h->kernel_launches += launch_unpack(h->stream, h->pos[h->cur], h->id[h->cur], n, stride, sx, 0, nullptr, 0.f, h->mg_id_off);
    CHECK_LAUNCH(h, "unpack(owned)");
    CU(h, cudaMemcpyAsync(xyz, sx, bytes, cudaMemcpyDeviceToHost, h->stream));
    h->vel_half = true;
    h->last_dt = dt;
    h->have_forces = false;
    h->list_valid = false;
    h->mg_pub_current = false;  // the atoms moved after the publication
    h->mg_keys_ready = false;
    h->steps_done++;
    return mg_async_epilogue(h);
}

// Waits for the asynchronous steps and reports what they could not: ghost capacity exceeded, neighbour buffer
// overflow, peer timeout.  n_ghost: ghosts of the last step, n_entries: its list entries (NaN placeholders have none).
int32_t nb200_mg_sync(nb200_handle* h, int64_t* n_ghost, int64_t* n_entries) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    CU(h, cudaSetDevice(h->device));
    unsigned int stat[4] = {0, 0, 0, 0}, err = 0;
    CU(h, cudaMemcpyAsync(stat, h->mg_ghost_stat, sizeof(stat), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(&err, h->mg_err, sizeof(err), cudaMemcpyDeviceToHost, h->stream));
    int32_t rc = read_counters(h);  // synchronises the stream (which waited for the ghost stream of every step)
    if (rc) return rc;
    if (h->mg_trace && h->mg_trace_ev[0] && h->mg_rank == 0) {  // tuning aid (NB200_MG_TRACE): timelines of the last <= 24 steps
        const long long last = h->mg_trace_step, first = std::max<long long>(last - 23, h->mg_trace_printed + 1);
        for (long long st_ = first; st_ <= last; ++st_) {
            float t[6] = {0, 0, 0, 0, 0, 0}, gap = 0.f;
            cudaEvent_t* e = &h->mg_trace_ev[(st_ % 32) * 6];
            for (int k = 1; k < 6; ++k) cudaEventElapsedTime(&t[k], e[0], e[k]);
            if (st_ > first) cudaEventElapsedTime(&gap, h->mg_trace_ev[((st_ - 1) % 32) * 6], e[0]);
            cudaGetLastError();
            fprintf(stderr, "[nb200 rank 0] step %lld (prev step took %.3f): integrate end %.3f | ghost tree %.3f | owned tree %.3f | owned pass end %.3f | ghost pass end %.3f\n",
                    st_, gap, t[1], t[2], t[3], t[4], t[5]);
        }
        h->mg_trace_printed = last;
    }
    CU(h, cudaMemsetAsync(h->mg_ghost_stat, 0, 2 * sizeof(unsigned int), h->stream));
    if (n_ghost) *n_ghost = stat[2];
    if (n_entries) *n_entries = (int64_t)h->counters_h->n_valid;
    if (err) {
        CU(h, cudaMemsetAsync(h->mg_err, 0, sizeof(unsigned int), h->stream));
        return fail(h, NB200_ERR_STATE, "peer %u did not publish within the time limit", err - 1u);
    }
    // keep 50 % headroom over the largest ghost count seen
    const int64_t galloc = mg_ghost_alloc(h);
    int64_t want = (int64_t)stat[0] + (int64_t)stat[0] / 2 + 8192;
    if (want > galloc) want = galloc;
    if (want > h->mg_ghost_cap) h->mg_ghost_cap = want;
    if (stat[1]) {
        h->have_forces = false;
        return fail(h, NB200_ERR_PAIR_OVERFLOW, "a step needed %u ghost slots, more than the asynchronous step provided; capacity raised to %lld — "
                    "the steps since the last nb200_mg_sync are invalid", stat[0], (long long)h->mg_ghost_cap);
    }
    if (h->async_overflow_possible) {
        h->async_overflow_possible = false;
        if (h->counters_h->overflow_sticky) {
            const unsigned long long need = h->counters_h->n_entries();
            CU(h, cudaMemsetAsync(&h->counters->overflow_sticky, 0, sizeof(unsigned int), h->stream));
            h->have_forces = false;
            h->list_valid = false;
            ensure_entries(h, (int64_t)need * 2);
            return fail(h, NB200_ERR_PAIR_OVERFLOW, "neighbour buffer overflowed during the asynchronous steps (needed >= %llu slots); buffer regrown", need);
        }
    }
    return NB200_OK;
}

// Publishes the owned atoms where they are now if the rank's publication is stale (the atoms moved or changed owner since
// it was made: host-buffer steps, a migration).  nb200_mg_search_force does this by itself; a driver that steps several
// slabs from ONE host thread calls it for all of them first, because a search waits for every peer's publication.
int32_t nb200_mg_republish(nb200_handle* h) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    if (h->mg_pub_current || h->mg_migration_pending) return NB200_OK;
    CU(h, cudaSetDevice(h->device));
    return mg_publish_now(h);
}

// Atoms this rank owns now (changes when atoms migrate).
int32_t nb200_mg_owned_count(nb200_handle* h, int32_t* n_own) {
    if (!h || !n_own) return NB200_ERR_BAD_ARG;
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    *n_own = h->mg_n_own;
    return NB200_OK;
}

// Global id (gathered index: the rank's own_begin + index in its hand-over order at setup) of the atom in every ROW of
// nb200_mg_get_owned.  Without migration the rows are the hand-over order; once atoms have migrated they are the
// rank's current curve order.
int32_t nb200_mg_get_owned_ids(nb200_handle* h, int32_t* ids) {
    if (!h || !ids) return NB200_ERR_BAD_ARG;
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    CU(h, cudaSetDevice(h->device));
    const int n = h->mg_n_own;
    if (h->mg_migrated) {
        CU(h, cudaMemcpyAsync(ids, h->id[h->cur], sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    } else {
        for (int k = 0; k < n; ++k) ids[k] = (int32_t)h->mg_own_begin + k;  // hand-over rows
    }
    return NB200_OK;
}

// Splitters of the Morton key space (world + 1 keys, split[0] = 0, split[world] = 2^30: rank g owns keys in
// [split[g], split[g+1])) and the migration interval: every `every`-th nb200_mg_integrate hands the atoms that left their
// rank's key range over to the new owner (0: never).  Same arguments on every rank; peer exchange only.
int32_t nb200_mg_set_migration(nb200_handle* h, const uint32_t* split, int32_t n_split, int32_t every) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    if (every < 0) return fail(h, NB200_ERR_BAD_ARG, "migration interval must be >= 0");
    if (every > 0) {
        if (!split || n_split != h->mg_world + 1 || n_split > 65) return fail(h, NB200_ERR_BAD_ARG, "need world + 1 = %d splitters (call nb200_mg_connect first)", h->mg_world + 1);
        for (int g = 0; g < h->mg_world; ++g)
            if (split[g] > split[g + 1]) return fail(h, NB200_ERR_BAD_ARG, "splitters must be non-decreasing");
        CU(h, cudaSetDevice(h->device));
        CU(h, cudaMemcpy(h->mg_split, split, sizeof(uint32_t) * (size_t)n_split, cudaMemcpyHostToDevice));
    }
    h->mg_migrate_every = every;
    h->mg_steps_since_migration = 0;
    return NB200_OK;
}

// mode 0 positions, 1 velocities (synchronised), 2 forces — one row per owned atom (row order: nb200_mg_get_owned_ids)
int32_t nb200_mg_get_owned(nb200_handle* h, float* out, int32_t stride, int32_t mode) {
    if (!h || !out) return NB200_ERR_BAD_ARG;
    if (!h->mg_active) return fail(h, NB200_ERR_STATE, "nb200_mg_set_owned has not been called");
    if (stride != 3 && stride != 4) return fail(h, NB200_ERR_BAD_ARG, "stride must be 3 or 4");
    if (mode < 0 || mode > 2) return fail(h, NB200_ERR_BAD_ARG, "mode must be 0, 1 or 2");
    if (mode == 2 && !h->have_forces) return fail(h, NB200_ERR_STATE, "forces not computed yet");
    CU(h, cudaSetDevice(h->device));
    const int n = h->mg_n_own;
    const float4* src = mode == 0 ? h->pos[h->cur] : (mode == 1 ? h->vel[h->cur] : h->force);
    const bool pending = (mode == 1) && h->vel_half && h->have_forces;
    h->kernel_launches += launch_unpack(h->stream, src, h->mg_migrated ? nullptr : h->id[h->cur], n, stride, h->stage_dev, mode,
                                        pending ? h->force : nullptr, 0.5f * h->last_dt, h->mg_id_off);
    CHECK_LAUNCH(h, "unpack(owned)");
    CU(h, cudaMemcpyAsync(out, h->stage_dev, sizeof(float) * (size_t)n * stride, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

int32_t nb200_mg_get_energies(nb200_handle* h, double* kinetic, double* potential) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->mg_active || !h->have_forces) return fail(h, NB200_ERR_STATE, "no multi-GPU forces yet");
    CU(h, cudaSetDevice(h->device));
    if (!h->pe_valid) {  // the step loop skips the energy accumulation: redo the force pass with it on the same list
        CU(h, cudaMemsetAsync(h->force, 0, sizeof(float4) * (size_t)h->mg_n_own, h->stream));
        int32_t rc = enqueue_force(h, true);
        if (rc) return rc;
    }
    h->kernel_launches += launch_energy(h->stream, h->vel[h->cur], h->force, h->mg_n_own, h->vel_half ? 0.5f * h->last_dt : 0.f, h->energy_dev);
    CHECK_LAUNCH(h, "energy(owned)");
    double e[2];
    CU(h, cudaMemcpyAsync(e, h->energy_dev, sizeof(e), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (kinetic) *kinetic = e[0];
    if (potential) *potential = e[1];
    return NB200_OK;
}

// this rank's list entries as (gathered index of the row atom, gathered index of the partner, d): half list — every
// pair with at least one owned atom, once; directed list — the complete rows of the owned atoms
int32_t nb200_mg_get_entries(nb200_handle* h, int32_t* a, int32_t* b, float* d, int64_t capacity, int64_t* written) {
    if (!h) return NB200_ERR_BAD_ARG;
    if (!h->mg_active || !h->list_valid) return fail(h, NB200_ERR_STATE, "no multi-GPU neighbour list");
    CU(h, cudaSetDevice(h->device));
    int32_t rc = read_counters(h);
    if (rc) return rc;
    const int64_t ne = (int64_t)h->counters_h->n_valid;
    if (written) *written = ne;
    if (ne == 0) return NB200_OK;
    if (capacity < ne) return fail(h, NB200_ERR_CAPACITY, "buffers hold %lld, list has %lld directed entries", (long long)capacity, (long long)ne);
    if (!a || !b || !d) return fail(h, NB200_ERR_BAD_ARG, "output pointers are NULL");
    if (h->exp_capacity < ne) {
        cudaFree(h->exp_a); cudaFree(h->exp_b); cudaFree(h->exp_d);
        h->exp_a = h->exp_b = nullptr; h->exp_d = nullptr; h->exp_capacity = 0;
        CU(h, dalloc(&h->exp_a, ne));
        CU(h, dalloc(&h->exp_b, ne));
        CU(h, dalloc(&h->exp_d, ne));
        h->exp_capacity = ne;
    }
    // sorted slot -> gathered index: owned slots through the hand-over id, ghost slots through the ghost's pre-sort index
    h->kernel_launches += launch_compose(h->stream, h->id[h->cur], h->mg_n_own, h->mg_gbase, h->n, (int)h->mg_own_begin - h->mg_id_off, h->mg_ggidx,
                                         (int32_t*)h->vals[1]);
    h->kernel_launches += launch_export_directed(h->stream, h->sm_count, h->segs, h->entries, h->counters, h->seg_capacity,
                                                 h->pos[h->cur], (const int32_t*)h->vals[1], h->n, h->exp_a, h->exp_b, h->exp_d, ne);
    CHECK_LAUNCH(h, "export_directed");
    CU(h, cudaMemcpyAsync(a, h->exp_a, sizeof(int32_t) * (size_t)ne, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(b, h->exp_b, sizeof(int32_t) * (size_t)ne, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(d, h->exp_d, sizeof(float) * (size_t)ne, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return NB200_OK;
}

int32_t nb200_timer_start(nb200_handle* h) {
    if (!h) return NB200_ERR_BAD_ARG;
    CU(h, cudaSetDevice(h->device));
    if (!h->sw_created) {
        CU(h, cudaEventCreate(&h->sw_start));
        CU(h, cudaEventCreate(&h->sw_stop));
        h->sw_created = true;
    }
    CU(h, cudaEventRecord(h->sw_start, h->stream));
    return NB200_OK;
}

int32_t nb200_timer_stop(nb200_handle* h, double* elapsed_ms) {
    if (!h || !elapsed_ms) return NB200_ERR_BAD_ARG;
    if (!h->sw_created) return fail(h, NB200_ERR_STATE, "nb200_timer_start was not called");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaEventRecord(h->sw_stop, h->stream));
    CU(h, cudaEventSynchronize(h->sw_stop));
    float ms = 0.f;
    CU(h, cudaEventElapsedTime(&ms, h->sw_start, h->sw_stop));
    *elapsed_ms = ms;
    return NB200_OK;
}

int32_t nb200_get_stats(nb200_handle* h, nb200_stats* out) {
    if (!h || !out) return NB200_ERR_BAD_ARG;
    CU(h, cudaSetDevice(h->device));
    int32_t rc = read_counters(h);
    if (rc) return rc;
    out->n_atoms = h->n;
    out->n_leaves = h->n_leaves;
    out->n_entries = (int64_t)h->counters_h->n_valid;
    out->n_slots = (int64_t)h->counters_h->n_entries();
    out->n_segments = (int64_t)h->counters_h->n_segments();
    out->entry_capacity = h->entry_capacity;
    out->kernel_launches = h->kernel_launches;
    out->steps_done = h->steps_done;
    out->regrows = h->regrows;
    out->n_pairs = list_pairs(h);
    out->list_half = h->list_half ? 1 : 0;
    return NB200_OK;
}

}  // extern "C"
