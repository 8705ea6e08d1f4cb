// nb200_internal.cuh — handle layout and kernel launch prototypes shared by the .cu files.
// Not part of the ABI (see include/naiveb200.h for that).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/naiveb200.h"

namespace nb200 {

constexpr int LEAF = NB200_LEAF_SIZE;  // atoms per LBVH leaf == warp width: lane <-> atom
static_assert(LEAF == 32, "lane <-> atom mapping assumes 32-atom leaves");

// ---- neighbour list layout: cluster-pair hit-mask TILES ------------------------------------------
// The list is not expanded into one slot per pair.  A TILE pairs one query leaf (32 curve-consecutive atoms,
// lane <-> atom) with a block of 32 target atoms the traversal gathered for it and holds
//   words [0, 32)  : sorted slot of target b (-1 = no target; multi-GPU: slots >= the ghost base are ghosts)
//   words [32, 64) : hit mask of query atom q: bit b set <=> (q, target b) passes the reference's predicate
// i.e. 256 bytes for up to 1024 candidate pairs, written by the traversal's distance pass with two coalesced
// 128-byte stores and no per-hit work at all.  Tiles are allocated in GROUPS (one per drain pass of a query
// leaf, <= TGT_CAP/32 tiles) with ONE packed atomic; a group header names the leaf and its tile range.  The first
// tile of a leaf's first group is its SELF tile (targets = the leaf's own atoms).
// Two forms.  HALF (default): each unique pair appears once, in the tile row of its curve-earlier atom (self
// tile: bits above the diagonal).  DIRECTED: each pair appears in the rows of both atoms.
// Readers iterate the set bits of their mask.  The step loop does not read the list at all: the traversal evaluates the
// pair forces from the tile while it is still in shared memory (FUSED, traverse.cu); force_tiles_kernel (forces.cu) is the
// reader for lists that are reused, for energies on demand and for nb200_set_fused_force(0).
constexpr int TILE_WORDS = 64;
struct GroupHdr {
    int32_t leaf;         // query leaf index == first sorted atom / 32
    uint32_t ntiles;      // tiles in this group (0: group did not fit the buffer) | GROUP_SELF
    uint64_t base_tile;   // first tile of the group: words [base_tile * 64, (base_tile + ntiles) * 64) of the tile buffer
};
static_assert(sizeof(GroupHdr) == 16, "GroupHdr is 16 bytes");
constexpr uint32_t GROUP_SELF = 0x80000000u;  // the group's first tile is the leaf's self tile

// Allocation counter of the traversal: ONE 64-bit word so a group costs a single atomic on the hot address
// (same-address atomics serialise in L2).
//   alloc = (TILES requested so far) << SEG_BITS | (groups so far)
constexpr int SEG_BITS = 27;  // up to 134 M groups; 37 bits of tiles
struct Counters {
    unsigned long long alloc;      // packed tiles / groups (keeps counting past capacity)
    unsigned long long n_valid;    // set mask bits = unique pairs (half list) or 2 x unique pairs (directed); one add per leaf
    unsigned int overflow;         // set when a group did not fit
    unsigned int overflow_sticky;  // like overflow but only cleared by nb200_sync (async step loops)
    unsigned long long n_export;   // pairs written by the export kernel
    unsigned int stack_overflow;   // sticky: a traversal warp ran out of stack (tree deeper than the stack allows)
    unsigned int sticky_saved;     // overflow_sticky as it was when a synchronous search began (restored when that search succeeds)
    __host__ __device__ unsigned long long n_tiles() const { return alloc >> SEG_BITS; }
    __host__ __device__ unsigned long long n_entries() const { return (alloc >> SEG_BITS) * TILE_WORDS; }  // words in use
    __host__ __device__ unsigned int n_segments() const { return (unsigned int)(alloc & ((1ull << SEG_BITS) - 1ull)); }
};
constexpr size_t COUNTERS_RESET_BYTES = 20;  // alloc, n_valid, overflow: cleared before each traversal

// LBVH internal node, 64 B: both child boxes live in the parent so one traversal step is a single
// round of loads.  c[0] = {left.min.xyz, left id}, c[1] = {left.max.xyz, right id},
// c[2] = {right.min.xyz, range first leaf}, c[3] = {right.max.xyz, range last leaf}.
// child id >= 0: internal node; < 0: leaf ~id.
struct __align__(16) Node {
    float4 c[4];
};

// what a rank needs to know about a peer's publication (peer_exchange.cu); pointers are peer-mapped device memory
// Header of a rank's published region (first 256 bytes, 32-bit words).  HDR_FLAG: publication count (the step flag);
// HDR_NPUB + parity: owned atoms in that parity's publication; HDR_OUT: records in the migration outbox.
enum { HDR_FLAG = 0, HDR_NPUB = 1, HDR_OUT = 3, HDR_CAP = 4 };
struct MgPeer {
    const float4* pos[2];
    const float4* box[2];
    const int32_t* id[2];   // global id of the atom in each published slot
    const unsigned int* flag;   // == header base
    const float4* out_pos;  // migration outbox: atoms that left this rank's key range at the last migration step
    const float4* out_vel;
    const int32_t* out_gid;
    const int32_t* out_dest;
    int cap;                // slots per published array (the layout follows the capacity, not the current atom count)
    int out_cap;
};

// Scratch that later kernels need in a known state, initialised by the tail of reorder_kernel instead of by separate
// memset operations (each stream operation costs 2-3 us of GPU idle; the step loop had five of them):
//   node_flag[0..n_flag) = -1 and the traversal counters = 0 for THIS step's build / traverse,
//   the radix sort's histogram, tickets and look-back words = 0 for the NEXT step's sort.
struct Housekeeping {
    int32_t* node_flag; int n_flag;
    uint32_t* counters; int n_counter_words;
    uint32_t* sort_hist; int n_hist;
    uint32_t* sort_ticket; int n_ticket;
    uint32_t* sort_status; long long n_status;
};

// what integrate_kernel<PUBLISH> writes besides the integration itself (atoms.cu)
struct MgPublish {
    float4* pub_pos; int32_t* pub_id; const int32_t* id_in; float4* pub_box; unsigned int* n_pub;
    int* slab_box6; int* slab_box6_next;
    float4* g_pos; uint32_t* g_keys; uint32_t* g_vals; int g_fill;
    unsigned int* flag; unsigned int* done;   // the last block publishes flag = flag + 1
    const uint32_t* prev_keys; const uint32_t* split; int world, rank;   // strays keep their last in-range sort key (atoms.cu)
};

// multi-GPU search arrays: two sorted segments with a tree each (traverse.cu)
struct MgSearch {
    int n_query;               // owned atoms (slots [0, n_query) query)
    int32_t* blist;            // owned leaves near the ghost tree (filled by boundary_leaves_kernel right before the pass)
    unsigned int* bcount;      // their number
};

struct ForceField {
    float eps, sigma, kcoul, cutoff;
    int shift;
};

struct StageTimer {
    bool enabled;
    int only_stage;  // >= 0: only this stage is bracketed with events (the others run back to back)
    static constexpr int MAX_EVENTS = 8192;
    cudaEvent_t ev[MAX_EVENTS];
    int stage_of[MAX_EVENTS];  // event i..i+1 brackets stage_of[i] (or -1)
    int n_ev;
    double ms[NB200_STAGE_COUNT];
    int64_t launches[NB200_STAGE_COUNT];
    bool created;
};

}  // namespace nb200

struct nb200_handle {
    int device;
    cudaStream_t stream;
    int sm_count;
    int64_t n_max;
    int32_t n;         // atoms currently held
    int32_t n_leaves;  // ceil(n / 32)

    // atom state, kept in Morton-sorted order, double buffered (cur = buffer holding live state)
    float4* pos[2];    // x, y, z, charge
    float4* vel[2];    // vx, vy, vz, 1/mass
    int32_t* id[2];    // original atom id (0-based) of the atom in each slot
    float4* force;     // fx, fy, fz, potential energy share   (same order as pos[cur])
    int cur;
    bool vel_half;     // vel holds v(t - dt/2 .. ) i.e. the closing half kick is still pending
    bool have_system;
    bool have_forces;
    bool list_valid;
    bool pe_valid;     // force[].w holds the potential-energy shares of the current list
    int reuse_every;       // step loop: neighbour list rebuilt every k-th step (1 = every step) with cutoff + reuse_skin
    float reuse_skin;
    int list_age;          // steps since the list was built
    unsigned int* reuse_d2;  // device [2]: largest offending squared displacement (float bits), sticky violation flag
    bool hk_sort_clean;    // the sort scratch was zeroed by the last reorder_kernel for (hk_n, hk_passes)
    int64_t hk_n;
    int hk_passes;
    int resort_interval;   // step loop: full Morton re-sort every k-th step (1 = every step), leaf refresh in between
    int steps_since_sort;
    int sort_passes_override;  // > 0: number of 8-bit sort passes from the top of the key (tuning aid)
    bool use_graph;    // step loop: replay two captured steps as a CUDA graph in steady state (NB200_NO_GRAPH in the environment disables it)
    void* graph_exec;  // cudaGraphExec_t of two steps, valid for graph_key
    unsigned char graph_key[192];
    bool graph_is_mg;  // graph_exec holds two SLAB steps (nb200_mg_step_async)
    int64_t graph_launches;  // kernel launches in one replay
    bool fused_force;  // step loop: pair forces evaluated inside the traversal (nb200_set_fused_force; default on)
    int list_mode;     // requested NB200_LIST_HALF / NB200_LIST_DIRECTED
    bool list_half;    // form of the list currently in `entries` (multi-GPU searches are always directed)

    // Morton keys / permutation, double buffered for the LSD passes
    uint32_t* keys[2];
    uint32_t* vals[2];
    uint32_t* sort_hist;    // [4][256]
    uint32_t* sort_status;  // [4][tiles][256] decoupled look-back words
    uint32_t* sort_ticket;  // [4]
    int64_t sort_tiles_cap;

    // tree
    float4* leaf_lo;     // min.xyz, (float bits) atoms in leaf
    float4* leaf_hi;     // max.xyz, (float bits) Morton key of first atom
    float4* leaf_sub;    // 4 sub-boxes per leaf: [leaf][4][lo,hi]
    nb200::Node* nodes;  // n_leaves - 1
    float4* node_lo;     // merged box of each internal node (build scratch, dumped by get_tree)
    float4* node_hi;
    int32_t* node_flag;  // Apetrei range hand-off word per split
    int32_t* frontier;   // [0] = count, [1..32]: the tree's first levels expanded to <= 32 entries (lbvh_build.cu)

    // neighbour list
    int32_t* entries;
    int64_t entry_capacity;
    nb200::GroupHdr* segs;   // group headers of the tile list
    int64_t seg_capacity;
    nb200::Counters* counters;    // device
    nb200::Counters* counters_h;  // pinned host mirror

    // host<->device staging
    float* stage_dev;       // n_max * 4 floats
    int64_t stage_floats;
    void* scratch_dev;      // generic scratch for the list-based reference entry points
    int64_t scratch_bytes;
    float* pinned;          // pinned host staging
    int64_t pinned_bytes;

    // export buffers (sized on demand)
    int32_t* exp_a;
    int32_t* exp_b;
    float* exp_d;
    int64_t exp_capacity;

    float box_min[3], box_max[3];
    int curve;     // 0 Morton, 1 Hilbert: the order the atoms are sorted in
    float cutoff;  // cutoff of the current list
    nb200::ForceField ff;
    float last_dt;

    double* energy_dev;  // [2] KE, PE partial sums; [2] thermostat sum

    // position log of nb200_simulate: frames are unpacked into a ring of staging buffers and copied out on a
    // second stream, so the D2H of frame k overlaps the steps after it
    cudaStream_t log_stream;
    float* log_stage[3];
    int64_t log_stage_floats;
    cudaEvent_t log_ready[3], log_copied[3];
    bool log_created;

    // multi-GPU (Morton-slab partition): the OWNED atoms are the resident sorted system
    // (pos/vel/id[cur], slots [0, mg_n_own); id = index in the hand-over order); the GHOSTS of a step form a second
    // sorted segment from slot mg_gbase = 32 * owned leaves on, with their own tree
    bool mg_active;
    int32_t mg_n_own;
    int32_t mg_n_ghost;          // ghosts found by the last synchronous search
    int32_t mg_nLo, mg_gbase;    // owned leaves; first ghost slot
    int32_t mg_n_gslots;         // ghost slots in the current search arrays (exact count or capacity)
    void* mg_pub;                // published region: [flag (256 B) | pos x2 | leaf boxes x2 | hand-over ids x2]  (peer_exchange.cu)
    int64_t mg_pub_bytes;
    float4* mg_pub_pos[2];
    float4* mg_pub_box[2];
    int32_t* mg_pub_id[2];
    unsigned int* mg_flag;
    int mg_parity;
    unsigned int mg_pub_step;    // publications so far; flag value = mg_pub_step
    nb200::MgPeer* mg_peers_dev;
    int mg_world, mg_rank, mg_max_peer_own;
    int64_t mg_own_begin;
    bool mg_connected;
    void* mg_ipc_opened[64];     // peer regions opened with cudaIpcOpenMemHandle (closed in destroy)
    unsigned int* mg_err;        // device [4]: [0] set by mg_pull_kernel when a peer never published, [2], [3] last-block-done counters
    unsigned long long* mg_grid;  // occupancy grid of the slab: 2 x 64 x 64 words (raw marks, dilated)
    bool mg_use_grid;             // decided by the synchronous search: is the slab ragged (AABB much larger than its atoms need)?
    int64_t mg_n_total;           // atoms of all ranks (nb200_mg_connect)
    unsigned int* mg_ghost_stat; // device [4]: max ghosts since the last sync, sticky overflow, latest count
    int64_t mg_ghost_cap;        // ghost slots the asynchronous step provides (0: no synchronous search has run yet)
    int64_t mg_gfill;            // ghost pre-sort slots the last integrate pre-filled with NaN placeholders
    bool mg_pub_current;         // the publication holds the owned atoms' current positions
    bool mg_keys_ready;          // keys[0] / vals[0] hold the owned atoms' keys of the current positions (written by the integrate)
    unsigned int* mg_stat_h;     // pinned [4]: copy of mg_ghost_stat made by every asynchronous step (read with a lag)
    cudaEvent_t mg_step_ev[16];  // completion of the last 16 asynchronous steps: bounds how far the host runs ahead
    bool mg_ev_created;
    int64_t mg_async_steps;      // asynchronous steps so far
    int64_t mg_async_subs;       // submissions (steps or graph launches) so far: index into mg_step_ev
    float4* mg_gpos;             // ghost pre-sort arrays: positions, gathered (global) indices, curve keys / permutation
    int32_t* mg_ggidx;
    uint32_t* mg_gkeys[2];
    uint32_t* mg_gvals[2];
    uint32_t* sort_hist2;        // scratch of the ghost sort
    uint32_t* sort_status2;
    uint32_t* sort_ticket2;
    bool hk2_clean;
    int64_t hk2_n;
    int32_t* frontier2;          // frontier of the ghost tree
    int32_t* mg_blist;           // owned leaves the ghost pass has to visit (n_max / 32 entries) + their count
    unsigned int* mg_bcount;
    float4* mg_sendbuf;          // NCCL exchange: owned positions in hand-over order
    // migration: ownership follows the atoms (key ranges), see peer_exchange.cu
    uint32_t* mg_split;          // device [world + 1]: rank g owns Morton keys in [split[g], split[g+1])
    float4* mg_out_pos;          // outbox (in the published region)
    float4* mg_out_vel;
    int32_t* mg_out_gid;
    int32_t* mg_out_dest;
    int64_t mg_out_cap;
    int mg_max_peer_out;
    int mg_id_off;               // offset already added to the atom ids (0 until nb200_mg_connect makes them global)
    bool mg_migrated;            // some migration has happened: nb200_mg_get_owned returns rows in the current curve order
    int mg_migrate_every;        // 0: atoms never change rank
    int mg_steps_since_migration;
    bool mg_migration_pending;   // the last integrate classified the atoms; the next search takes the immigrants in
    bool mg_sort_extra;          // this step's owned sort needs the extra top pass (leaver marker)
    int32_t mg_last_out, mg_last_in;  // atoms that left / arrived at the last migration
    const uint32_t* mg_sorted_keys;   // sorted keys of the last search, aligned with the owned slots (null: not available)
    int32_t mg_n_pre;            // owned pre-sort entries of the step (owned atoms of the last step + immigrants)
    cudaStream_t mg_stream2;     // the ghost side of the asynchronous step
    cudaEvent_t mg_ev_int, mg_ev_ghost, mg_ev_owned;
    int mg_ahead;                // submissions the host may run ahead of the GPU (<= 16; NB200_MG_AHEAD)
    bool mg_graph_multi;
    bool mg_pull_late;          // enqueue the halo pull with the rest of the ghost side, after the owned pass was launched (default)
    bool mg_ghost_prio_normal;  // tuning: ghost stream at normal priority         // NB200_MG_GRAPH: graph replay of the slab step also for world > 1 (measured slower)
    bool mg_trace;               // NB200_MG_TRACE: timeline of the asynchronous step (tuning aid)
    cudaEvent_t mg_trace_ev[32 * 6];
    long long mg_trace_step, mg_trace_printed;
    int* mg_box;         // slab AABB (ordered-int encoding), 2 x 8 ints: one per publication parity
    unsigned int* mg_ghost_count;    // device
    unsigned int* mg_ghost_count_h;  // pinned
    bool owns_stream;

    nb200::StageTimer timer;
    cudaEvent_t sw_start, sw_stop;
    bool sw_created;
    int64_t kernel_launches;
    int64_t steps_done;
    int64_t regrows;
    bool async_overflow_possible;

    char err[512];
};

namespace nb200 {

// ---- launchers (each returns the number of kernels it launched) ------------------------------
int launch_pack(cudaStream_t s, const float* xyz_dev, int stride, const float* vel_dev, const float* mass_dev,
                const float* charge_dev, int n, float4* pos, float4* vel, int32_t* id);
// pos/vel of sorted slot s <- caller arrays (original order) through id[]; .w lanes are kept
int launch_refresh(cudaStream_t s, const float* xyz_dev, const float* vel_dev, int stride, const int32_t* id, int n,
                   float4* pos, float4* vel, const float* bmin = nullptr, const float* bmax = nullptr, int hilbert = 0,
                   uint32_t* keys = nullptr, uint32_t* vals = nullptr, int id_off = 0);
int launch_unpack_state(cudaStream_t s, const float4* pos, const float4* vel, const int32_t* id, int n, int stride, float* out_pos,
                        float* out_vel);
int launch_morton(cudaStream_t s, const float4* pos, int n, const float* bmin, const float* bmax, uint32_t* keys,
                  uint32_t* vals, int hilbert);
// pos_out == nullptr: positions are updated in place
int launch_integrate(cudaStream_t s, float4* pos, float4* vel, const float4* force, int n, float kick_dt, float dt,
                     const float* bmin, const float* bmax, uint32_t* keys, uint32_t* vals, int hilbert, float4* pos_out = nullptr,
                     const MgPublish* pub = nullptr);
// sorts (keys[0], vals[0]) using the [1] buffers as ping-pong; result ends in buffer *out_buf
// scratch_clean: hist / ticket / status were zeroed by the previous reorder_kernel (Housekeeping) for exactly this n and passes
int launch_sort(cudaStream_t s, uint32_t* keys[2], uint32_t* vals[2], int64_t n, uint32_t* hist, uint32_t* status,
                uint32_t* ticket, int* out_buf, int low_bit = 0, int passes = 4, bool scratch_clean = false);
Housekeeping sort_housekeeping(int64_t n, int passes, uint32_t* hist, uint32_t* status, uint32_t* ticket);
int64_t sort_tiles(int64_t n);
int launch_reorder(cudaStream_t s, const uint32_t* perm, const uint32_t* keys_sorted, const float4* pos_in,
                   const float4* vel_in, const int32_t* id_in, float4* pos_out, float4* vel_out, int32_t* id_out,
                   float4* force_zero, float4* leaf_lo, float4* leaf_hi, float4* leaf_sub, int n, float cutoff,
                   const Housekeeping* hk = nullptr);
// A leaf whose AABB is wider than this along some axis is "wide" (its run crosses a coarse cell boundary): only such
// leaves get sub-boxes (reorder_kernel) and use them (traverse_kernel) — the two must agree bit for bit.
__host__ __device__ inline float wide_leaf_limit(float cutoff) { return 3.0f * cutoff; }
// off: leaf / node index offset of a second tree kept in the same arrays (multi-GPU ghost tree); pointers are NOT pre-shifted
int launch_build(cudaStream_t s, const float4* leaf_lo, const float4* leaf_hi, int n_leaves, Node* nodes, float4* node_lo,
                 float4* node_hi, int32_t* node_flag, bool flags_clean = false, int off = 0);
// frontier: 64 ints (count, <= 32 entries) followed by 64 float4 (the entries' own boxes, lo / hi interleaved)
constexpr int FRONTIER_WORDS = 64 + 64 * 4;
int launch_frontier(cudaStream_t s, const Node* nodes, int n_leaves, int32_t* frontier, int off, const float4* node_lo, const float4* node_hi,
                    const float4* leaf_lo, const float4* leaf_hi);
int launch_traverse(cudaStream_t s, int sm_count, const Node* nodes, const int32_t* frontier, const float4* leaf_lo, const float4* leaf_hi,
                    const float4* leaf_sub, const float4* pos, int n, int n_leaves, float cutoff, int32_t* entries, int64_t entry_capacity,
                    GroupHdr* segs, int64_t seg_capacity, Counters* counters, bool half, long long* dbg = nullptr,
                    const MgSearch* mg = nullptr, bool counters_clean = false, const ForceField* fused_ff = nullptr,
                    float4* fused_force = nullptr);
int launch_force(cudaStream_t s, int sm_count, const GroupHdr* segs, const int32_t* entries, const Counters* counters,
                 int64_t seg_capacity, const float4* pos, float4* force, int n, ForceField ff, bool with_pe, bool half,
                 bool check_cutoff = false, int ghost_base = 0x7fffffff);
// list reuse: flags (sticky) any atom whose squared displacement since the list was built exceeds limit2
int launch_displacement_check(cudaStream_t s, const float4* pos, const float4* pos_ref, int n, float limit2, unsigned int* out2);
int launch_export(cudaStream_t s, int sm_count, const GroupHdr* segs, const int32_t* entries, Counters* counters,
                  int64_t seg_capacity, const float4* pos, const int32_t* id, int n, int32_t* a, int32_t* b, float* d, int64_t capacity,
                  int index_base);
int launch_export_directed(cudaStream_t s, int sm_count, const GroupHdr* segs, const int32_t* entries, Counters* counters,
                           int64_t seg_capacity, const float4* pos, const int32_t* id, int n, int32_t* a, int32_t* b, float* d,
                           int64_t capacity);
int launch_slab_box(cudaStream_t s, const float4* pos, int n, int* box6, bool init = true);  // init = false: extend the box
int launch_slab_box_init(cudaStream_t s, int* box6);
int launch_ghost_select(cudaStream_t s, const float4* all_pos, int64_t n_all, int64_t own_begin, int n_own, const int* box6,
                        float cutoff, float4* gpos, int32_t* ggidx, unsigned int* ghost_count, int64_t ghost_capacity,
                        const unsigned long long* grid, const float* bmin, const float* bmax, int hilbert, uint32_t* gkeys,
                        uint32_t* gvals);
int launch_mg_publish(cudaStream_t s, const float4* pos, int n_own, float4* box, unsigned int* flag, unsigned int value);
int launch_mg_release_flag(cudaStream_t s, unsigned int* flag, unsigned int value);
int launch_mg_ghost_fill(cudaStream_t s, float4* gpos, uint32_t* keys, uint32_t* vals, int from, int to, const float* bmin, const float* bmax,
                         int hilbert);
int launch_mg_pull(cudaStream_t s, const MgPeer* peers_dev, int world, int rank, int max_peer_own, int parity,
                   const float4* own_pos, int n_own, const int* box6, float cutoff, float4* gpos, int32_t* ggidx, unsigned int* ghost_count,
                   int64_t ghost_capacity, unsigned int* err, long long spin_limit_cycles, unsigned int* ghost_stat, const float* bmin,
                   const float* bmax, int hilbert, uint32_t* gkeys, uint32_t* gvals, unsigned long long* grid2, unsigned int* done,
                   const uint32_t* split = nullptr, bool keep_count = false, bool wait_flags = true);
// migration (peer_exchange.cu): atoms whose Morton key left this rank's range go to the outbox (and stay as ghosts of
// this step); atoms the peers sent here are appended behind the owned atoms before the sort
int launch_mg_classify(cudaStream_t s, const float4* pos, const float4* vel, const int32_t* id, uint32_t* sort_keys, int n, const float* bmin,
                       const float* bmax, int hilbert, const uint32_t* split, int world, int rank, float4* out_pos, float4* out_vel,
                       int32_t* out_gid, int32_t* out_dest, unsigned int* out_count, int out_cap, float4* gpos, int32_t* ggidx, uint32_t* gkeys,
                       uint32_t* gvals, unsigned int* ghost_count, unsigned int ghost_cap);
int launch_mg_immigrate(cudaStream_t s, const MgPeer* peers_dev, int world, int rank, int max_out_cap, float4* pos, float4* vel, int32_t* id,
                        uint32_t* sort_keys, uint32_t* sort_vals, int n_old, int room, unsigned int* in_count, const float* bmin,
                        const float* bmax, int hilbert, unsigned int* err, long long spin_limit_cycles);
int launch_add_offset(cudaStream_t s, int32_t* v, int n, int off);
// occupancy grid of the slab (peer_exchange.cu): grid2 = 2 x 64 x 64 words (raw marks, dilated grid)
int launch_mg_grid(cudaStream_t s, const float4* own_pos, int n_own, const float* bmin, const float* bmax, float cutoff,
                   unsigned long long* grid2);
int launch_compose(cudaStream_t s, const int32_t* id_sorted, int n_own, int ghost_base, int n, int own_begin, const int32_t* ghost_gidx,
                   int32_t* out);
int launch_unpack(cudaStream_t s, const float4* src, const int32_t* id, int n, int stride, float* out_dev, int mode,
                  const float4* force, float half_dt, int id_off = 0);
int launch_unsort4(cudaStream_t s, const float4* src, const int32_t* id, int n, float4* dst, int id_off = 0);  // dst[id[s] - id_off] = src[s]
int launch_energy(cudaStream_t s, const float4* vel, const float4* force, int n, float half_dt, double* out2);
int launch_rescale_velocity(cudaStream_t s, float4* vel, const float4* force, int n, float half_dt, float tf, float gamma, int physical,
                            double* sum_dev);
int launch_neighbor_counts(cudaStream_t s, int sm_count, const GroupHdr* segs, const int32_t* entries, const Counters* counters,
                           int64_t seg_capacity, const int32_t* id, int n, int32_t* counts, bool half);
int launch_lj_literal(cudaStream_t s, const int32_t* a, const float* d, int64_t np, int index_base, int n, double* acc,
                      float* force);
int launch_coulomb_literal(cudaStream_t s, const int32_t* a, const int32_t* b, const float* d, int64_t np, int index_base,
                           const float* charge, int n, float* force);
int launch_sum_forces(cudaStream_t s, float* out, const float* f1, const float* f2, int64_t n3);
int launch_verlet_literal(cudaStream_t s, float* pos, float* vel, const float* f, const float* fnext, const float* mass,
                          int n, float dt, const float* bmin, const float* bmax, int reflect);
// system setup (setup.cu): staged system (stride-4 positions and velocities, mass, charge) drawn on the device
int launch_setup_draw(cudaStream_t s, int n, uint64_t seed, const float* bmin, const float* bmax, float minmass, float maxmass,
                      float mincharge, float maxcharge, float temperature, int randomvelocity, float* sx, float* sv, float* sm,
                      float* sq, double* sum3);
// mark the lower-id atom of every listed pair and give each marked atom a new position (draw number `round`)
int launch_prune_redraw(cudaStream_t s, int n, uint64_t seed, uint32_t round, const float* bmin, const float* bmax,
                        const int32_t* pair_a, const int32_t* pair_b, int64_t np, int32_t* mark, float* sx,
                        unsigned long long* redrawn);

void carveout_sort(int pct);
void carveout_build(int pct);
void carveout_peer(int pct);
void carveout_atoms(int pct);
void carveout_traverse(int pct);

}  // namespace nb200
