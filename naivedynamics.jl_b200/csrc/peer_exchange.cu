// peer_exchange.cu — multi-GPU halo exchange WITHOUT a collective: ghosts are pulled by our own kernel with
// loads from the peers' memory over NVLink / NVSwitch.
//
// Every rank PUBLISHES, in one peer-mapped allocation (CUDA IPC between processes, plain pointers inside one
// process), double buffered by step parity:
//     flag (monotonic publication counter) | owned positions float4{x,y,z,q}[n_own] x2 | leaf boxes x2 | atom ids x2
// where a publication leaf is 32 consecutive owned atoms in the rank's CURRENT curve order (the owned atoms are
// resident in sorted order like a single-GPU system, so these groups are as compact as the tree's leaves one step
// ago), its box is recomputed every step, and the id is the atom's index in the rank's hand-over order.
// After kick-drift a rank writes positions + boxes of step s into buffer s&1 and then releases flag = s+1 at
// system scope.  mg_pull_kernel on another rank spins (acquire, system scope) until the peer's flag reaches the
// step it needs, tests the peer's leaf boxes against its own slab box dilated by the cutoff, and copies only the
// atoms of nearby leaves that are themselves within the cutoff of the slab box: ~1.5 MB per rank per step at
// 1 M atoms per GPU instead of the 16 B x N all-gather (128 MB per rank at 8 GPUs).  Buffer s&1 is rewritten at
// step s+2, which a rank can only reach after it has seen every peer publish s+1, i.e. after every peer has
// finished pulling step s — the flags are the only synchronisation, there is no barrier and no NCCL call.
#include <cstdlib>

#include "nb200_internal.cuh"
#include "curve.cuh"
#include "slab_grid.cuh"

namespace nb200 {

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// box of every publication leaf (32 consecutive owned atoms): [leaf][0] = min, [leaf][1] = max
__global__ void mg_leafbox_kernel(const float4* __restrict__ pos, int n, float4* __restrict__ box) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const float inf = __int_as_float(0x7f800000);
    float3 lo = make_float3(inf, inf, inf), hi = make_float3(-inf, -inf, -inf);
    if (i < n) {
        const float4 p = pos[i];
        lo = make_float3(p.x, p.y, p.z);
        hi = lo;
    }
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(full, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(full, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(full, lo.z, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(full, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(full, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(full, hi.z, o));
    }
    if (lane == 0 && (i - lane) < n) {
        box[2 * (size_t)(i >> 5)] = make_float4(lo.x, lo.y, lo.z, 0.f);
        box[2 * (size_t)(i >> 5) + 1] = make_float4(hi.x, hi.y, hi.z, 0.f);
    }
}

// Everything this stream wrote before (positions, boxes) becomes visible to the peers, then the counter moves.
__global__ void mg_release_flag_kernel(unsigned int* flag, unsigned int value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- occupancy grid of the slab -----------------------------------------------------------------------------------
// A Morton slab of equal atom count is not a box: on clustered data its AABB can span the whole domain, and "within
// the cutoff of the slab's AABB" would select almost every foreign atom as a ghost.  The slab is therefore also
// described by a 64^3 occupancy grid (one 64-bit word per (y,z) row): cells holding an owned atom are marked and
// dilated by R = ceil(cutoff / cell) cells, so a foreign atom within the cutoff of ANY owned atom always falls into a
// set cell (cell indices are clamped monotonically, which cannot increase an index distance).  A foreign atom is a
// ghost iff it passes BOTH the AABB test and the grid test; a peer leaf is pulled iff its box passes both.
__global__ void mg_grid_mark_kernel(const float4* __restrict__ pos, int n, GridQ q, unsigned long long* __restrict__ raw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned full = 0xffffffffu;
    int row = -1 - (int)(threadIdx.x & 31);  // out-of-range lanes: a row of their own, nothing to mark
    unsigned long long bit = 0ull;
    if (i < n) {
        const float4 p = pos[i];
        const int cx = grid_cell(p.x, q.lo[0], q.scale[0]), cy = grid_cell(p.y, q.lo[1], q.scale[1]), cz = grid_cell(p.z, q.lo[2], q.scale[2]);
        row = cz * GRID + cy;
        bit = 1ull << cx;
    }
    // 32 consecutive owned atoms share a handful of rows: one atomic per distinct row of the warp, and only if the bits
    // are not there yet (read at L2 — an L1-cached read would keep returning the zeros of the memset)
    const unsigned peers = __match_any_sync(full, row);
    const unsigned lo = __reduce_or_sync(peers, (unsigned)bit), hi = __reduce_or_sync(peers, (unsigned)(bit >> 32));
    const unsigned long long bits = ((unsigned long long)hi << 32) | lo;
    if (row >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
        unsigned long long* w = &raw[row];
        if ((__ldcg(w) & bits) != bits) atomicOr(w, bits);
    }
}

// grid[z][y] = OR over |dz|,|dy| <= R of the x-dilated raw rows; R >= GRID means "no grid information" (all ones)
__global__ void mg_grid_dilate_kernel(const unsigned long long* __restrict__ raw, int R, unsigned long long* __restrict__ grid) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= GRID * GRID) return;
    if (R >= GRID) { grid[t] = ~0ull; return; }
    const int z = t / GRID, y = t % GRID;
    unsigned long long acc = 0ull;
    for (int dz = -R; dz <= R; ++dz) {
        const int zz = z + dz;
        if (zz < 0 || zz >= GRID) continue;
        for (int dy = -R; dy <= R; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= GRID) continue;
            acc |= raw[zz * GRID + yy];
        }
    }
    unsigned long long d = acc;
    for (int k = 1; k <= R; ++k) d |= (acc << k) | (acc >> k);
    grid[t] = d;
}

// ---- migration ------------------------------------------------------------------------------------------------------
// Ownership follows the atoms: rank g owns the atoms whose 30-bit MORTON key (global box) lies in [split[g], split[g+1]).
// Every k-th step, right after the kick-drift, each rank sends the atoms that left its range: the record (position,
// velocity, global id, destination) goes to an OUTBOX in its published region, the atom's sort key becomes 0xffffffff
// (it sorts behind every real atom and falls off the end of the owned segment) and — the atom is still a neighbour —
// it joins this step's ghosts.  The destination finds it in the peers' outboxes and appends it to its own pre-sort
// arrays; its ghost pull skips atoms that key into its own range (they arrive through the outbox, not as ghosts).
__global__ void mg_classify_kernel(const float4* __restrict__ pos, const float4* __restrict__ vel, const int32_t* __restrict__ id,
                                   uint32_t* __restrict__ sort_keys, int n, BoxQ mq, BoxQ sq, const uint32_t* __restrict__ split, int world, int rank,
                                   float4* __restrict__ out_pos, float4* __restrict__ out_vel, int32_t* __restrict__ out_gid,
                                   int32_t* __restrict__ out_dest, unsigned int* __restrict__ out_count, unsigned int out_cap,
                                   float4* __restrict__ gpos, int32_t* __restrict__ ggidx, uint32_t* __restrict__ gkeys, uint32_t* __restrict__ gvals,
                                   unsigned int* __restrict__ ghost_count, unsigned int ghost_cap) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    int dest = rank;
    if (s < n) {
        p = pos[s];
        dest = mg_owner(morton30(p.x, p.y, p.z, mq), split, world);
    }
    const bool leaves = dest != rank;
    const unsigned m = __ballot_sync(full, leaves);
    if (m == 0u) return;
    unsigned ob = 0, gb = 0;
    if (lane == 0) {
        ob = atomicAdd(out_count, (unsigned)__popc(m));
        gb = atomicAdd(ghost_count, (unsigned)__popc(m));
    }
    ob = __shfl_sync(full, ob, 0);
    gb = __shfl_sync(full, gb, 0);
    if (leaves) {
        const unsigned k = __popc(m & ((1u << lane) - 1u));
        const int gid = id[s];
        if (ob + k < out_cap) {
            out_pos[ob + k] = p;
            out_vel[ob + k] = vel[s];
            out_gid[ob + k] = gid;
            out_dest[ob + k] = dest;
        }
        if (gb + k < ghost_cap) {
            gpos[gb + k] = p;
            ggidx[gb + k] = gid;
            gkeys[gb + k] = morton30(p.x, p.y, p.z, sq);
            gvals[gb + k] = gb + k;
        }
        sort_keys[s] = 0xffffffffu;
    }
}

// grid = (blocks over the largest outbox, world): records addressed to me are appended behind my owned atoms
__global__ void mg_immigrate_kernel(const MgPeer* __restrict__ peers, int rank, float4* __restrict__ pos, float4* __restrict__ vel,
                                    int32_t* __restrict__ id, uint32_t* __restrict__ sort_keys, uint32_t* __restrict__ sort_vals, int n_old, int room,
                                    unsigned int* __restrict__ in_count, BoxQ sq) {
    const int p = blockIdx.y;
    if (p == rank) return;
    const MgPeer P = peers[p];
    const unsigned cnt = min(__ldcg(&P.flag[HDR_OUT]), (unsigned)P.out_cap);
    const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool mine = r < cnt && __ldcg(&P.out_dest[r]) == rank;
    const unsigned m = __ballot_sync(full, mine);
    if (m == 0u) return;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(in_count, (unsigned)__popc(m));
    base = __shfl_sync(full, base, 0);
    if (mine) {
        const unsigned k = base + __popc(m & ((1u << lane) - 1u));
        if ((int)k < room) {
            const float4 q = __ldcg(&P.out_pos[r]);
            const int slot = n_old + (int)k;
            pos[slot] = q;
            vel[slot] = __ldcg(&P.out_vel[r]);
            id[slot] = __ldcg(&P.out_gid[r]);
            sort_keys[slot] = morton30(q.x, q.y, q.z, sq);
            sort_vals[slot] = (uint32_t)slot;
        }
    }
}

__global__ void add_offset_kernel(int32_t* __restrict__ v, int n, int off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += off;
}

// Waits for the publication flags of all peers: ONE thread per peer polls (acquire, system scope) with a growing
// back-off.  (When every block of the pull kernel polled for itself, ~250 pollers per rank kept re-reading a flag in the
// slower peer's memory over NVLink for as long as that peer was still busy with the previous step — and its traversal
// took up to twice as long.)  A peer that never publishes is reported after the time limit instead of hanging the GPU.
// The step every peer must have published is this rank's OWN publication count (its flag word): the ranks move in
// lockstep, and reading it on the device keeps the step loop free of per-step host values (CUDA-graph replay).
__global__ void mg_wait_kernel(const MgPeer* __restrict__ peers, int world, int rank, unsigned int* __restrict__ err,
                               long long spin_limit_cycles) {
    const int p = threadIdx.x;
    if (p >= world || p == rank) return;
    const unsigned int want_flag = *(volatile const unsigned int*)peers[rank].flag;
    const unsigned int* flag = peers[p].flag;
    const long long t0 = clock64();
    unsigned ns = 64;
    if (*(volatile unsigned int*)err) return;  // a peer already failed to publish: do not wait the time limit again on every step
    while (ld_acquire_sys(flag) < want_flag) {
        __nanosleep(ns);
        if (ns < 2048) ns <<= 1;
        if (clock64() - t0 > spin_limit_cycles) {
            atomicExch(err, 1u + (unsigned)p);
            break;
        }
    }
}

// grid = (blocks over the largest peer's leaves, world).  One warp tests 32 publication leaves of peer
// blockIdx.y against my slab box, then pulls the atoms of the near ones with coalesced 512-byte peer loads, PULL_BATCH
// leaves at a time: their loads are in flight together and ONE atomic reserves the ghost slots of the batch (one
// atomic per leaf left every warp waiting for ~3 k serialised round trips to the same counter).
// The last block to finish records the ghost statistics the asynchronous step needs (largest count, overflow, latest).
constexpr int PULL_BATCH = 4;
__global__ void __launch_bounds__(TPB)
    mg_pull_kernel(const MgPeer* __restrict__ peers, int rank, int parity, const int* __restrict__ box6,
                   float cutoff, float4* __restrict__ pos_out, int32_t* __restrict__ gidx_out,
                   unsigned int* __restrict__ ghost_count, unsigned int ghost_capacity, BoxQ bq, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                   const unsigned long long* __restrict__ grid, GridQ gq, unsigned int* __restrict__ stat, unsigned int* __restrict__ done,
                   const uint32_t* __restrict__ split, int world, BoxQ mq) {
    const int p = blockIdx.y;
    const MgPeer P = peers[p];
    const int p_n = p != rank ? (int)__ldcg(&P.flag[HDR_NPUB + parity]) : 0;  // owned atoms in the peer's publication of this step
    const int n_leaves = (p_n + 31) >> 5;
    const bool work = p != rank && (long long)blockIdx.x * TPB < n_leaves;
    if (work) {
        // (mg_wait_kernel has seen every peer's publication flag of this step)
        const unsigned full = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        const float3 lo = make_float3(ord2f(box6[0]), ord2f(box6[1]), ord2f(box6[2]));
        const float3 hi = make_float3(ord2f(box6[3]), ord2f(box6[4]), ord2f(box6[5]));
        const float r2 = cutoff * cutoff;
        const float r2pad = fmaf(r2, 4e-6f, r2) + 1e-37f;  // same conservative pad as the traversal
        const float4* __restrict__ pbox = P.box[parity];
        const float4* __restrict__ ppos = P.pos[parity];
        const int32_t* __restrict__ pid = P.id[parity];
        const int leaf = blockIdx.x * TPB + threadIdx.x;
        bool near_leaf = false;
        if (leaf < n_leaves) {
            const float4 blo = __ldcg(&pbox[2 * (size_t)leaf]), bhi = __ldcg(&pbox[2 * (size_t)leaf + 1]);
            const float gx = fmaxf(0.f, fmaxf(lo.x - bhi.x, blo.x - hi.x));
            const float gy = fmaxf(0.f, fmaxf(lo.y - bhi.y, blo.y - hi.y));
            const float gz = fmaxf(0.f, fmaxf(lo.z - bhi.z, blo.z - hi.z));
            near_leaf = gx * gx + gy * gy + gz * gz <= r2pad && (grid == nullptr || grid_box(grid, gq, blo, bhi));
        }
        unsigned sel = __ballot_sync(full, near_leaf);
        const int leaf0 = leaf - lane;
        while (sel) {
            int a[PULL_BATCH];
            float4 q[PULL_BATCH];
            bool ghost[PULL_BATCH];
#pragma unroll
            for (int u = 0; u < PULL_BATCH; ++u) {
                a[u] = -1;
                if (sel) {
                    const int b = __ffs(sel) - 1;
                    sel &= sel - 1;
                    a[u] = (leaf0 + b) * 32 + lane;  // atom of the peer's owned array
                }
                q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a[u] >= 0 && a[u] < p_n) q[u] = __ldcg(&ppos[a[u]]);
            }
            unsigned m[PULL_BATCH];
            int total = 0;
#pragma unroll
            for (int u = 0; u < PULL_BATCH; ++u) {
                ghost[u] = false;
                if (a[u] >= 0 && a[u] < p_n) {
                    const float gx = fmaxf(0.f, fmaxf(lo.x - q[u].x, q[u].x - hi.x));
                    const float gy = fmaxf(0.f, fmaxf(lo.y - q[u].y, q[u].y - hi.y));
                    const float gz = fmaxf(0.f, fmaxf(lo.z - q[u].z, q[u].z - hi.z));
                    ghost[u] = gx * gx + gy * gy + gz * gz <= r2pad && (grid == nullptr || grid_point(grid, gq, q[u].x, q[u].y, q[u].z));
                    // migration step: an atom whose key is in MY range arrives through the peer's outbox as an owned atom
                    if (ghost[u] && split && mg_owner(morton30(q[u].x, q[u].y, q[u].z, mq), split, world) == rank) ghost[u] = false;
                }
                m[u] = __ballot_sync(full, ghost[u]);
                total += __popc(m[u]);
            }
            if (total) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(ghost_count, (unsigned)total);
                base = __shfl_sync(full, base, 0);
#pragma unroll
                for (int u = 0; u < PULL_BATCH; ++u) {
                    if (ghost[u]) {
                        const unsigned g = base + __popc(m[u] & ((1u << lane) - 1u));
                        if (g < ghost_capacity) {
                            pos_out[g] = q[u];
                            gidx_out[g] = __ldcg(&pid[a[u]]);  // global id of the atom
                            keys[g] = morton30(q[u].x, q[u].y, q[u].z, bq);
                            vals[g] = (uint32_t)g;
                        }
                    }
                    base += __popc(m[u]);
                }
            }
        }
    }
    if (stat) {
        // last block done: remember the largest ghost count seen and whether the capacity was ever exceeded (the
        // asynchronous step cannot look at the count before it launches the rest of the pipeline)
        __shared__ bool last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x * gridDim.y - 1;
        __syncthreads();
        if (last && threadIdx.x == 0) {
            *done = 0u;
            const unsigned int g = *(volatile unsigned int*)ghost_count;
            if (g > stat[0]) stat[0] = g;          // max ghosts since the last nb200_mg_sync
            if (g > ghost_capacity) stat[1] = 1u;  // sticky overflow
            stat[2] = g;                           // latest
        }
    }
}

}  // namespace

int launch_mg_publish(cudaStream_t s, const float4* pos, int n_own, float4* box, unsigned int* flag, unsigned int value) {
    mg_leafbox_kernel<<<(n_own + TPB - 1) / TPB, TPB, 0, s>>>(pos, n_own, box);
    mg_release_flag_kernel<<<1, 1, 0, s>>>(flag, value);
    return 2;
}

// ghost pre-sort slots [from, to): inert NaN placeholders with their keys (see integrate_kernel<PUBLISH>, which does
// this inside the step loop)
__global__ void mg_ghost_fill_kernel(float4* __restrict__ gpos, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int from, int to, BoxQ q) {
    const int j = from + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= to) return;
    const float nan = __int_as_float(0x7fc00000);
    gpos[j] = make_float4(nan, nan, nan, 0.f);
    keys[j] = morton30(nan, nan, nan, q);
    vals[j] = (uint32_t)j;
}

int launch_mg_ghost_fill(cudaStream_t s, float4* gpos, uint32_t* keys, uint32_t* vals, int from, int to, const float* bmin, const float* bmax,
                         int hilbert) {
    if (to <= from) return 0;
    mg_ghost_fill_kernel<<<(to - from + TPB - 1) / TPB, TPB, 0, s>>>(gpos, keys, vals, from, to, make_boxq(bmin, bmax, hilbert));
    return 1;
}

// grid2: [0, 4096) raw marks, [4096, 8192) dilated grid (what the selection kernels read)
int launch_mg_grid(cudaStream_t s, const float4* own_pos, int n_own, const float* bmin, const float* bmax, float cutoff,
                   unsigned long long* grid2) {
    const GridQ q = make_gridq(bmin, bmax);
    float smax = fmaxf(q.scale[0], fmaxf(q.scale[1], q.scale[2]));
    int R = smax > 0.f ? (int)ceilf(cutoff * smax) : GRID;  // cells; a degenerate box carries no grid information
    if (R < 1) R = 1;
    if (R > 8) R = GRID;                                     // huge cutoffs: the AABB test alone decides
    cudaMemsetAsync(grid2, 0, sizeof(unsigned long long) * GRID * GRID, s);
    mg_grid_mark_kernel<<<(n_own + TPB - 1) / TPB, TPB, 0, s>>>(own_pos, n_own, q, grid2);
    mg_grid_dilate_kernel<<<(GRID * GRID + TPB - 1) / TPB, TPB, 0, s>>>(grid2, R, grid2 + GRID * GRID);
    return 2;
}

int launch_mg_release_flag(cudaStream_t s, unsigned int* flag, unsigned int value) {
    mg_release_flag_kernel<<<1, 1, 0, s>>>(flag, value);
    return 1;
}

// Ghosts of this step into the ghost PRE-SORT arrays (slot g in [0, count)): positions, global ids, curve keys.
// split != nullptr (migration step): atoms that key into this rank's own range are skipped (they arrive through the
// outboxes); keep_count: the leavers already sit at the front of the ghost arrays; wait_flags = false: the caller has
// already waited for the peers' flags on this stream.
int launch_mg_pull(cudaStream_t s, const MgPeer* peers_dev, int world, int rank, int max_peer_own, int parity,
                   const float4* own_pos, int n_own, const int* box6, float cutoff, float4* gpos, int32_t* ggidx, unsigned int* ghost_count,
                   int64_t ghost_capacity, unsigned int* err, long long spin_limit_cycles, unsigned int* ghost_stat, const float* bmin,
                   const float* bmax, int hilbert, uint32_t* gkeys, uint32_t* gvals, unsigned long long* grid2, unsigned int* done,
                   const uint32_t* split, bool keep_count, bool wait_flags) {
    if (!keep_count) cudaMemsetAsync(ghost_count, 0, sizeof(unsigned int), s);
    // grid2 == nullptr: the slab is compact (its AABB is about as large as its atoms need), the AABB test alone decides
    int launches = grid2 ? launch_mg_grid(s, own_pos, n_own, bmin, bmax, cutoff, grid2) : 0;
    GridQ gq = make_gridq(bmin, bmax);
    const unsigned long long* occupancy = grid2 ? grid2 + GRID * GRID : nullptr;
    const BoxQ bq = make_boxq(bmin, bmax, hilbert);
    if (world > 1) {
        const int max_leaves = (max_peer_own + 31) / 32;
        dim3 grid((max_leaves + TPB - 1) / TPB, world);
        if (wait_flags) {
            mg_wait_kernel<<<1, 64, 0, s>>>(peers_dev, world, rank, err, spin_limit_cycles);
            ++launches;
        }
        mg_pull_kernel<<<grid, TPB, 0, s>>>(peers_dev, rank, parity, box6, cutoff, gpos, ggidx, ghost_count,
                                            (unsigned int)ghost_capacity, bq, gkeys, gvals, occupancy, gq, ghost_stat, done, split, world,
                                            make_boxq(bmin, bmax, 0));
        ++launches;
    } else if (ghost_stat) {
        cudaMemsetAsync(ghost_stat + 2, 0, sizeof(unsigned int), s);  // a single slab has no ghosts
    }
    return launches;
}

int launch_mg_classify(cudaStream_t s, const float4* pos, const float4* vel, const int32_t* id, uint32_t* sort_keys, int n, const float* bmin,
                       const float* bmax, int hilbert, const uint32_t* split, int world, int rank, float4* out_pos, float4* out_vel,
                       int32_t* out_gid, int32_t* out_dest, unsigned int* out_count, int out_cap, float4* gpos, int32_t* ggidx, uint32_t* gkeys,
                       uint32_t* gvals, unsigned int* ghost_count, unsigned int ghost_cap) {
    cudaMemsetAsync(out_count, 0, sizeof(unsigned int), s);
    cudaMemsetAsync(ghost_count, 0, sizeof(unsigned int), s);
    if (getenv("NB200_DEBUG_NO_LEAVERS")) return 0;  // (tuning aid: migration steps with their sync but without any hand-over)
    mg_classify_kernel<<<(n + TPB - 1) / TPB, TPB, 0, s>>>(pos, vel, id, sort_keys, n, make_boxq(bmin, bmax, 0), make_boxq(bmin, bmax, hilbert), split,
                                                          world, rank, out_pos, out_vel, out_gid, out_dest, out_count, (unsigned)out_cap, gpos, ggidx,
                                                          gkeys, gvals, ghost_count, ghost_cap);
    return 1;
}

int launch_mg_immigrate(cudaStream_t s, const MgPeer* peers_dev, int world, int rank, int max_out_cap, float4* pos, float4* vel, int32_t* id,
                        uint32_t* sort_keys, uint32_t* sort_vals, int n_old, int room, unsigned int* in_count, const float* bmin,
                        const float* bmax, int hilbert, unsigned int* err, long long spin_limit_cycles) {
    cudaMemsetAsync(in_count, 0, sizeof(unsigned int), s);
    mg_wait_kernel<<<1, 64, 0, s>>>(peers_dev, world, rank, err, spin_limit_cycles);
    dim3 grid((max_out_cap + TPB - 1) / TPB, world);
    mg_immigrate_kernel<<<grid, TPB, 0, s>>>(peers_dev, rank, pos, vel, id, sort_keys, sort_vals, n_old, room, in_count,
                                             make_boxq(bmin, bmax, hilbert));
    return 2;
}

int launch_add_offset(cudaStream_t s, int32_t* v, int n, int off) {
    if (n > 0 && off != 0) add_offset_kernel<<<(n + TPB - 1) / TPB, TPB, 0, s>>>(v, n, off);
    return 1;
}

// Tuning (NB200_CARVEOUT, see atoms.cu)
void carveout_peer(int pct) {
    cudaFuncSetAttribute(mg_classify_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(mg_wait_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(mg_pull_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(mg_ghost_fill_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(mg_grid_mark_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(mg_grid_dilate_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

}  // namespace nb200
