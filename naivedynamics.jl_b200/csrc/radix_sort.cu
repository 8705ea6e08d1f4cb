// radix_sort.cu — onesweep-style LSD radix sort of (u32 key, u32 value) pairs.
//
// Replaces `sortperm(pos.morton_code)` + 3x `permute!` (BVHTraverse.jl:570-573; serial Base sort).
// 4 passes of 8 bits.  One upfront kernel histograms all four digits in a single read of the keys;
// each pass is ONE kernel: tiles take a ticket (so a tile only ever waits on tiles that already
// run), rank their keys with warp match-any (stable), publish the per-digit tile aggregate and
// resolve their global offsets by decoupled look-back on single 32-bit status words
// (2 flag bits + 30 count bits, so publication needs no fence), then exchange through shared
// memory and write digit runs coalesced.
// HBM traffic per pass: 8 B read + 8 B written per pair (+4 B/key once for the histogram).
#include "nb200_internal.cuh"

namespace nb200 {

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
#ifndef NB200_SORT_ITEMS
#define NB200_SORT_ITEMS 16
#endif
constexpr int ITEMS = NB200_SORT_ITEMS;
constexpr int TILE = SORT_THREADS * ITEMS;  // pairs per tile
constexpr int PASSES = 4;

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

__global__ void __launch_bounds__(256) sort_hist_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                                        uint32_t* __restrict__ hist, int low_bit, int passes) {
    __shared__ uint32_t sh[PASSES * RADIX];
    for (int i = threadIdx.x; i < PASSES * RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t k = keys[i];
#pragma unroll
        for (int p = 0; p < PASSES; ++p)
            if (p < passes) atomicAdd(&sh[p * RADIX + ((k >> (low_bit + p * RADIX_BITS)) & (RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PASSES * RADIX; i += blockDim.x) {
        uint32_t c = sh[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// exclusive scan of one value per thread over a 256-thread block
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp /*[8]*/) {
    const unsigned full = 0xffffffffu;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(full, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w)
        if (w < warp) woff += s_warp[w];
    __syncthreads();  // s_warp reusable afterwards
    return woff + incl - v;
}

__global__ void __launch_bounds__(SORT_THREADS)
    sort_pass_kernel(const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin, uint32_t* __restrict__ kout,
                     uint32_t* __restrict__ vout, uint32_t n, int shift, const uint32_t* __restrict__ hist,
                     volatile uint32_t* status, uint32_t* ticket) {
    __shared__ uint32_t s_wcount[SORT_WARPS][RADIX];  // per-warp digit counts -> warp offsets
    __shared__ uint32_t s_keys[TILE];
    __shared__ uint32_t s_vals[TILE];
    __shared__ uint32_t s_tilebase[RADIX];  // first slot of digit d in the tile-sorted order
    __shared__ uint32_t s_gbase[RADIX];     // global slot of tile-sorted position p with digit d = s_gbase[d] + p
    __shared__ uint32_t s_scan[SORT_WARPS];
    __shared__ uint32_t s_tile;

    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&s_wcount[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * TILE;
    const uint32_t tile_count = min((uint32_t)TILE, n - tile_base);

    // ---- load: warp w owns a contiguous 512-pair chunk, item k of lane l is chunk[k*32 + l] ----------
    // (the values are only needed for the exchange and are loaded there: 16 registers fewer live across the ranking)
    uint32_t key[ITEMS], rank[ITEMS];
    const uint32_t wbase = warp * (32 * ITEMS);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        uint32_t p = wbase + k * 32 + lane;
        key[k] = p < tile_count ? kin[tile_base + p] : 0xffffffffu;
    }

    // ---- stable rank inside the warp chunk -----------------------------------------------------------
    // Three independent sweeps instead of one serial chain per item: (1) all match-any votes, (2) one shared-memory
    // atomic per distinct digit and item by the group's first lane — a warp's atomics on one address are performed
    // in issue order, which is what keeps the sort stable — (3) broadcast of the group base + position in the group.
    // Invalid items vote with a digit of their own so they neither match anything nor diverge.
    {
        uint32_t peers[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const bool valid = wbase + k * 32 + lane < tile_count;
            const uint32_t d = valid ? ((key[k] >> shift) & (RADIX - 1)) : (uint32_t)(RADIX + lane);
            peers[k] = __match_any_sync(full, d);
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const bool valid = wbase + k * 32 + lane < tile_count;
            const uint32_t d = (key[k] >> shift) & (RADIX - 1);
            rank[k] = 0;
            if (valid && lane == __ffs(peers[k]) - 1) rank[k] = atomicAdd(&s_wcount[warp][d], (uint32_t)__popc(peers[k]));
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k)
            rank[k] = __shfl_sync(full, rank[k], __ffs(peers[k]) - 1) + __popc(peers[k] & lt_mask);
    }
    __syncthreads();

    // ---- per-digit (thread == digit): warp offsets, tile aggregate, look-back --------------------------
    {
        const int d = tid;
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t c = s_wcount[w][d];
            s_wcount[w][d] = sum;
            sum += c;
        }
        const uint32_t tile_digit_count = sum;
        volatile uint32_t* my_status = status + (size_t)tile * RADIX + d;
        uint32_t excl = 0;
        if (tile == 0) {
            *my_status = FLAG_PREFIX | tile_digit_count;
        } else {
            *my_status = FLAG_AGG | tile_digit_count;
            // Walk back over the predecessors' status words, LOOKBACK of them in flight at a time: with all
            // tiles resident in one wave the walk is long (every tile publishes its aggregate at about the
            // same moment), and one dependent L2 load per predecessor made the pass latency-bound.
            constexpr int LOOKBACK = 16;
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t w[LOOKBACK];
#pragma unroll
                for (int b = 0; b < LOOKBACK; ++b) w[b] = (t - b >= 0) ? status[(size_t)(t - b) * RADIX + d] : (2u << 30) /* FLAG_PREFIX | 0 */;
#pragma unroll
                for (int b = 0; b < LOOKBACK; ++b) {
                    if (done) break;
                    uint32_t flag = w[b] & FLAG_MASK;
                    if (flag == 0) break;  // not published yet: re-read from here
                    excl += w[b] & VALUE_MASK;
                    --t;
                    if (flag == FLAG_PREFIX) done = true;
                }
            }
            *my_status = FLAG_PREFIX | (excl + tile_digit_count);
        }
        uint32_t tb = block_excl_scan_256(tile_digit_count, s_scan);  // tile-local digit base
        uint32_t gb = block_excl_scan_256(hist[d], s_scan);           // global digit base
        s_tilebase[d] = tb;
        s_gbase[d] = gb + excl - tb;
    }
    __syncthreads();

    // ---- exchange through shared memory into tile-sorted order -----------------------------------------
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        uint32_t p = wbase + k * 32 + lane;
        if (p < tile_count) {
            uint32_t d = (key[k] >> shift) & (RADIX - 1);
            uint32_t slot = s_tilebase[d] + s_wcount[warp][d] + rank[k];
            s_keys[slot] = key[k];
            s_vals[slot] = vin[tile_base + p];
        }
    }
    __syncthreads();

    // ---- coalesced write of digit runs --------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        uint32_t p = tid + k * SORT_THREADS;
        if (p < tile_count) {
            uint32_t kk = s_keys[p];
            uint32_t d = (kk >> shift) & (RADIX - 1);
            uint32_t g = s_gbase[d] + p;
            kout[g] = kk;
            vout[g] = s_vals[p];
        }
    }
}

}  // namespace

int64_t sort_tiles(int64_t n) { return (n + TILE - 1) / TILE; }

// Sorts by key bits [low_bit, low_bit + 8*passes): the pipeline drops the lowest bits of the 30-bit curve key
// when the atom count does not need them (stable sort: ties keep the previous step's order).
Housekeeping sort_housekeeping(int64_t n, int passes, uint32_t* hist, uint32_t* status, uint32_t* ticket) {
    Housekeeping hk = {};
    hk.sort_hist = hist; hk.n_hist = PASSES * RADIX;
    hk.sort_ticket = ticket; hk.n_ticket = PASSES;
    hk.sort_status = status; hk.n_status = (long long)passes * sort_tiles(n) * RADIX;
    return hk;
}

int launch_sort(cudaStream_t s, uint32_t* keys[2], uint32_t* vals[2], int64_t n, uint32_t* hist, uint32_t* status,
                uint32_t* ticket, int* out_buf, int low_bit, int passes, bool scratch_clean) {
    int launches = 0;
    if (n <= 0) { *out_buf = 0; return 0; }
    const int64_t tiles = sort_tiles(n);
    if (!scratch_clean) {
        cudaMemsetAsync(hist, 0, sizeof(uint32_t) * PASSES * RADIX, s);
        cudaMemsetAsync(ticket, 0, sizeof(uint32_t) * PASSES, s);
        cudaMemsetAsync(status, 0, sizeof(uint32_t) * PASSES * tiles * RADIX, s);
    }
    int hblocks = (int)((n + 256 * 16 - 1) / (256 * 16));
    if (hblocks > 148 * 8) hblocks = 148 * 8;
    sort_hist_kernel<<<hblocks, 256, 0, s>>>(keys[0], n, hist, low_bit, passes);
    ++launches;
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        sort_pass_kernel<<<(unsigned)tiles, SORT_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1],
                                                                 (uint32_t)n, low_bit + p * RADIX_BITS, hist + p * RADIX,
                                                                 status + (size_t)p * tiles * RADIX, ticket + p);
        ++launches;
        cur ^= 1;
    }
    *out_buf = cur;
    return launches;
}

// Tuning (NB200_CARVEOUT, see atoms.cu)
void carveout_sort(int pct) {
    cudaFuncSetAttribute(sort_hist_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(sort_pass_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

}  // namespace nb200
