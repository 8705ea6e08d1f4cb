// slab_grid.cuh — 64^3 occupancy grid of a rank's slab, shared by the ghost selection kernels (peer_exchange.cu, atoms.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nb200 {

constexpr int GRID = 64;

struct GridQ {
    float lo[3], scale[3];  // cell = clamp(int((p - lo) * scale), 0, 63)
};
__device__ __forceinline__ int grid_cell(float p, float lo, float scale) {
    return min(max(__float2int_rd((p - lo) * scale), 0), GRID - 1);  // NaN -> 0
}

__device__ __forceinline__ bool grid_point(const unsigned long long* __restrict__ grid, const GridQ& q, float x, float y, float z) {
    const int cx = grid_cell(x, q.lo[0], q.scale[0]), cy = grid_cell(y, q.lo[1], q.scale[1]), cz = grid_cell(z, q.lo[2], q.scale[2]);
    return (grid[cz * GRID + cy] >> cx) & 1ull;
}

__device__ __forceinline__ bool grid_box(const unsigned long long* __restrict__ grid, const GridQ& q, const float4& blo, const float4& bhi) {
    const int x0 = grid_cell(blo.x, q.lo[0], q.scale[0]), x1 = grid_cell(bhi.x, q.lo[0], q.scale[0]);
    const int y0 = grid_cell(blo.y, q.lo[1], q.scale[1]), y1 = grid_cell(bhi.y, q.lo[1], q.scale[1]);
    const int z0 = grid_cell(blo.z, q.lo[2], q.scale[2]), z1 = grid_cell(bhi.z, q.lo[2], q.scale[2]);
    const unsigned long long xm = (x1 - x0 >= 63 ? ~0ull : ((1ull << (x1 - x0 + 1)) - 1ull) << x0);
    for (int z = z0; z <= z1; ++z)
        for (int y = y0; y <= y1; ++y)
            if (grid[z * GRID + y] & xm) return true;
    return false;
}


inline GridQ make_gridq(const float* bmin, const float* bmax) {
    GridQ q;
    for (int d = 0; d < 3; ++d) {
        const float ext = bmax[d] - bmin[d];
        q.lo[d] = bmin[d];
        q.scale[d] = ext > 0.f ? (float)GRID / ext : 0.f;
    }
    return q;
}

}  // namespace nb200
