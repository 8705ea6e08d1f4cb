// Microbenchmark: cost of scattering pair forces with 16-byte vector reductions (red.global.add.v4.f32)
// into a window of ~600 atoms ahead of the warp's own leaf, 26 per lane (the half-list force pattern).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float4* force, int nleaf, int per_lane) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nleaf) return;
    unsigned h = warp * 2654435761u + lane * 40503u;
    int n = nleaf * 32;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int k = 0; k < per_lane; ++k) {
        h = h * 1664525u + 1013904223u;
        int j = warp * 32 + 32 + (h >> 8) % 576;
        if (j >= n) j -= n;
        float4 v = make_float4(1.f + k, 2.f, 3.f, 0.f);
        if (MODE == 0) atomicAdd(&force[j], v);
        if (MODE == 1) { atomicAdd(&force[j].x, v.x); atomicAdd(&force[j].y, v.y); atomicAdd(&force[j].z, v.z); }
        if (MODE == 2) { float4 p = __ldg(&force[j]); acc.x += p.x; acc.y += p.y; acc.z += p.z; }  // gather only
    }
    if (MODE == 2) force[warp * 32 + lane] = acc;
}
template <int MODE> void run(const char* name) {
    int nleaf = 31250, per_lane = 26;
    float4* f; cudaMalloc(&f, (size_t)nleaf * 32 * 16); cudaMemset(f, 0, (size_t)nleaf * 32 * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<(nleaf + 7) / 8, 256>>>(f, nleaf, per_lane);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) k<MODE><<<(nleaf + 7) / 8, 256>>>(f, nleaf, per_lane);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-22s %.4f ms per launch (26M pair scatters)  %s\n", name, ms / 10, cudaGetErrorString(cudaGetLastError()));
    cudaFree(f);
}
int main() { run<0>("red.v4.f32"); run<1>("3x red.f32"); run<2>("gather float4 (ref)"); return 0; }
