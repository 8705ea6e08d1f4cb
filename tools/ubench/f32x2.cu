// Microbenchmark: issue/pipe rate of scalar FADD/FMUL vs packed add.f32x2 / mul.f32x2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[8]; unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
    float c = seed * 0.999f; unsigned long long c2 = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] = __fadd_rn(a[i], c); a[i] = __fmul_rn(a[i], c); }
            if (MODE == 1) { p[i] = add2(p[i], c2); p[i] = mul2(p[i], c2); }
            if (MODE == 2) { a[i] = __fadd_rn(a[i], c); a[i] = __fmul_rn(a[i], c); p[i] = add2(p[i], c2); p[i] = mul2(p[i], c2);}  // mix
            if (MODE == 3) { a[i] = __fadd_rn(a[i], c); a[i] = __fmul_rn(a[i], c); p[i] = p[i] + 0x9e3779b9ull * (p[i] >> 7); } // FP + int
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int flops_per_iter_thread) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 100, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)148 * 8 * 256 * iters * flops_per_iter_thread;
    printf("%-28s %.3f ms  %.2f Tflop-ops/s (scalar-equivalent ops)  err=%s\n", name, ms, ops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    run<0>("scalar fadd+fmul", 16);
    run<1>("packed add2+mul2", 32);
    run<2>("scalar + packed mix", 48);
    run<3>("scalar + int", 16);
    return 0;
}
