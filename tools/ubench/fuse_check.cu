__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__global__ void chk(const unsigned long long* in, unsigned long long* out) {
    unsigned long long a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64], q = in[threadIdx.x + 96];
    unsigned long long dx = sub2(a, q), dy = sub2(b, q), dz = sub2(c, q);
    out[threadIdx.x] = add2(add2(mul2(dx, dx), mul2(dy, dy)), mul2(dz, dz));
}
__global__ void chk_scalar(const float* in, float* out) {
    float a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64], q = in[threadIdx.x+96];
    float dx = __fsub_rn(a, q), dy = __fsub_rn(b, q), dz = __fsub_rn(c, q);
    out[threadIdx.x] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
