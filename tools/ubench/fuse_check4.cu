__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long sq2(unsigned long long a) {
    unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(a), "l"(0x8000000080000000ull)); return r; }
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__global__ void chk(const unsigned long long* in, unsigned long long* out) {
    unsigned long long a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64], q = in[threadIdx.x + 96];
    unsigned long long dx = sub2(a, q), dy = sub2(b, q), dz = sub2(c, q);
    out[threadIdx.x] = add2(add2(sq2(dx), sq2(dy)), sq2(dz));
}
