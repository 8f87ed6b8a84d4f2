// Which pipe runs HSET2 (__hne2_mask)?  Throughput of VIADDMNMX.S16x2 alone, HSET2 alone, and both interleaved.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
template<int MODE> __global__ void k(uint32_t* out, long long* cyc, int iters, uint32_t seed)
{
    uint32_t x[8], m[8]; uint32_t y = seed * 3 + 1, z = seed ^ 0x1234;
    for (int i = 0; i < 8; ++i) { x[i] = seed + threadIdx.x * 8 + i; m[i] = 0x3c004000u + i + threadIdx.x; }
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int u = 0; u < 2; ++u)
        #pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 2) x[i] = __viaddmax_s16x2(x[i], y, z);
            if (MODE == 1 || MODE == 2) { __half2 a = *reinterpret_cast<__half2*>(&m[i]); __half2 b = *reinterpret_cast<__half2*>(&y); m[i] = __hne2_mask(a, b) + m[i] * 0 + (m[i] ^ 0x10001u); }
            if (MODE == 3) { __half2 a = *reinterpret_cast<__half2*>(&m[i]); __half2 b = *reinterpret_cast<__half2*>(&x[i]); uint32_t mk = __hne2_mask(a, b); x[i] = __viaddmax_s16x2(x[i], mk, z); }
            if (MODE == 4) { uint32_t mk; asm volatile("prmt.b32 %0,%1,%2,%3;" : "=r"(mk) : "r"(m[i]), "r"(y), "r"(x[i])); x[i] = __viaddmax_s16x2(x[i], mk, z); }
        }
    }
    long long t1 = clock64();
    uint32_t a = 0; for (int i = 0; i < 8; ++i) a ^= x[i] ^ m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template<typename K> void run(const char* name, K kern, int threads) {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 4 * 2048); cudaMalloc(&cyc, 64);
    const int iters = 20000;
    kern<<<1, threads>>>(out, cyc, 10, 1); kern<<<1, threads>>>(out, cyc, iters, 1); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s threads=%4d cycles per 16-op group per warp-slot = %.2f  [%s]\n", name, threads, (double)c / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    for (int th : {128, 1024}) {
        run("VIADDMNMX.S16x2 x16", k<0>, th);
        run("HSET2(hne2_mask)+LOP x16", k<1>, th);
        run("both independent x16+x16", k<2>, th);
        run("dependent HSET2->VIADDMNMX x16", k<3>, th);
        run("dependent PRMT->VIADDMNMX x16", k<4>, th);
    }
}
