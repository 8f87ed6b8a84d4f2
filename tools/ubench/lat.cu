// Latency / throughput microbenchmarks for the instructions on the NW critical path (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define FULL 0xffffffffu
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s){uint32_t d; asm volatile("prmt.b32 %0,%1,%2,%3;":"=r"(d):"r"(a),"r"(b),"r"(s)); return d;}

template<int MODE> __global__ void lat_kernel(uint32_t* out, long long* cyc, int iters, uint32_t seed)
{
    const int lane = threadIdx.x & 31;
    uint32_t x = seed + threadIdx.x, y = seed * 3 + 1, z = seed ^ 0x1234;
    const int src = (lane + 31) & 31;
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        #pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) x = __shfl_sync(FULL, x, src);                       // SHFL.IDX chain
            if (MODE == 1) x = __shfl_up_sync(FULL, x, 1);                      // SHFL.UP chain
            if (MODE == 2) x = __vmaxs2(x, y) + 0;                              // VIMNMX.S16x2 chain
            if (MODE == 3) x = __vimax3_s16x2(x, y, z);                         // VIMNMX3 chain
            if (MODE == 4) x = __viaddmax_s16x2(x, y, z);                       // VIADDMNMX chain
            if (MODE == 5) x = prmt(x, y, z);                                   // PRMT chain
            if (MODE == 6) { x = __shfl_sync(FULL, x, src); x = prmt(x, y, 0x3210); x = __vmaxs2(x, z); }  // the NW chain
            if (MODE == 7) { x = __viaddmax_s32((int)x, (int)y, (int)z); }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// throughput: 8 independent chains per thread
template<int MODE> __global__ void thr_kernel(uint32_t* out, long long* cyc, int iters, uint32_t seed)
{
    const int lane = threadIdx.x & 31;
    uint32_t x[8]; uint32_t y = seed * 3 + 1, z = seed ^ 0x1234;
    for (int k = 0; k < 8; ++k) x[k] = seed + threadIdx.x * 8 + k;
    const int src = (lane + 31) & 31;
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        #pragma unroll
        for (int u = 0; u < 2; ++u)
        #pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) x[k] = __shfl_sync(FULL, x[k], src);
            if (MODE == 2) x[k] = __vmaxs2(x[k], y);
            if (MODE == 3) x[k] = __vimax3_s16x2(x[k], y, z);
            if (MODE == 4) x[k] = __viaddmax_s16x2(x[k], y, z);
            if (MODE == 5) x[k] = prmt(x[k], y, z);
        }
    }
    long long t1 = clock64();
    uint32_t a = 0; for (int k = 0; k < 8; ++k) a ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<typename K> void run(const char* name, K kern, int threads, int iters, int per_iter)
{
    uint32_t* out; long long* cyc; cudaMalloc(&out, 4 * 2048); cudaMalloc(&cyc, 8 * 4);
    kern<<<1, threads>>>(out, cyc, 10, 1); cudaDeviceSynchronize();
    kern<<<1, threads>>>(out, cyc, iters, 1); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s threads=%4d  cycles per op (per warp) = %.2f   [err=%s]\n", name, threads, (double)c / iters / per_iter, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    const int it = 20000;
    run("lat SHFL.IDX", lat_kernel<0>, 32, it, 16);
    run("lat SHFL.UP", lat_kernel<1>, 32, it, 16);
    run("lat VIMNMX.S16x2", lat_kernel<2>, 32, it, 16);
    run("lat VIMNMX3.S16x2", lat_kernel<3>, 32, it, 16);
    run("lat VIADDMNMX.S16x2", lat_kernel<4>, 32, it, 16);
    run("lat PRMT", lat_kernel<5>, 32, it, 16);
    run("lat SHFL+PRMT+VIMNMX (NW chain)", lat_kernel<6>, 32, it, 16);
    run("lat VIADDMNMX s32", lat_kernel<7>, 32, it, 16);
    for (int th : {32, 128, 256, 512, 1024}) {
        run("thr SHFL.IDX (per SM, all warps)", thr_kernel<0>, th, it, 16);
    }
    for (int th : {128, 512, 1024}) {
        run("thr VIMNMX.S16x2", thr_kernel<2>, th, it, 16);
        run("thr VIMNMX3.S16x2", thr_kernel<3>, th, it, 16);
        run("thr VIADDMNMX.S16x2", thr_kernel<4>, th, it, 16);
        run("thr PRMT", thr_kernel<5>, th, it, 16);
    }
    return 0;
}
