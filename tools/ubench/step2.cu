// In-situ cost of one column step of the lag-2 packed sweep (nw_lag2.cuh: sweep16l2<4>), one warp alone and several
// warps per SM: variants remove one ingredient at a time (results are then meaningless; only the time matters).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o step2.e step2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define FULL_MASK 0xffffffffu
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s){uint32_t d; asm("prmt.b32 %0,%1,%2,%3;":"=r"(d):"r"(a),"r"(b),"r"(s)); return d;}
constexpr int COPY = 256 + 32 + 16;
constexpr int WORDS = 2 * COPY + 32 + 64;

// V bits: 1 no lane-31 STS, 2 no operand LDS (stale vectors), 4 no SHFL, 8 no weight PRMTs, 16 STS.128 every 4 steps,
//         32 selector-ring form: w = prmt(A[r], B[r], selc) with one LDS.128 per 4 steps
template <int R, int V>
__device__ __forceinline__ void sweep(uint32_t (&h)[R], uint32_t& dprev, const uint32_t (&sel)[R], const uint32_t (&selb)[R], const uint32_t upsel,
                                      const int src_lane, const uint32_t* __restrict__ ringm, const uint32_t* __restrict__ sin,
                                      uint32_t* sout, const int lane, const int cb, uint32_t& q0, uint32_t& q1)
{
    const int i0 = cb - 2 * lane + 2 * (lane & 1);
    const uint32_t* const wlo = ringm + (i0 & 255);
    const uint32_t* const whi = ringm + ((i0 - 64) & 255);
    uint4 clo = *reinterpret_cast<const uint4*>(wlo);
    uint4 chi = *reinterpret_cast<const uint4*>(whi);
    uint4 tin = *reinterpret_cast<const uint4*>(sin);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        const uint32_t cl[4] = {clo.x, clo.y, clo.z, clo.w};
        const uint32_t ch[4] = {chi.x, chi.y, chi.z, chi.w};
        const uint32_t tn[4] = {tin.x, tin.y, tin.z, tin.w};
        if (!(V & 2) && k4 < 7) {
            clo = *reinterpret_cast<const uint4*>(wlo + 4 * k4 + 4);
            if (!(V & 32)) chi = *reinterpret_cast<const uint4*>(whi + 4 * k4 + 4);
            tin = *reinterpret_cast<const uint4*>(sin + 4 * k4 + 4);
        }
        uint4 o;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = 4 * k4 + kk;
            const uint32_t up0 = prmt(q0, tn[kk], upsel);
            uint32_t t[R];
            {
                uint32_t diag = dprev;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    uint32_t w;
                    if (V & 8) w = cl[kk];
                    else if (V & 32) w = prmt(sel[r], selb[r], cl[kk]);
                    else w = prmt(cl[kk], ch[kk], sel[r]);
                    t[r] = __viaddmax_s16x2(diag, w, h[r]);
                    diag = h[r];
                }
            }
            dprev = up0;
            uint32_t g = up0;
#pragma unroll
            for (int r = 0; r < R; r += 2) {
                const uint32_t ga = __vmaxs2(t[r], g);
                h[r] = ga;
                g = __vimax3_s16x2(t[r + 1], t[r], g);
                h[r + 1] = g;
            }
            q0 = q1;
            if (V & 4) q1 = h[R - 1] + lane; else q1 = __shfl_sync(FULL_MASK, h[R - 1], src_lane);
            if (V & 16) {
                if (kk == 0) o.x = h[R - 1]; else if (kk == 1) o.y = h[R - 1]; else if (kk == 2) o.z = h[R - 1]; else o.w = h[R - 1];
                if (kk == 3 && lane == 31) *reinterpret_cast<uint4*>(sout + 4 * k4) = o;
            } else if (!(V & 1)) { if (lane == 31) sout[k] = h[R - 1]; }
        }
    }
}

template <int V> __global__ void k(uint32_t* out, long long* cyc, int nblocks)
{
    extern __shared__ __align__(16) uint32_t smem_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* smem = smem_all + warp * WORDS;
    for (int i = lane; i < WORDS; i += 32) smem[i] = 0x02020202u + ((i * 2654435761u >> 13) & 0x01010101u);
    __syncwarp();
    uint32_t h[4] = {1u, 2u, 3u, 4u}, sel[4], selb[4], dprev = 0, q0 = lane, q1 = lane + 1;
    for (int r = 0; r < 4; ++r) { sel[r] = 0xC080u | ((lane + r) & 3) | ((4u + ((lane * 3 + r) & 3)) << 8); selb[r] = sel[r] * 3; }
    const uint32_t upsel = lane == 0 ? 0x1054u : 0x3210u;
    const uint32_t* ringm = smem + (lane & 1) * COPY;
    uint32_t* sin = smem + 2 * COPY; uint32_t* sout = sin + 32;
    __syncthreads();
    long long t0 = clock64();
    for (int b = 0; b < nblocks; ++b) sweep<4, V>(h, dprev, sel, selb, upsel, (lane + 31) & 31, ringm, sin, sout + ((b & 1) << 5), lane, b << 5, q0, q1);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = h[0] ^ h[1] ^ h[2] ^ h[3] ^ dprev;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int V> void run(const char* name) {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int nb = 4000;
    printf("%-56s", name);
    for (int warps : {1, 4, 8, 16}) {
        cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * WORDS * 4);
        k<V><<<1, 32 * warps, warps * WORDS * 4>>>(out, cyc, 10); k<V><<<1, 32 * warps, warps * WORDS * 4>>>(out, cyc, nb); cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf(" %2dw: %6.1f", warps, (double)c / nb / 32);
    }
    printf("   cycles/step  [%s]\n", cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<0>("V0 full step (4 regs, lag 2)");
    run<1>("V1 without lane-31 STS");
    run<2>("V2 without operand LDS.128 (stale vectors)");
    run<4>("V4 without SHFL");
    run<8>("V8 without the 4 weight PRMTs");
    run<7>("V7 without STS, LDS, SHFL: ALU only");
    run<15>("V15 without STS, LDS, SHFL, weight PRMTs");
    run<16>("V16 one STS.128 per 4 steps");
    run<32>("V32 selector ring: 1 column LDS.128 per 4 steps");
    run<48>("V48 selector ring + STS.128");
}
