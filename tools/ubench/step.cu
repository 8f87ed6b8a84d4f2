// In-situ cost of one column step of the packed sweep (nw_packed.cuh: sweep16<4,false>), one warp alone:
// variants remove one ingredient at a time (results are then meaningless; only the time matters).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define FULL_MASK 0xffffffffu
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s){uint32_t d; asm("prmt.b32 %0,%1,%2,%3;":"=r"(d):"r"(a),"r"(b),"r"(s)); return d;}
constexpr int RING_COPY_WORDS = 136;

template <int R, int V>
__device__ __forceinline__ void sweep(uint32_t (&h)[R], uint32_t& dprev, const uint32_t (&sel)[R], const uint32_t upsel,
                                      const int src_lane, const uint32_t* __restrict__ ringm, const uint32_t* __restrict__ sin,
                                      uint32_t* sout, const int lane, const int cb, uint32_t& scar)
{
    const int i0 = cb - lane + (lane & 3);
    uint4 clo = *reinterpret_cast<const uint4*>(ringm + (i0 & 127));
    uint4 chi = *reinterpret_cast<const uint4*>(ringm + ((i0 - 32) & 127));
    uint4 tin = *reinterpret_cast<const uint4*>(sin);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        const uint32_t cl[4] = {clo.x, clo.y, clo.z, clo.w};
        const uint32_t ch[4] = {chi.x, chi.y, chi.z, chi.w};
        const uint32_t tn[4] = {tin.x, tin.y, tin.z, tin.w};
        if (V != 2 && k4 < 7) {
            clo = *reinterpret_cast<const uint4*>(ringm + ((i0 + 4 * k4 + 4) & 127));
            chi = *reinterpret_cast<const uint4*>(ringm + ((i0 + 4 * k4 + 4 - 32) & 127));
            tin = *reinterpret_cast<const uint4*>(sin + 4 * k4 + 4);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = 4 * k4 + kk;
            const uint32_t s = scar;
            uint32_t t[R];
            if (V != 5) {
                uint32_t diag = dprev;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t w = (V == 3) ? cl[kk] : prmt(cl[kk], ch[kk], sel[r]);
                    t[r] = __viaddmax_s16x2(diag, w, h[r]);
                    diag = h[r];
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) t[r] = h[r];
            }
            const uint32_t up0 = prmt(s, tn[kk], upsel);
            dprev = up0;
            uint32_t P[R];
            P[0] = t[0];
            if (V != 5) {
#pragma unroll
                for (int r = 1; r + 1 < R; ++r) P[r] = __vmaxs2(t[r], P[r - 1]);
            } else {
#pragma unroll
                for (int r = 1; r + 1 < R; ++r) P[r] = t[r];
            }
            {
                const uint32_t g = __vimax3_s16x2(t[R - 1], P[R - 2], up0);
                h[R - 1] = g;
                if (V == 4) scar = g + lane; else scar = __shfl_sync(FULL_MASK, h[R - 1], src_lane);
            }
            if (V != 5) {
#pragma unroll
                for (int r = 0; r + 1 < R; ++r) h[r] = (r == 0) ? __vmaxs2(t[0], up0) : __vimax3_s16x2(t[r], P[r - 1], up0);
            }
            if (V != 1 && V != 5) { if (lane == 31) sout[k] = h[R - 1]; }
        }
    }
}

template <int V> __global__ void k(uint32_t* out, long long* cyc, int nblocks)
{
    __shared__ __align__(16) uint32_t smem[4 * RING_COPY_WORDS + 64];
    const int lane = threadIdx.x;
    for (int i = lane; i < 4 * RING_COPY_WORDS + 64; i += 32) smem[i] = 0x02020202u + ((i * 2654435761u >> 13) & 0x01010101u);
    __syncwarp();
    uint32_t h[4] = {1u, 2u, 3u, 4u}, sel[4], dprev = 0, scar = lane;
    for (int r = 0; r < 4; ++r) sel[r] = 0xC080u | ((lane + r) & 3) | ((4u + ((lane * 3 + r) & 3)) << 8);
    const uint32_t upsel = lane == 0 ? 0x1054u : 0x3210u;
    const uint32_t* ringm = smem + (lane & 3) * RING_COPY_WORDS;
    uint32_t* sin = smem + 4 * RING_COPY_WORDS; uint32_t* sout = sin + 32;
    long long t0 = clock64();
    for (int b = 0; b < nblocks; ++b) sweep<4, V>(h, dprev, sel, upsel, (lane + 31) & 31, ringm, sin, sout, lane, b << 5, scar);
    long long t1 = clock64();
    out[lane] = h[0] ^ h[1] ^ h[2] ^ h[3] ^ dprev;
    if (lane == 0) cyc[0] = t1 - t0;
}
template <int V> void run(const char* name) {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    const int nb = 4000;
    k<V><<<1, 32>>>(out, cyc, 10); k<V><<<1, 32>>>(out, cyc, nb); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %.1f cycles/step  [%s]\n", name, (double)c / nb / 32, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<0>("V0 full step (4 regs)");
    run<1>("V1 without lane-31 STS");
    run<2>("V2 without operand LDS.128 (stale vectors)");
    run<3>("V3 without the 4 weight PRMTs");
    run<4>("V4 without SHFL (chain broken)");
    run<5>("V5 chain only: SHFL + PRMT + VIMNMX3");
}
