// Write-pattern ceiling: (A) contiguous fill, (B) per-warp 128-byte row segments over 256 rows (the pass-2 pattern),
// (C) same with 32-byte-aligned segments.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void fill_contig(int4* p, size_t n16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = make_int4(1,2,3,4);
}
// table rows x pitch ints; warp task = (strip of 256 rows, tile of TB blocks); per block: 256 rows x 32 ints, row segment start = cb - L (+mis)
__global__ void fill_rows(int* t, long long pitch, int nrows, int ncols, int tile_blocks, int aligned) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nstrips = nrows / 256, nblocks = ncols / 32, ntiles = (nblocks + tile_blocks - 1) / tile_blocks;
    const long long ntasks = (long long)nstrips * ntiles;
    for (long long task = (long long)blockIdx.x * nw + warp; task < ntasks; task += (long long)gridDim.x * nw) {
        const int s = task / ntiles, m = task % ntiles;
        for (int b = m * tile_blocks; b < min(nblocks, (m + 1) * tile_blocks); ++b) {
            const int cb = b * 32;
            if (cb < 64) continue;
            for (int L = 0; L < 32; ++L)
                for (int r = 0; r < 4; ++r) {
                    int* lo = t + (long long)(s * 256 + L * 4 + r) * pitch + cb + lane - (aligned ? 0 : L);
                    int* hi = t + (long long)(s * 256 + 128 + L * 4 + r) * pitch + cb + lane - 32 - (aligned ? 0 : L);
                    *lo = cb + L; *hi = cb - L;
                }
        }
    }
}
int main() {
    const int nrows = 22016, ncols = 22528; long long pitch = 22542;
    int* t; cudaMalloc(&t, (size_t)(nrows + 1) * (pitch + 8) * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto f, double bytes) {
        f(); cudaDeviceSynchronize(); cudaEventRecord(e0); for (int i = 0; i < 3; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
        printf("%-46s %.3f ms  %.0f GB/s  [%s]\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    const double bytes = (double)nrows * ncols * 4;
    timeit("contiguous int4 fill", [&] { fill_contig<<<148 * 8, 256>>>((int4*)t, (size_t)(bytes / 16)); }, bytes);
    for (int tb : {2, 32}) for (int w : {8, 16, 32}) for (int al : {0, 1}) {
        char nm[96]; snprintf(nm, 96, "row segments tile_blocks=%d warps/SM=%d pitch=%lld %s", tb, w, pitch, al ? "no-skew" : "skewed");
        timeit(nm, [&] { fill_rows<<<148 * (w / 8), 256>>>(t, pitch, nrows, ncols, tb, al); }, bytes);
    }
    pitch = 22544;
    timeit("row segments tb=2 w=16 pitch=22544 no-skew", [&] { fill_rows<<<148 * 2, 256>>>(t, pitch, nrows, ncols, 2, 1); }, bytes);
    return 0;
}
