// nw_batch.cuh -- batch of independent pairs (BASELINE.json configs[4]): one warp per pair, no HBM boundary traffic.
//
// Same recurrence and the same G = H + i + j change of variable as nw_kernels.cuh (reference arithmetic:
// src/serial/serial.cpp:12-31).  A pair of len2 rows is covered by ceil(len2 / (32*R)) row strips handled one after
// the other by the SAME warp; the boundary row between two strips of a pair lives in a per-warp scratch row
// (L2-resident, updated in place).  For the headline shape (1000 x 1000, R = 32) there is exactly one strip and the
// only global traffic is the 2 x 1000 sequence bytes in and one int32 score out.
#pragma once
#include "nw_kernels.cuh"

namespace nw {

struct BatchParams {
    const uint8_t* S1;      // npairs x len1 bytes (columns)
    const uint8_t* S2;      // npairs x len2 bytes (rows)
    int32_t* scores;        // npairs
    int32_t* scratch;       // total_warps x scratch_pitch ints (only used when nstrips > 1)
    long long npairs;
    long long scratch_pitch;
    int len1, len2;
    int nstrips, pad_top;
    int generic;
    int w_match, w_mis, gap;   // G-form weights and the gap score (nw_kernels.cuh); default 3, 2, -1
    uint8_t code[256];      // 4-letter path: byte value -> 0..3
};

template <bool GENERIC>
__device__ __forceinline__ uint32_t batch_col_operand(const uint8_t* s1, int c, int len1, const uint8_t* lut, int wm, int wx)
{
    if (c < 0 || c >= len1) return GENERIC ? 0x100u : (uint32_t)wx * 0x01010101u;
    const uint32_t v = s1[c];
    return GENERIC ? v : weight_word(lut[v], wm, wx);
}

template <int R, bool GENERIC>
__global__ void __launch_bounds__(256) nw_batch_kernel(const BatchParams p)
{
    extern __shared__ uint32_t nw_smem[];
    __shared__ uint8_t lut[256];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) lut[x] = p.code[x];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t* W = nw_smem + warp * SMEM_WORDS_PER_WARP;
    int* sin = (int*)(W + 64);
    int* sout = sin + 32;
    const uint32_t* Wl = W + 32 - lane;
    const long long gw = (long long)blockIdx.x * nwarps + warp, nw_total = (long long)gridDim.x * nwarps;
    int32_t* scratch = p.scratch + gw * p.scratch_pitch;
    const int ncols = p.len1;
    const int nblocks = (ncols + 62) >> 5;
    int32_t* const nofull_t[1] = {nullptr};
    const int nofull_h[1] = {0};

    for (long long pair = gw; pair < p.npairs; pair += nw_total) {
        const uint8_t* s1 = p.S1 + pair * p.len1;
        const uint8_t* s2 = p.S2 + pair * p.len2;
        int h[R];
        for (int s = 0; s < p.nstrips; ++s) {
            RowOperands<R, GENERIC> ro;
            const int k0 = s * 32 * R + lane * R - p.pad_top;      // index into s2 of this lane's first row
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int k = k0 + r;
                uint32_t v;
                if (k >= 0) v = GENERIC ? (uint32_t)s2[k] : (0x5550u | lut[s2[k]]);
                else v = GENERIC ? 0x200u : 0xCCCCu;
                ro.sel[r] = v;
                if (GENERIC) ro.wx[GENERIC ? r : 0] = (v == 0x200u) ? -1 : p.w_mis;
            }
            ro.wm = p.w_match;
            int dprev = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) h[r] = 0;

            W[lane] = GENERIC ? 0x100u : (uint32_t)p.w_mis * 0x01010101u;
            W[lane + 32] = batch_col_operand<GENERIC>(s1, lane, ncols, lut, p.w_match, p.w_mis);
            uint32_t wnext = batch_col_operand<GENERIC>(s1, 32 + lane, ncols, lut, p.w_match, p.w_mis);
            int pre = 0;
            if (s > 0 && lane < ncols) pre = scratch[lane];
            for (int b = 0; b < nblocks; ++b) {
                const int cb = b << 5;
                if (b > 0) {
                    W[lane] = W[lane + 32];
                    W[lane + 32] = wnext;
                    wnext = batch_col_operand<GENERIC>(s1, cb + 32 + lane, ncols, lut, p.w_match, p.w_mis);
                }
                sin[lane] = pre;
                if (s > 0 && cb + 32 + lane < ncols) pre = scratch[cb + 32 + lane];
                __syncwarp();
                if (cb >= 31 && cb + 31 < ncols)
                    sweep32<R, GENERIC, false, false>(h, dprev, ro, Wl, sin, sout, lane, cb, ncols, nofull_t, nofull_h);
                else
                    sweep32<R, GENERIC, false, true>(h, dprev, ro, Wl, sin, sout, lane, cb, ncols, nofull_t, nofull_h);
                __syncwarp();
                if (s + 1 < p.nstrips) {
                    const int oc = cb - 31 + lane;
                    if (oc >= 0 && oc < ncols) scratch[oc] = sout[lane];
                }
            }
            __syncwarp();
        }
        // lane 31 holds G[len2][len1] (or 0 when there is no interior); H = G - i - j
        if (lane == 31) p.scores[pair] = ((ncols > 0 && p.nstrips > 0) ? h[R - 1] : 0) + p.gap * (p.len1 + p.len2);
    }
}

}  // namespace nw

// ---- packed (s16x2) batch kernel ----------------------------------------------------------------------------------------
// One warp per pair, the 64-virtual-lane sweep of nw_packed.cuh (two cells per register), no re-basing: the host only
// selects this kernel when max weight * min(len1, len2) + 64 fits 15 bits, so absolute G values fit the 16-bit lanes.
// Rows beyond one strip (64*R rows) are handled strip after strip by the same warp through a per-warp scratch row.
#include "nw_packed.cuh"

namespace nw {

// profile word of column c for the packed batch kernel: all-zero (weight 0) before the first column
__device__ __forceinline__ uint32_t batch16_col_operand(const uint8_t* s1, int c, int len1, const uint8_t* lut, int wm, int wx)
{
    if (c >= len1) return (uint32_t)wx * 0x01010101u;
    return weight_word(lut[s1[c]], wm, wx);
}

template <int R>
__global__ void __launch_bounds__(128) nw_batch16_kernel(const BatchParams p)
{
    extern __shared__ __align__(16) uint32_t nw_smem[];
    __shared__ uint8_t lut[256];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) lut[x] = p.code[x];
    __syncthreads();

    constexpr int SH = 64 * R;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t* ring = nw_smem + warp * SMEM16_WORDS_PER_WARP;
    uint32_t* sin = ring + 4 * RING_COPY_WORDS;
    uint32_t* sout = sin + 32;
    const uint32_t* ringm = ring + (lane & 3) * RING_COPY_WORDS;
    const long long gw = (long long)blockIdx.x * nwarps + warp, nw_total = (long long)gridDim.x * nwarps;
    int32_t* scratch = p.scratch + gw * p.scratch_pitch;
    const int ncols = p.len1;
    const int nblocks = (ncols + 63 + 31) >> 5;
    const uint32_t upsel = (lane == 0) ? 0x1054u : 0x3210u;
    const int src_lane = (lane + 31) & 31;

    for (long long pair = gw; pair < p.npairs; pair += nw_total) {
        const uint8_t* s1 = p.S1 + pair * p.len1;
        const uint8_t* s2 = p.S2 + pair * p.len2;
        uint32_t h[R];
        for (int s = 0; s < p.nstrips; ++s) {
            uint32_t sel[R];
            const int klo0 = s * SH + lane * R - p.pad_top, khi0 = klo0 + 32 * R;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int klo = klo0 + r, khi = khi0 + r;
                const uint32_t n0 = (klo >= 0) ? lut[s2[klo]] : 0x8u;
                const uint32_t n2 = (khi >= 0) ? 4u + lut[s2[khi]] : 0xCu;
                sel[r] = n0 | 0x80u | (n2 << 8) | 0xC000u;
            }
            uint32_t dprev = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) h[r] = 0;
            uint32_t wnext = batch16_col_operand(s1, lane, ncols, lut, p.w_match, p.w_mis);
            int pre = 0;
            if (s > 0 && lane < ncols) pre = scratch[lane];
            uint32_t scar = __shfl_sync(FULL_MASK, h[R - 1], src_lane);
            // No predicated edge blocks here: a whole table has G = 0 on its top row and left column, and columns before
            // the first one carry the all-zero profile word (weight 0), so a half that has not started yet just keeps
            // computing zeros; past the last column the lanes compute values nobody reads (the strip's bottom row and
            // the score are taken from lane 31's per-step outputs at the right step).
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                ring[m * RING_COPY_WORDS + ((lane + 32 + m) & 127)] = 0u;       // columns [-96, 0) of every copy
                ring[m * RING_COPY_WORDS + ((lane + 64 + m) & 127)] = 0u;
                ring[m * RING_COPY_WORDS + ((lane + 96 + m) & 127)] = 0u;
            }
            int score_g = 0;
            for (int b = 0; b < nblocks; ++b) {
                const int cb = b << 5;
#pragma unroll
                for (int m = 0; m < 4; ++m) ring[m * RING_COPY_WORDS + ((cb + lane + m) & 127)] = wnext;
                wnext = batch16_col_operand(s1, cb + 32 + lane, ncols, lut, p.w_match, p.w_mis);
                sin[lane] = (uint32_t)pre & 0xffffu;
                pre = 0;
                if (s > 0 && cb + 32 + lane < ncols) pre = scratch[cb + 32 + lane];
                __syncwarp();
                sweep16<R, false, false>(h, dprev, sel, upsel, src_lane, ringm, sin, sout, lane, cb, ncols, scar);
                __syncwarp();
                const int oc = cb - 63 + lane;           // column whose last-row value lane 31 produced at step k = lane
                if (s + 1 < p.nstrips) {
                    if (oc >= 0 && oc < ncols) scratch[oc] = (int)sout[lane] >> 16;
                } else if (oc == ncols - 1) {
                    score_g = (int)sout[lane] >> 16;     // G[len2][len1]
                }
            }
            __syncwarp();
            if (s + 1 == p.nstrips) {
                // exactly one lane of one block saw column ncols-1
                score_g = __reduce_max_sync(FULL_MASK, score_g);
                if (lane == 0) p.scores[pair] = ((ncols > 0) ? score_g : 0) + p.gap * (p.len1 + p.len2);
            }
        }
        if (p.nstrips == 0 && lane == 0) p.scores[pair] = p.gap * (p.len1 + p.len2);
    }
}

}  // namespace nw
