// nw_packed2.cuh -- boundary-mode packed strip kernel, TWO columns per lane per step.
//
// Same arithmetic and data layout as nw_packed.cuh (packed s16x2 lanes, 64 virtual lanes per warp, tagged boundary rows,
// per-warp re-basing; reference recurrence: src/serial/serial.cpp:12-31).  The difference is the schedule: a single
// pair is bound by the wavefront's critical path, and with one column per step that path pays one shuffle latency
// (~26 cycles, profiles/r01_ubench_latency.log) per column.  Here virtual lane v works on columns 2(t-v) and 2(t-v)+1 at
// step t, the two last-row values of a step travel in two back-to-back shuffles, and the shuffle latency is paid once
// per two columns.  The price is twice the skew between virtual lanes (a strip starts ~190 columns after its
// predecessor instead of ~135), which is worth it when the table is much wider than (number of strips) x (skew).
#pragma once
#include "nw_packed.cuh"

namespace nw {

constexpr int K2_RING_PITCH = 256 + 16;                         // 256-column ring per copy, 16 words of bank skew
constexpr int SMEM16K2_WORDS_PER_WARP = 2 * K2_RING_PITCH + 64 + 64;   // two ring copies + 64 top inputs + 64 outputs

template <int R, bool PRED>
__device__ __forceinline__ void sweep16k2(uint32_t (&h)[R], uint32_t& dprev, const uint32_t (&sel)[R],
                                          const uint32_t upsel, const int src_lane, const uint32_t* __restrict__ ringm,
                                          const uint32_t* __restrict__ sin, uint32_t* sout, const int lane, const int cb2,
                                          const int ncols, uint32_t& scarA, uint32_t& scarB)
{
    const int i0 = cb2 - 2 * lane + 2 * (lane & 1);      // ring index (before & 255) of column cA at k = 0; 4 | i0
    uint4 clo = *reinterpret_cast<const uint4*>(ringm + (i0 & 255));
    uint4 chi = *reinterpret_cast<const uint4*>(ringm + ((i0 - 64) & 255));
    uint4 tin = *reinterpret_cast<const uint4*>(sin);
#pragma unroll
    for (int g = 0; g < 16; ++g) {                       // two steps = four columns per operand vector
        const uint32_t cl[4] = {clo.x, clo.y, clo.z, clo.w};
        const uint32_t ch[4] = {chi.x, chi.y, chi.z, chi.w};
        const uint32_t tn[4] = {tin.x, tin.y, tin.z, tin.w};
        if (g < 15) {
            clo = *reinterpret_cast<const uint4*>(ringm + ((i0 + 4 * g + 4) & 255));
            chi = *reinterpret_cast<const uint4*>(ringm + ((i0 + 4 * g + 4 - 64) & 255));
            tin = *reinterpret_cast<const uint4*>(sin + 4 * g + 4);
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const int k = 2 * g + kk;
            const uint32_t sA = scarA, sB = scarB;       // lane-1's last row after its columns A and B of the previous step
            uint32_t maskA = 0xffffffffu, maskB = 0xffffffffu;
            if (PRED) {
                const int cA = cb2 + 2 * k - 2 * lane;
                maskA = ((unsigned)cA < (unsigned)ncols ? 0x0000ffffu : 0u) |
                        ((unsigned)(cA - 64) < (unsigned)ncols ? 0xffff0000u : 0u);
                maskB = ((unsigned)(cA + 1) < (unsigned)ncols ? 0x0000ffffu : 0u) |
                        ((unsigned)(cA - 63) < (unsigned)ncols ? 0xffff0000u : 0u);
            }
            // ---- column A ------------------------------------------------------------------------------------------------
            uint32_t t[R], P[R], gA[R];
            {
                uint32_t diag = dprev;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t w = prmt(cl[2 * kk], ch[2 * kk], sel[r]);
                    t[r] = __viaddmax_s16x2(diag, w, h[r]);
                    diag = h[r];
                }
            }
            P[0] = t[0];
#pragma unroll
            for (int r = 1; r + 1 < R; ++r) P[r] = __vmaxs2(t[r], P[r - 1]);
            const uint32_t upA = prmt(sA, tn[2 * kk], upsel);
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                uint32_t gg;
                if (r == 0) gg = __vmaxs2(t[0], upA);
                else gg = __vimax3_s16x2(t[r], P[r - 1], upA);
                gA[r] = PRED ? ((gg & maskA) | (h[r] & ~maskA)) : gg;
                if (r == R - 1) scarA = __shfl_sync(FULL_MASK, gA[R - 1], src_lane);
            }
            if (lane == 31) sout[2 * k] = gA[R - 1];
            // ---- column B (left neighbour = column A of this step) -----------------------------------------------------------
            {
                uint32_t diag = upA;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t w = prmt(cl[2 * kk + 1], ch[2 * kk + 1], sel[r]);
                    t[r] = __viaddmax_s16x2(diag, w, gA[r]);
                    diag = gA[r];
                }
            }
            P[0] = t[0];
#pragma unroll
            for (int r = 1; r + 1 < R; ++r) P[r] = __vmaxs2(t[r], P[r - 1]);
            const uint32_t upB = prmt(sB, tn[2 * kk + 1], upsel);
            dprev = upB;
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                uint32_t gg;
                if (r == 0) gg = __vmaxs2(t[0], upB);
                else gg = __vimax3_s16x2(t[r], P[r - 1], upB);
                h[r] = PRED ? ((gg & maskB) | (gA[r] & ~maskB)) : gg;
                if (r == R - 1) scarB = __shfl_sync(FULL_MASK, h[R - 1], src_lane);
            }
            if (lane == 31) sout[2 * k + 1] = h[R - 1];
        }
    }
}

template <int R>
__device__ __forceinline__ void run_strip16k2(const StripParams& p, const int s, const int lane, uint32_t* smem)
{
    constexpr int SH = 64 * R;
    uint32_t* ring = smem;
    uint32_t* sin = smem + 2 * K2_RING_PITCH;
    uint32_t* sout = sin + 64;
    const uint32_t* ringm = ring + (lane & 1) * K2_RING_PITCH;
    const int ncols = p.ncols;
    const int i_lo = s * SH + lane * R - p.pad_top;   // table row just above the low half's first row (may be <= 0)
    const int i_hi = i_lo + 32 * R;

    uint32_t sel[R];
#pragma unroll
    for (int r = 0; r < R; ++r) sel[r] = p.rsel[(s * 32 + lane) * R + r];
    const uint32_t upsel = (lane == 0) ? 0x1054u : 0x3210u;
    const int src_lane = (lane + 31) & 31;

    int base = 0;
    uint32_t h[R];
    uint32_t dprev = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) h[r] = 0;
    if (p.halo != nullptr) {
        int lo[R + 1], hi[R + 1];
        int mn = 0x7fffffff;
#pragma unroll
        for (int r = -1; r < R; ++r) {
            const int a = i_lo + 1 + r, b = i_hi + 1 + r;
            lo[r + 1] = (a >= 1) ? poll_tagged(p.halo + a, p.epoch, p.halo_sys).y : 0;
            hi[r + 1] = (b >= 1) ? poll_tagged(p.halo + b, p.epoch, p.halo_sys).y : 0;
            mn = min(mn, min(lo[r + 1], hi[r + 1]));
        }
        base = max(__reduce_min_sync(FULL_MASK, mn) - 8, 0);
        dprev = ((uint32_t)(lo[0] - base) & 0xffffu) | ((uint32_t)(hi[0] - base) << 16);
#pragma unroll
        for (int r = 0; r < R; ++r) h[r] = ((uint32_t)(lo[r + 1] - base) & 0xffffu) | ((uint32_t)(hi[r + 1] - base) << 16);
    }

    int2* tout = p.brow + (long long)s * p.pitch;
    const int2* tin = p.brow + (long long)(s - 1) * p.pitch;
    if (lane == 31) st_tagged_gpu(tout, p.epoch, ((int)h[R - 1] >> 16) + base);      // j = 0: the boundary column

    const uint32_t* wq = p.wq;
    const int wmax = ncols + WQ_PAD - 1;
    uint32_t wn0 = wq[min(lane, wmax)], wn1 = wq[min(32 + lane, wmax)];
    int2 pre0 = make_int2(0, 0), pre1 = make_int2(0, 0);
    if (s > 0) {
        if (lane < ncols) pre0 = ld_tagged_gpu(tin + lane + 1);
        if (32 + lane < ncols) pre1 = ld_tagged_gpu(tin + 32 + lane + 1);
    }

    // the high half of lane 31 finishes column ncols-1 as its A or B column at step t with 2t - 126 <= ncols-1 <= 2t - 125
    const int nsteps = ((ncols + 125) >> 1) + 1;
    const int nblocks = (nsteps + 31) >> 5;
    uint32_t scarA = __shfl_sync(FULL_MASK, h[R - 1], src_lane), scarB = scarA;
    for (int b = 0; b < nblocks; ++b) {
        const int cb2 = b << 6;
        // column operands of [cb2, cb2+64) into the two skewed ring copies; the ring then holds [cb2-192, cb2+64)
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            ring[m * K2_RING_PITCH + ((cb2 + lane + 2 * m) & 255)] = wn0;
            ring[m * K2_RING_PITCH + ((cb2 + 32 + lane + 2 * m) & 255)] = wn1;
        }
        wn0 = wq[min(cb2 + 64 + lane, wmax)];
        wn1 = wq[min(cb2 + 96 + lane, wmax)];
        // top boundary row of [cb2, cb2+64) -> stored form, low half
        if (cb2 < ncols) {
            int v0 = 0, v1 = 0;
            if (s > 0) {
                const int c0 = cb2 + lane, c1 = c0 + 32;
                const bool need0 = c0 < ncols, need1 = c1 < ncols;
                while (!__all_sync(FULL_MASK, (!need0 || pre0.x == p.epoch) && (!need1 || pre1.x == p.epoch))) {
                    if (need0 && pre0.x != p.epoch) pre0 = ld_tagged_gpu(tin + c0 + 1);
                    if (need1 && pre1.x != p.epoch) pre1 = ld_tagged_gpu(tin + c1 + 1);
                }
                v0 = pre0.y;
                v1 = pre1.y;
                if (c0 + 64 < ncols) pre0 = ld_tagged_gpu(tin + c0 + 65);
                if (c1 + 64 < ncols) pre1 = ld_tagged_gpu(tin + c1 + 65);
            }
            sin[lane] = (uint32_t)(v0 - base) & 0xffffu;
            sin[32 + lane] = (uint32_t)(v1 - base) & 0xffffu;
        }
        __syncwarp();
        if (cb2 >= 128 && cb2 + 63 < ncols)
            sweep16k2<R, false>(h, dprev, sel, upsel, src_lane, ringm, sin, sout, lane, cb2, ncols, scarA, scarB);
        else
            sweep16k2<R, true>(h, dprev, sel, upsel, src_lane, ringm, sin, sout, lane, cb2, ncols, scarA, scarB);
        __syncwarp();
        {
            // sout[j] holds column cb2 - 126 + j of the strip's last row (high half of lane 31)
            const int oc0 = cb2 - 126 + lane, oc1 = oc0 + 32;
            if (oc0 >= 0 && oc0 < ncols) st_tagged_gpu(tout + oc0 + 1, p.epoch, ((int)sout[lane] >> 16) + base);
            if (oc1 >= 0 && oc1 < ncols) st_tagged_gpu(tout + oc1 + 1, p.epoch, ((int)sout[32 + lane] >> 16) + base);
        }
        if ((b & 15) == 15) {                            // re-base: keep the stored values small
            uint32_t mm = dprev;
#pragma unroll
            for (int r = 0; r < R; ++r) mm = __vmins2(mm, h[r]);
            int m = min((int)(short)(mm & 0xffffu), (int)mm >> 16);
            m = __reduce_min_sync(FULL_MASK, m);
            const int D = m - 8;
            if (D > 0) {
                const uint32_t Dp = (uint32_t)D * 0x10001u;       // every half is >= D: no borrow between halves
#pragma unroll
                for (int r = 0; r < R; ++r) h[r] -= Dp;
                dprev -= Dp;
                scarA -= Dp;
                scarB -= Dp;
                base += D;
            }
        }
    }

    if (p.rcol != nullptr) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int a = i_lo + 1 + r, b = i_hi + 1 + r;
            const int va = (int)(short)(h[r] & 0xffffu) + base, vb = ((int)h[r] >> 16) + base;
            if (p.rcol_sys) {
                if (a >= 1) st_tagged_sys(p.rcol + a, p.epoch, va);
                if (b >= 1) st_tagged_sys(p.rcol + b, p.epoch, vb);
            } else {
                if (a >= 1) st_tagged_gpu(p.rcol + a, p.epoch, va);
                if (b >= 1) st_tagged_gpu(p.rcol + b, p.epoch, vb);
            }
        }
    }
    __syncwarp();
}

template <int R>
__global__ void __launch_bounds__(512) nw_strip16k2_kernel(const StripParams p)
{
    extern __shared__ __align__(16) uint32_t nw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t* smem = nw_smem + warp * SMEM16K2_WORDS_PER_WARP;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    if (p.ack_in != nullptr) {
        if (threadIdx.x == 0) {
            int a;
            do {
                asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(a) : "l"(p.ack_in) : "memory");
                if (a < p.epoch - 2) __nanosleep(500);
            } while (a < p.epoch - 2);
        }
        __syncthreads();
    }
    for (int s = slot; s < p.nstrips; s += nslots) run_strip16k2<R>(p, s, lane, smem);
}

}  // namespace nw
