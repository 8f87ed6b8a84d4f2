// nw_ws.cuh -- the boundary-mode strip sweep, warp-specialised: every strip is worked on by a PAIR of warps on the same
// scheduler.
//
// Same arithmetic, packing, lag-2 schedule and data layout as nw_lag2.cuh (reference recurrence: src/serial/serial.cpp:
// 12-31).  nw_lag2.cuh interleaves the block loop's memory work (profile ring refill, top-row fetch + tag check +
// conversion, bottom-row publication) with the sweep of the same warp; with one warp per scheduler every one of those
// instructions and every latency they expose is paid by the sweep (measured: 43 cycles per step against 35 for the bare
// sweep).  Here
//   * the COMPUTE warp runs nothing but the 32-step sweeps: integer-pipe instructions, one SHFL and one predicated STS per
//     step, one LDS.128 pair per four steps, and ONE named barrier per block;
//   * its HELPER warp (same scheduler, otherwise idle issue slots) does all global-memory work one block ahead: it polls
//     the tagged words of the strip above, converts them to the compute warp's 16-bit stored form, refills the selector
//     ring three blocks ahead, and after the barrier turns the block's bottom row into tagged words for the strip below.
// The pair shares a 3 KB slice of shared memory; `bar.sync id, 64` at the end of every block hands sin[(b+1)&1] and the
// ring to the compute warp and sout[b&1] to the helper.  The base of the compute warp's 16-bit window travels through
// ctrl[] (written before a barrier, read after it).
#pragma once
#include "nw_lag2.cuh"

namespace nw {

#ifndef NW_WS_POLL_NS
#define NW_WS_POLL_NS 20
#endif
#ifndef NW_WS_SLACK_NS
#define NW_WS_SLACK_NS 1000
#endif
constexpr int WS_CTRL_WORDS = 8;
constexpr int WS_SMEM_WORDS_PER_PAIR = 2 * L2_COPY_WORDS + 64 + 64 + WS_CTRL_WORDS;   // ring copies, 2 x sin, 2 x sout, ctrl

__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// ---- compute warp ---------------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void ws_compute_strip(const StripParams& p, const int s, const int lane, uint32_t* smem, const int bar)
{
    constexpr int SH = 64 * R;
    uint32_t* ring = smem;
    uint32_t* sin = smem + 2 * L2_COPY_WORDS;
    uint32_t* sout = sin + 64;
    volatile int* ctrl = reinterpret_cast<volatile int*>(sout + 64);
    const uint32_t* ringm = ring + (lane & 1) * L2_COPY_WORDS;
    const int ncols = p.ncols;
    const int i_lo = s * SH + lane * R - p.pad_top;   // table row just above the low half's first row (may be <= 0)
    const int i_hi = i_lo + 32 * R;

    uint32_t rowa[R], rowb[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint2 ab = reinterpret_cast<const uint2*>(p.rsel)[(s * 32 + lane) * R + r];
        rowa[r] = ab.x;
        rowb[r] = ab.y;
    }
    const uint32_t upsel = (lane == 0) ? 0x1054u : 0x3210u;
    const int src_lane = (lane + 31) & 31;

    // left boundary column.  Whole table: G = 0.  Column strip: the neighbour's right column (absolute G); the warp's
    // base starts at its minimum so that the stored values are small.
    int base = 0;
    uint32_t h[R];
    uint32_t dprev = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) h[r] = 0;
    if (p.halo != nullptr) {
        wait_halo_politely(p, (s + 1) * SH - p.pad_top);
        int lo[R + 1], hi[R + 1];
        int mn = 0x7fffffff;
#pragma unroll
        for (int r = -1; r < R; ++r) {
            const int a = i_lo + 1 + r, b = i_hi + 1 + r;
            lo[r + 1] = (a >= 1) ? poll_tagged(p, p.halo + a, p.epoch, p.halo_sys).y : 0;
            hi[r + 1] = (b >= 1) ? poll_tagged(p, p.halo + b, p.epoch, p.halo_sys).y : 0;
            mn = min(mn, min(lo[r + 1], hi[r + 1]));
        }
        base = max(__reduce_min_sync(FULL_MASK, mn) - p.margin, 0);
        dprev = ((uint32_t)(lo[0] - base) & 0xffffu) | ((uint32_t)(hi[0] - base) << 16);
#pragma unroll
        for (int r = 0; r < R; ++r) h[r] = ((uint32_t)(lo[r + 1] - base) & 0xffffu) | ((uint32_t)(hi[r + 1] - base) << 16);
    }
    int2* tout = p.brow + (long long)s * p.pitch;

    // (P) the helper learns the base of block 0; (Q) the helper has prepared block 0
    if (lane == 0) ctrl[0] = base;
    pair_barrier(bar);
    pair_barrier(bar);
    // j = 0, the boundary column -- stored only now, when the helper has this strip's first block ready: the strips of a
    // chain wake up one after the other, and only the next one or two poll closely for their first block
    if (lane == 31) st_tagged_gpu(tout, p.epoch, ((int)h[R - 1] >> 16) + base);
    if (p.times != nullptr && lane == 0) { p.times[4 * s] = global_ns(); p.times[4 * s + 2] = (unsigned long long)clock64(); }

    const int nblocks = (ncols + L2_SKEW + 31) >> 5;   // the high half of lane 31 reaches column ncols-1 at t = ncols+125
    uint32_t q0 = __shfl_sync(FULL_MASK, h[R - 1], src_lane), q1 = q0;
#if NW_L2_DBG & 32
    long long accS = 0, accB = 0;
#endif
    for (int b = 0; b < nblocks; ++b) {
#if NW_L2_DBG & 32
        const long long T0 = clock64();
#endif
        sweep16l2<R>(h, dprev, rowa, rowb, upsel, src_lane, ringm, sin + ((b & 1) << 5), sout + ((b & 1) << 5), lane, b << 5,
                     q0, q1, [] {}, [] {}, [] {});
#if NW_L2_DBG & 32
        const long long T1 = clock64();
#endif
        int D = 0;
        if ((b & 31) == 31) {                            // re-base: keep the stored values small
            uint32_t mm = dprev;
#pragma unroll
            for (int r = 0; r < R; ++r) mm = __vmins2(mm, h[r]);
            int m = min((int)(short)(mm & 0xffffu), (int)mm >> 16);
            m = __reduce_min_sync(FULL_MASK, m);
            D = max(m - p.margin, 0);                         // in-flight shuffles are up to two steps (6) older than h
            const uint32_t Dp = (uint32_t)D * 0x10001u;       // every half is >= D: no borrow between halves
#pragma unroll
            for (int r = 0; r < R; ++r) h[r] -= Dp;
            dprev -= Dp;
            q0 -= Dp;
            q1 -= Dp;
            base += D;
        }
        if (lane == 0) ctrl[(b + 1) & 1] = base;         // the base of block b+1, for the helper
        pair_barrier(bar);                               // sout[b&1] -> helper; sin[(b+1)&1] and the ring -> this warp
        if (D > 0) {
            // the helper converted the next block's top row with the old base: move it to the new one
            uint32_t* sn = sin + (((b & 1) ^ 1) << 5) + lane;
            *sn = (*sn - (uint32_t)D) & 0xffffu;
            __syncwarp();
        }
#if NW_L2_DBG & 32
        const long long T2 = clock64();
        accS += T1 - T0; accB += T2 - T1;
#endif
    }
#if NW_L2_DBG & 32
    if (lane == 0 && (s == 0 || s == 2)) printf("ws strip %d: per block: sweep %lld, barrier+rest %lld cycles\n", s, accS / nblocks, accB / nblocks);
#endif
    if (p.times != nullptr && lane == 0) { p.times[4 * s + 1] = global_ns(); p.times[4 * s + 3] = (unsigned long long)clock64(); }

    // right boundary column of this lane's rows (absolute G): every half is frozen at its value in the last column
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
    pair_barrier(bar);                                   // (Z) the helper has published the last block: smem may be reused
}

// ---- helper warp ----------------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void ws_helper_strip(const StripParams& p, const int s, const int lane, uint32_t* smem, const int bar)
{
    uint32_t* ring = smem;
    uint32_t* sin = smem + 2 * L2_COPY_WORDS;
    uint32_t* sout = sin + 64;
    volatile int* ctrl = reinterpret_cast<volatile int*>(sout + 64);
    const int ncols = p.ncols, clast = ncols - 1;
    const int nblocks = (ncols + L2_SKEW + 31) >> 5;
    const uint32_t* wq = p.wq;
    int2* tout = p.brow + (long long)s * p.pitch;
    const int2* tin = p.brow + (long long)(s - 1) * p.pitch;      // (strip 0 never reads it)

    // selector ring: column c at word c & 255 of copy 0 and (c + 2) & 255 of copy 1, words 0..31 mirrored at 256..287
    auto put_ring = [&](int c, uint32_t v) {
        const int w0 = c & 255, w1 = (c + 2) & 255;
        ring[w0] = v;
        if (w0 < L2_MIRROR) ring[256 + w0] = v;
        ring[L2_COPY_WORDS + w1] = v;
        if (w1 < L2_MIRROR) ring[L2_COPY_WORDS + 256 + w1] = v;
    };
    // top boundary row of block nb in the compute warp's stored form (waits for late words)
    auto prepare_top = [&](int nb, int base) {
        int v = 0;
        if (s > 0) {
            const int c = (nb << 5) + lane;
            const bool need = c <= clast;      // past the last column the row above is frozen, like this strip's rows
            const int2* a = tin + c + 1;
            int2 pre = make_int2(p.epoch, base);
            if (need) pre = ld_tagged_gpu(a);
            SpinGuard sg;
            while (!__all_sync(FULL_MASK, pre.x == p.epoch)) {
                __nanosleep(NW_WS_POLL_NS);
                if (pre.x != p.epoch) pre = ld_tagged_gpu(a);
                if (sg.expired_warp(p)) break;
            }
            v = pre.y;
        }
        sin[((nb & 1) << 5) + lane] = (uint32_t)(v - base) & 0xffffu;
    };
    // bottom row of block pb (finished by lane 31's high half at step k = lane of that block) as tagged words
    auto publish = [&](int pb, int base) {
        const int oc = (pb << 5) - L2_SKEW + lane;
        if ((unsigned)oc < (unsigned)ncols)
            st_tagged_gpu(tout + oc + 1, p.epoch, ((int)sout[((pb & 1) << 5) + lane] >> 16) + base);
    };

    // ring: "both halves virtual" for columns < 0, then the first three blocks
    for (int x = lane; x < 2 * L2_COPY_WORDS; x += 32) ring[x] = 0xCC88u;
    __syncwarp();
    for (int c0 = 0; c0 < 96; c0 += 32) put_ring(c0 + lane, wq[c0 + lane]);
    if (s > 0) {
        // Wait politely: most warps of a long chain wait for milliseconds, and hundreds of warps polling L2 in a tight loop
        // slow down the warps that work.  The strip above stores its boundary word (j = 0) when it starts and publishes
        // its first block ~160 steps later.
        SpinGuard sg;
        while (ld_tagged_gpu(tin).x != p.epoch && !sg.expired(p)) __nanosleep(400);
        __nanosleep(NW_WS_SLACK_NS);
    }
    pair_barrier(bar);                                   // (P)
    int base_cur = ctrl[0], base_prev = base_cur;
    prepare_top(0, base_cur);
    pair_barrier(bar);                                   // (Q)
    for (int b = 0; b < nblocks; ++b) {
        uint32_t wnext = wq[((b + 3) << 5) + lane];      // ring columns of block b+3 (zero-padded array): load early ...
        if (b > 0) publish(b - 1, base_prev);
        if (b + 1 < nblocks) prepare_top(b + 1, base_cur);
        put_ring(((b + 3) << 5) + lane, wnext);          // ... store late
        pair_barrier(bar);                               // end of the compute warp's block b
        base_prev = base_cur;
        base_cur = ctrl[(b + 1) & 1];
    }
    publish(nblocks - 1, base_prev);
    pair_barrier(bar);                                   // (Z)
}

template <int R>
__global__ void __launch_bounds__(512) nw_strip16ws_kernel(const StripParams p)
{
    extern __shared__ __align__(16) uint32_t nw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, npairs = blockDim.x >> 6;
    const bool helper = warp >= npairs;
    const int pair = helper ? warp - npairs : warp;      // warps w and w + npairs share a scheduler when npairs % 4 == 0
    uint32_t* smem = nw_smem + pair * WS_SMEM_WORDS_PER_PAIR;
    const int slot = blockIdx.x * npairs + pair, nslots = gridDim.x * npairs;
    wait_mailbox_free(p);
    for (int s = slot; s < p.nstrips; s += nslots) {
        if (helper) ws_helper_strip<R>(p, s, lane, smem, 1 + pair);
        else ws_compute_strip<R>(p, s, lane, smem, 1 + pair);
    }
}

}  // namespace nw
