// nw_lag2.cuh -- the boundary-mode strip sweep, packed s16x2 lanes, virtual lane v TWO columns behind v-1.
//
// Same recurrence, G = H + i + j change of variable, packed 16-bit halves, tagged boundary rows and re-basing as
// nw_packed.cuh (reference arithmetic: src/serial/serial.cpp:12-31).  What changes is the schedule of a warp:
//   * virtual lane v works on column t - 2v at step t (low half of lane L: column t - 2L; high half: t - 2L - 64), so the
//     last row of lane L-1 that lane L needs at step t was produced at step t-2: the shuffle issued two steps earlier is
//     consumed now, and the 24-cycle SHFL latency leaves the per-step dependency chain.  What remains loop-carried is
//     VIADDMNMX -> VIMNMX (9 cycles), so a step costs what its 3R+1 integer-pipe instructions cost (2 cycles each with
//     one warp per scheduler) instead of the 36-cycle SHFL -> PRMT -> VIMNMX3 chain of the one-column skew.
//   * no predicated edge blocks: columns before the first and after the last carry the all-zero profile word (weight
//     0).  G is monotone along rows and columns, so with weight 0 a cell keeps the value of its left neighbour: before
//     column 0 a half keeps its left-boundary value, after the last column it keeps the right-boundary value, which is
//     what the warp hands on (right column) at the end.  Every block runs the same straight-line code.
//   * the operand roles are swapped: w = PRMT(A[r], B[r], S[c]) with two row words per register (weights by letter of
//     the low / high half's row) and ONE selector word per column pair (c, c - 64), so a lane loads one column vector
//     per four steps instead of two (the shared-memory pipe is shared by the SM's four warps and was the next limit:
//     38 -> 31 cycles per step in isolation, tools/ubench/step2.cu);
//   * the column selector ring (256 columns, two bank-skewed copies so that each lane's 4-column window is one aligned
//     LDS.128, plus a 32-word mirror so that a block's window never wraps) is refilled by cp.async (LDGSTS) two blocks ahead;
//   * the top boundary row also arrives by cp.async (.cg: from L2, where the producer's stores land), late in the
//     previous block, into a staging buffer of tagged words.  A register load there would do, but its long-latency
//     scoreboard gets shared with the sweep's SHFL / LDS scoreboards and stalls the sweep for an L2 round trip per block
//     (measured: +7 cycles per step); cp.async completion is tracked by its own group counter instead; the top boundary row is fetched late in the previous
//     block (so a strip trails its predecessor by the hand-off latency, not by a whole block more); the bottom row of
//     block b is published while block b+1 is being computed.
// The price is twice the skew between virtual lanes: a strip starts 126 + 32 columns (plus the hand-off latency) after
// its predecessor.
#pragma once
#include "nw_packed.cuh"
#include <cstdio>
#ifndef NW_L2_DBG
#define NW_L2_DBG 0      // development only: bit mask of block-loop ingredients to leave out (timing experiments)
#endif

namespace nw {

constexpr int L2_MIRROR = 32;                                         // words 256..287 of a copy repeat words 0..31, so a
                                                                      // 32-word window never wraps: LDS with immediates
constexpr int L2_COPY_WORDS = 256 + L2_MIRROR + 16;                   // 256-column ring + mirror + 16 words of bank skew
constexpr int L2_SMEM_WORDS_PER_WARP = 2 * L2_COPY_WORDS + 64 + 64 + 128;   // two ring copies + 2 x top inputs + 2 x bottom
                                                                            // outputs + 2 x 32 staged tagged words
constexpr int L2_SKEW = 126;                                          // columns between virtual lane 0 and virtual lane 63

__device__ __forceinline__ void cp_async8(uint32_t* smem_dst, const uint32_t* gsrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4_if(bool pred, uint32_t* smem_dst, const uint32_t* gsrc)      // predicated, no branch
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p cp.async.ca.shared.global [%0], [%1], 4; }" ::"r"(d), "l"(gsrc),
                 "r"((int)pred)
                 : "memory");
}
__device__ __forceinline__ void cp_async8_if(bool pred, uint32_t* smem_dst, const uint32_t* gsrc)      // predicated, no branch
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p cp.async.ca.shared.global [%0], [%1], 8; }" ::"r"(d), "l"(gsrc),
                 "r"((int)pred)
                 : "memory");
}
__device__ __forceinline__ void cp_async16_cg_if(bool pred, void* smem_dst, const void* gsrc)   // L2 only: coherent with peers' stores
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p cp.async.cg.shared.global [%0], [%1], 16; }" ::"r"(d), "l"(gsrc),
                 "r"((int)pred)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One 32-step block.  q0 / q1: the shuffles issued two steps / one step ago.  hook0 / hook2 / hook7 run before steps 0, 8 and
// 28 of the 32: the block loop's own work (publishing, prefetching, converting the next block's top row), which would
// otherwise sit between two sweeps with its latencies exposed (measured: 350-400 cycles per block with all warps of an SM
// active, against 1120 for the sweep).
template <int R, class H0, class H2, class H7>
__device__ __forceinline__ void sweep16l2(uint32_t (&h)[R], uint32_t& dprev, const uint32_t (&rowa)[R], const uint32_t (&rowb)[R], const uint32_t upsel,
                                          const int src_lane, const uint32_t* __restrict__ ringm,
                                          const uint32_t* __restrict__ sin, uint32_t* sout, const int lane, const int cb,
                                          uint32_t& q0, uint32_t& q1, H0&& hook0, H2&& hook2, H7&& hook7)
{
    const int i0 = cb - 2 * lane + 2 * (lane & 1);      // ring index (before & 255) of this lane's low column at k = 0; 4 | i0
    const uint32_t* const wlo = ringm + (i0 & 255);     // 32-word window (the mirror makes it contiguous)
    uint4 clo = *reinterpret_cast<const uint4*>(wlo);
    uint4 tin = *reinterpret_cast<const uint4*>(sin);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        const uint32_t cl[4] = {clo.x, clo.y, clo.z, clo.w};
        const uint32_t tn[4] = {tin.x, tin.y, tin.z, tin.w};
        if (k4 < 7) {
            clo = *reinterpret_cast<const uint4*>(wlo + 4 * k4 + 4);
            tin = *reinterpret_cast<const uint4*>(sin + 4 * k4 + 4);
        }
        if (k4 == 0) hook0();
        if (k4 == 2) hook2();
        if (k4 == 7) hook7();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = 4 * k4 + kk;
            // {low: row above at column c, high: row above at column c - 64}; lane 0: low from the strip's top boundary
            // row, high from lane 31's low half
            const uint32_t up0 = prmt(q0, tn[kk], upsel);
            uint32_t t[R];
            {
                uint32_t diag = dprev;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t w = prmt(rowa[r], rowb[r], cl[kk]);
                    t[r] = __viaddmax_s16x2(diag, w, h[r]);       // max(G[i-1][j-1] + w, G[i][j-1])
                    diag = h[r];
                }
            }
            dprev = up0;
            uint32_t g = up0;                                     // running max down the rows; every second link is a max3
#pragma unroll
            for (int r = 0; r < R; r += 2) {
                const uint32_t ga = __vmaxs2(t[r], g);
                h[r] = ga;
                if (r + 1 < R) {
                    g = __vimax3_s16x2(t[r + 1], t[r], g);
                    h[r + 1] = g;
                } else {
                    g = ga;
                }
            }
            q0 = q1;
            q1 = __shfl_sync(FULL_MASK, h[R - 1], src_lane);
            if (lane == 31) sout[k] = h[R - 1];
        }
    }
}

// STAIR: score mode along a staircase -- the strip has its own width (StripParams::widths); instantiated separately so
// that the plain kernel's code is exactly what it was.
template <int R, bool STAIR>
__device__ __forceinline__ void run_strip16l2(const StripParams& p, const int s, const int lane, uint32_t* smem)
{
    constexpr int SH = 64 * R;
    uint32_t* ring = smem;
    uint32_t* sin = smem + 2 * L2_COPY_WORDS;          // 2 x 32, by block parity
    uint32_t* sout = sin + 64;                         // 2 x 32, by block parity
    int2* stag = reinterpret_cast<int2*>(sout + 64);   // 2 x 32 tagged words of the top boundary row, by block parity
    const uint32_t* ringm = ring + (lane & 1) * L2_COPY_WORDS;
    // staircase score mode: this strip's own width.  Beyond it the strip is frozen exactly like every strip is beyond the
    // table's last column -- its selector words are virtual (weight 0) from there on
    const int ncols = STAIR ? p.widths[s] : p.ncols;
    const int xcut = (ncols < p.ncols) ? ncols : 0x7fffff00;
    const uint32_t* const tail = STAIR ? p.tails + s * 64 : nullptr;
    const uint32_t* const virt = p.wq + p.ncols + 64;                  // an all-virtual word (the right padding of wq)
    const int q_lo = s * SH + lane * R;               // first padded row of the low half; the high half is 32*R below
    const int i_lo = q_lo - p.pad_top;                // table row just above the low half's first row (may be <= 0)
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

    // selector ring: "both halves virtual" (columns < 0), then columns [0, 64) on their way
    __syncwarp();
    for (int x = lane; x < 2 * L2_COPY_WORDS; x += 32) ring[x] = 0xCC88u;
    __syncwarp();
    const uint32_t* wq = p.wq;
    // lanes 0..15 fill copy 0 (column c at word c & 255), lanes 16..31 copy 1 (column c at word (c + 2) & 255)
    const int fx = 2 * (lane & 15);
    uint32_t* const fdst = ring + (lane >> 4) * L2_COPY_WORDS;
    const int fskew = (lane >> 4) * 2;
    auto fill = [&](int c0) {      // columns [c0, c0 + 32); the array is zero-padded on both sides
        const int w = (c0 + fx + fskew) & 255;
        if (STAIR) {       // word by word: the cut may fall between the two (selects, no branch)
            const int ca = c0 + fx, cb2 = ca + 1;
            if (ca < xcut) chk_wq(p, ca);
            if (cb2 < xcut) chk_wq(p, cb2);
            const uint32_t* sa = (ca >= xcut) ? ((ca < xcut + 64) ? tail + (ca - xcut) : virt) : wq + ca;
            const uint32_t* sb = (cb2 >= xcut) ? ((cb2 < xcut + 64) ? tail + (cb2 - xcut) : virt) : wq + cb2;
            cp_async4_if(true, fdst + w, sa);
            cp_async4_if(true, fdst + w + 1, sb);
            cp_async4_if(w < L2_MIRROR, fdst + 256 + w, sa);
            cp_async4_if(w < L2_MIRROR, fdst + 256 + w + 1, sb);
        } else {
            chk_wq(p, c0 + fx);
            chk_wq(p, c0 + fx + 1);
            cp_async8(fdst + w, wq + c0 + fx);
            cp_async8_if(w < L2_MIRROR, fdst + 256 + w, wq + c0 + fx);
        }
        cp_async_commit();
    };
    fill(0);
    fill(32);

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
            if (a >= 1) chk_row(p, a);
            if (b >= 1) chk_row(p, b);
            lo[r + 1] = (a >= 1) ? poll_tagged(p, p.halo + a, p.epoch, p.halo_sys).y : 0;
            hi[r + 1] = (b >= 1) ? poll_tagged(p, p.halo + b, p.epoch, p.halo_sys).y : 0;
            mn = min(mn, min(lo[r + 1], hi[r + 1]));
        }
        base = max(__reduce_min_sync(FULL_MASK, mn) - p.margin, 0);
        dprev = ((uint32_t)(lo[0] - base) & 0xffffu) | ((uint32_t)(hi[0] - base) << 16);
#pragma unroll
        for (int r = 0; r < R; ++r) h[r] = ((uint32_t)(lo[r + 1] - base) & 0xffffu) | ((uint32_t)(hi[r + 1] - base) << 16);
    }

    // The sweep and the work interleaved with it are branch-free (a branch the compiler cannot prove warp-uniform makes it
    // guard every later shuffle with a convergence check, which cuts the unrolled sweep into small scheduling regions):
    // out-of-range bottom-row words go to the spare words behind the row, strip 0 "prefetches" its own row and ignores it.
    int2* tout = p.brow + (long long)s * p.pitch;
    const int2* tin = (s > 0) ? p.brow + (long long)(s - 1) * p.pitch : tout;
    int2* const dump = tout + ncols + 1 + (lane & 7);
    if (lane == 31) st_tagged_gpu(tout, p.epoch, ((int)h[R - 1] >> 16) + base);      // j = 0: the boundary column

    // top boundary row of block b: tagged word of column min(cb + lane, ncols - 1) (past the last column the row above is
    // frozen at its last value, like this strip's own rows)
    const int clast = ncols - 1;
    const int want = (s > 0) ? p.epoch : 0;          // strip 0: top row = 0, nothing to wait for
    // staged prefetch of the top-row words of columns [c0, c0 + 32): lanes 0..15 move two words each (the word of interior
    // column c sits at index c + 1; c0 + 1 + 2x is odd, so the 16-byte alignment comes from the shifted row base)
    auto fetch_top = [&](int c0, int par) {
        // (strip 0 has no row above: no fetch.  Its "row above" used to be its own row a block ahead of its stores, i.e.
        //  words of an earlier fill that have long left L2 -- one DRAM round trip per block on the critical path.)
        if (s > 0 && lane < 16 && c0 + 2 * lane <= clast) chk_brow(p, s - 1, c0 + 1 + 2 * lane + 1);
        cp_async16_cg_if(s > 0 && lane < 16 && c0 + 2 * lane <= clast, stag + (par << 5) + 2 * lane, tin + c0 + 1 + 2 * lane);
        cp_async_commit();
    };
    fetch_top(0, 0);

    const int nblocks = (ncols + L2_SKEW + 31) >> 5;   // the high half of lane 31 reaches column ncols-1 at t = ncols+125
    uint32_t q0 = __shfl_sync(FULL_MASK, h[R - 1], src_lane), q1 = q0;
    int pub_base = base;
#if NW_L2_DBG & 1024
    long long dbg_miss = 0, dbg_wait = 0;
#endif
    // top boundary row of block nb, from the staged words to stored form in sin[nb & 1] (waits for late words)
    auto convert_top = [&](int nb, bool first) {
        // the staged words have landed; the ring columns fetched in this block (the newest group, needed three blocks
        // from now, and from DRAM when this strip is the first to touch them) may still be on their way
        if (first) cp_async_wait<0>(); else cp_async_wait<1>();
        __syncwarp();                                  // ... for every lane
        int2 pre = stag[((nb & 1) << 5) + lane];
        int v = 0;
        if (s > 0 && !(NW_L2_DBG & 64)) {
            // past the last column the row above is frozen and this strip's rows are too: any small value will do
            const bool need = (nb << 5) + lane <= clast;
            const int2* a = tin + (nb << 5) + lane + 1;
            if (need) chk_brow(p, s - 1, (nb << 5) + lane + 1);
            SpinGuard sg;
#if NW_L2_DBG & 1024
            const long long tw0 = clock64();
            bool missed = false;
#endif
            while (!__all_sync(FULL_MASK, !need || pre.x == want)) {
                __nanosleep(32);
                if (need && pre.x != want) pre = ld_tagged_gpu(a);
                if (sg.expired_warp(p)) break;
#if NW_L2_DBG & 1024
                missed = true;
#endif
            }
#if NW_L2_DBG & 1024
            if (missed && nb > 0) { ++dbg_miss; dbg_wait += clock64() - tw0; }
#endif
            v = need ? pre.y : base;
        }
        sin[((nb & 1) << 5) + lane] = (uint32_t)(v - base) & 0xffffu;
    };
    if (s > 0 && !(NW_L2_DBG & 64)) {
        // Wait politely.  Most warps of a long chain wait for milliseconds; hundreds of warps polling L2 in a tight loop
        // slow down the warps that work (measured: 41 -> 56 cycles per step for every strip).  The predecessor stores its
        // boundary word (j = 0) when it starts and publishes its first block ~160 steps (~3.5 us) later.
        SpinGuard sg;
        while (ld_tagged_gpu(tin).x != want && !sg.expired(p)) __nanosleep(400);
        __nanosleep(1500);
#if NW_L2_DBG & 2048
        while (ld_tagged_gpu(tin + min(NW_L2_SLACK, clast) + 1).x != want) __nanosleep(1000);     // start far behind the write front
#endif
    }
    convert_top(0, true);
    if (p.times != nullptr && lane == 0) { p.times[4 * s] = global_ns(); p.times[4 * s + 2] = (unsigned long long)clock64(); }
    __syncwarp();
    fill(64);
    for (int b = 0; b < nblocks; ++b) {
        const int cb = b << 5;
        uint32_t* so = sout + ((b & 1) << 5);
        const uint32_t* so_prev = sout + (((b & 1) ^ 1) << 5);
        sweep16l2<R>(h, dprev, rowa, rowb, upsel, src_lane, ringm, sin + ((b & 1) << 5), so, lane, cb, q0, q1,
            [&] {       // prefetch the top row of the next block; publish the bottom row of the previous one (the column
                        // finished by lane 31's high half at step k = lane; block 0: out of range for every lane)
                if (!(NW_L2_DBG & 128)) fetch_top(cb + 32, (b & 1) ^ 1);
                const int oc = cb - 32 - L2_SKEW + lane;
                int2* const dst = ((unsigned)oc < (unsigned)ncols) ? tout + oc + 1 : dump;
                chk_brow(p, s, dst - tout);
                if (!(NW_L2_DBG & 256)) st_tagged_gpu(dst, p.epoch, ((int)so_prev[lane] >> 16) + pub_base);
            },
            [&] { if (!(NW_L2_DBG & 512)) fill(cb + 96); },
            [&] { convert_top(b + 1, false); });
        __syncwarp();                                  // sin of the next block is visible; sout of this block is complete
        pub_base = base;
        if ((b & 31) == 31) {                            // re-base: keep the stored values small
            uint32_t mm = dprev;
#pragma unroll
            for (int r = 0; r < R; ++r) mm = __vmins2(mm, h[r]);
            int m = min((int)(short)(mm & 0xffffu), (int)mm >> 16);
            m = __reduce_min_sync(FULL_MASK, m);
            const int D = m - p.margin;                       // in-flight shuffles are up to two steps (6) older than h
            if (D > 0) {
                const uint32_t Dp = (uint32_t)D * 0x10001u;       // every half is >= D: no borrow between halves
#pragma unroll
                for (int r = 0; r < R; ++r) h[r] -= Dp;
                dprev -= Dp;
                q0 -= Dp;
                q1 -= Dp;
                base += D;
                // the next block's top row is already in stored form: move it to the new base too
                uint32_t* sn = sin + (((b & 1) ^ 1) << 5) + lane;
                *sn = (*sn - (uint32_t)D) & 0xffffu;
                __syncwarp();
            }
        }
    }
#if NW_L2_DBG & 1024
    if (lane == 0 && (s % 37 == 1 || s == p.nstrips - 1))
        printf("strip %d: %lld misses in %d blocks, %lld cycles waiting\n", s, dbg_miss, nblocks, dbg_wait);
#endif
    cp_async_wait<0>();
    __syncwarp();
    {   // the last block's bottom row
        const int oc = ((nblocks - 1) << 5) - L2_SKEW + lane;
        if (oc >= 0 && oc < ncols) {
            chk_brow(p, s, oc + 1);
            st_tagged_gpu(tout + oc + 1, p.epoch, ((int)sout[(((nblocks - 1) & 1) << 5) + lane] >> 16) + pub_base);
        }
    }

    if (p.times != nullptr && lane == 0) { p.times[4 * s + 1] = global_ns(); p.times[4 * s + 3] = (unsigned long long)clock64(); }

    // right boundary column of this lane's rows (absolute G): every half is frozen at its value in the last column
    if (p.rcol != nullptr) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int a = i_lo + 1 + r, b = i_hi + 1 + r;
            const int va = (int)(short)(h[r] & 0xffffu) + base, vb = ((int)h[r] >> 16) + base;
            // (a, b <= n2: rows below the table exist when the padding sits at the bottom)
            if (a >= 1 && (!STAIR || a <= p.n2)) chk_row(p, a);
            if (b >= 1 && (!STAIR || b <= p.n2)) chk_row(p, b);
            if (p.rcol_sys) {
                if (a >= 1 && (!STAIR || a <= p.n2)) st_tagged_sys(p.rcol + a, p.epoch, va);
                if (b >= 1 && (!STAIR || b <= p.n2)) st_tagged_sys(p.rcol + b, p.epoch, vb);
            } else {
                if (a >= 1 && (!STAIR || a <= p.n2)) st_tagged_gpu(p.rcol + a, p.epoch, va);
                if (b >= 1 && (!STAIR || b <= p.n2)) st_tagged_gpu(p.rcol + b, p.epoch, vb);
            }
        }
    }
    __syncwarp();
}

template <int R, bool STAIR = false>
__global__ void __launch_bounds__(512) nw_strip16l2_kernel(const StripParams p)
{
    extern __shared__ __align__(16) uint32_t nw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t* smem = nw_smem + warp * L2_SMEM_WORDS_PER_WARP;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    wait_mailbox_free(p);
    for (int s = slot; s < p.nstrips; s += nslots) run_strip16l2<R, STAIR>(p, s, lane, smem);
}

}  // namespace nw
