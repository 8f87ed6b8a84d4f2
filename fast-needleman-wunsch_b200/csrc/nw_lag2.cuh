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
//   * the column profile ring (256 columns, two bank-skewed copies so that each lane's 4-column window is one aligned
//     LDS.128) is refilled by cp.async (LDGSTS) two blocks ahead; the top boundary row is fetched late in the previous
//     block (so a strip trails its predecessor by the hand-off latency, not by a whole block more); the bottom row of
//     block b is published while block b+1 is being computed.
// The price is twice the skew between virtual lanes: a strip starts 126 + 32 columns (plus the hand-off latency) after
// its predecessor.
#pragma once
#include "nw_packed.cuh"

namespace nw {

constexpr int L2_COPY_WORDS = 256 + 16;                               // 256-column ring + 16 words of bank skew
constexpr int L2_SMEM_WORDS_PER_WARP = 2 * L2_COPY_WORDS + 32 + 64;   // two ring copies + top inputs + 2 x bottom outputs
constexpr int L2_SKEW = 126;                                          // columns between virtual lane 0 and virtual lane 63

__device__ __forceinline__ void cp_async8(uint32_t* smem_dst, const uint32_t* gsrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One 32-step block.  q0 / q1: the shuffles issued two steps / one step ago.  mid2 / mid5 are run between steps (after 8
// and after 20 of the 32): independent work of the block loop that would otherwise sit, latency exposed, between sweeps.
template <int R, class Mid2, class Mid5>
__device__ __forceinline__ void sweep16l2(uint32_t (&h)[R], uint32_t& dprev, const uint32_t (&sel)[R], const uint32_t upsel,
                                          const int src_lane, const uint32_t* __restrict__ ringm,
                                          const uint32_t* __restrict__ sin, uint32_t* sout, const int lane, const int cb,
                                          uint32_t& q0, uint32_t& q1, Mid2&& mid2, Mid5&& mid5)
{
    const int i0 = cb - 2 * lane + 2 * (lane & 1);      // ring index (before & 255) of this lane's low column at k = 0; 4 | i0
    uint4 clo = *reinterpret_cast<const uint4*>(ringm + (i0 & 255));
    uint4 chi = *reinterpret_cast<const uint4*>(ringm + ((i0 - 64) & 255));
    uint4 tin = *reinterpret_cast<const uint4*>(sin);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        const uint32_t cl[4] = {clo.x, clo.y, clo.z, clo.w};
        const uint32_t ch[4] = {chi.x, chi.y, chi.z, chi.w};
        const uint32_t tn[4] = {tin.x, tin.y, tin.z, tin.w};
        if (k4 < 7) {
            clo = *reinterpret_cast<const uint4*>(ringm + ((i0 + 4 * k4 + 4) & 255));
            chi = *reinterpret_cast<const uint4*>(ringm + ((i0 + 4 * k4 + 4 - 64) & 255));
            tin = *reinterpret_cast<const uint4*>(sin + 4 * k4 + 4);
        }
        if (k4 == 2) mid2();
        if (k4 == 5) mid5();
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
                    const uint32_t w = prmt(cl[kk], ch[kk], sel[r]);
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

template <int R>
__device__ __forceinline__ void run_strip16l2(const StripParams& p, const int s, const int lane, uint32_t* smem)
{
    constexpr int SH = 64 * R;
    uint32_t* ring = smem;
    uint32_t* sin = smem + 2 * L2_COPY_WORDS;
    uint32_t* sout = sin + 32;                         // 2 x 32, by block parity
    const uint32_t* ringm = ring + (lane & 1) * L2_COPY_WORDS;
    const int ncols = p.ncols;
    const int q_lo = s * SH + lane * R;               // first padded row of the low half; the high half is 32*R below
    const int i_lo = q_lo - p.pad_top;                // table row just above the low half's first row (may be <= 0)
    const int i_hi = i_lo + 32 * R;

    uint32_t sel[R];
#pragma unroll
    for (int r = 0; r < R; ++r) sel[r] = p.rsel[(s * 32 + lane) * R + r];
    const uint32_t upsel = (lane == 0) ? 0x1054u : 0x3210u;
    const int src_lane = (lane + 31) & 31;

    // profile ring: zero (columns < 0), then columns [0, 64) on their way
    __syncwarp();
    for (int x = lane; x < 2 * L2_COPY_WORDS; x += 32) ring[x] = 0u;
    __syncwarp();
    const uint32_t* wq = p.wq;
    // lanes 0..15 fill copy 0 (column c at word c & 255), lanes 16..31 copy 1 (column c at word (c + 2) & 255)
    const int fx = 2 * (lane & 15);
    uint32_t* const fdst = ring + (lane >> 4) * L2_COPY_WORDS;
    const int fskew = (lane >> 4) * 2;
    auto fill = [&](int c0) {      // columns [c0, c0 + 32); the array is zero-padded on both sides
        cp_async8(fdst + ((c0 + fx + fskew) & 255), wq + c0 + fx);
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
        int lo[R + 1], hi[R + 1];
        int mn = 0x7fffffff;
#pragma unroll
        for (int r = -1; r < R; ++r) {
            const int a = i_lo + 1 + r, b = i_hi + 1 + r;
            lo[r + 1] = (a >= 1) ? poll_tagged(p.halo + a, p.epoch, p.halo_sys).y : 0;
            hi[r + 1] = (b >= 1) ? poll_tagged(p.halo + b, p.epoch, p.halo_sys).y : 0;
            mn = min(mn, min(lo[r + 1], hi[r + 1]));
        }
        base = max(__reduce_min_sync(FULL_MASK, mn) - 16, 0);
        dprev = ((uint32_t)(lo[0] - base) & 0xffffu) | ((uint32_t)(hi[0] - base) << 16);
#pragma unroll
        for (int r = 0; r < R; ++r) h[r] = ((uint32_t)(lo[r + 1] - base) & 0xffffu) | ((uint32_t)(hi[r + 1] - base) << 16);
    }

    int2* tout = p.brow + (long long)s * p.pitch;
    const int2* tin = p.brow + (long long)(s - 1) * p.pitch;
    if (lane == 31) st_tagged_gpu(tout, p.epoch, ((int)h[R - 1] >> 16) + base);      // j = 0: the boundary column

    // top boundary row of block b: tagged word of column min(cb + lane, ncols - 1) (past the last column the row above is
    // frozen at its last value, like this strip's own rows)
    const int clast = ncols - 1;
    int2 pre = make_int2(0, 0);
    if (s > 0) pre = ld_tagged_gpu(tin + min(lane, clast) + 1);

    const int nblocks = (ncols + L2_SKEW + 31) >> 5;   // the high half of lane 31 reaches column ncols-1 at t = ncols+125
    uint32_t q0 = __shfl_sync(FULL_MASK, h[R - 1], src_lane), q1 = q0;
    int pub_base = base;
    for (int b = 0; b < nblocks; ++b) {
        const int cb = b << 5;
        cp_async_wait<1>();                            // columns [cb, cb+32) have landed (this lane's part)
        if (s > 0) {
            const int2* a = tin + min(cb + lane, clast) + 1;
            while (!__all_sync(FULL_MASK, pre.x == p.epoch)) {
                if (pre.x != p.epoch) pre = ld_tagged_gpu(a);
            }
        }
        sin[lane] = (uint32_t)(pre.y - base) & 0xffffu;
        __syncwarp();                                  // ring + sin visible to every lane; previous block's sout complete
        fill(cb + 64);
        uint32_t* so = sout + ((b & 1) << 5);
        const uint32_t* so_prev = sout + (((b & 1) ^ 1) << 5);
        sweep16l2<R>(h, dprev, sel, upsel, src_lane, ringm, sin, so, lane, cb, q0, q1,
            [&] {       // publish the bottom row of the previous block: column finished by lane 31's high half at step k = lane
                const int oc = cb - 32 - L2_SKEW + lane;
                if (b > 0 && oc >= 0 && oc < ncols) st_tagged_gpu(tout + oc + 1, p.epoch, ((int)so_prev[lane] >> 16) + pub_base);
            },
            [&] {       // top boundary row of the next block, as late as its latency allows
                if (s > 0) pre = ld_tagged_gpu(tin + min(cb + 32 + lane, clast) + 1);
            });
        __syncwarp();                                  // every lane is done with sin before the next block rewrites it
        pub_base = base;
        if ((b & 31) == 31) {                            // re-base: keep the stored values small
            uint32_t mm = dprev;
#pragma unroll
            for (int r = 0; r < R; ++r) mm = __vmins2(mm, h[r]);
            int m = min((int)(short)(mm & 0xffffu), (int)mm >> 16);
            m = __reduce_min_sync(FULL_MASK, m);
            const int D = m - 16;                        // in-flight shuffles are up to two steps (6) older than h
            if (D > 0) {
                const uint32_t Dp = (uint32_t)D * 0x10001u;       // every half is >= D: no borrow between halves
#pragma unroll
                for (int r = 0; r < R; ++r) h[r] -= Dp;
                dprev -= Dp;
                q0 -= Dp;
                q1 -= Dp;
                base += D;
            }
        }
    }
    cp_async_wait<0>();
    __syncwarp();
    {   // the last block's bottom row
        const int oc = ((nblocks - 1) << 5) - L2_SKEW + lane;
        if (oc >= 0 && oc < ncols)
            st_tagged_gpu(tout + oc + 1, p.epoch, ((int)sout[(((nblocks - 1) & 1) << 5) + lane] >> 16) + pub_base);
    }

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
    __syncwarp();
}

template <int R>
__global__ void __launch_bounds__(512) nw_strip16l2_kernel(const StripParams p)
{
    extern __shared__ __align__(16) uint32_t nw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t* smem = nw_smem + warp * L2_SMEM_WORDS_PER_WARP;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    if (p.ack_in != nullptr) {          // do not overwrite a mailbox the consumer has not finished reading
        if (threadIdx.x == 0) {
            int a;
            do {
                asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(a) : "l"(p.ack_in) : "memory");
                if (a < p.epoch - 2) __nanosleep(500);
            } while (a < p.epoch - 2);
        }
        __syncthreads();
    }
    for (int s = slot; s < p.nstrips; s += nslots) run_strip16l2<R>(p, s, lane, smem);
}

}  // namespace nw
