// nw_packed.cuh -- the boundary-mode strip sweep on packed 16-bit lanes (DPX s16x2): two table rows per register.
//
// Same recurrence, same G = H + i + j change of variable and the same strip / tagged-boundary-row scheme as
// nw_kernels.cuh (reference arithmetic: src/serial/serial.cpp:12-31), but every 32-bit register carries TWO cells:
//   * a warp is 64 "virtual lanes" of R rows each: lane L holds virtual lane L in the low halves and virtual lane
//     L+32 in the high halves of h[0..R-1]; virtual lane v works on column t - v at step t, so the low half of lane L
//     is at column c = t - L and the high half at column c - 32;
//   * one rotating shuffle of the packed h[R-1] feeds both halves of the next lane (lane 0 takes its low half from the
//     strip's top boundary row and its high half from lane 31's low half: one PRMT with a per-lane selector);
//   * one PRMT builds both substitution weights from the two column profile words, then VIADDMNMX.S16x2 and
//     VIMNMX(3).S16x2 update both cells: 3 integer-pipe instructions per TWO cells.
// 16 bits are enough because cells that are live in a warp at the same time differ by at most 3 * (rows + columns of
// the live window) < 2^12; the warp keeps a running int32 `base` (stored = G - base) and re-bases every 32 blocks.
// Boundary rows / columns leave the kernel as absolute int32 G, exactly like the 32-bit kernel, so the two kernels
// interoperate (column-strip pipelines, checkpoint rows, finish kernel).
#pragma once
#include "nw_kernels.cuh"

namespace nw {

constexpr int TILE_ROW_WORDS = 40;     // full-table pass 2: 8 carried-over columns + the 32 columns of the current block
constexpr int RING_COPY_WORDS = 136;                       // 128-column ring + 8 words of bank skew per copy
constexpr int SMEM16_WORDS_PER_WARP = 4 * RING_COPY_WORDS + 32 + 32;   // 4 ring copies + top inputs + bottom outputs

// One 32-step block.  The per-step loop-carried chain is  SHFL -> PRMT -> max  (the running max down the rows is
// computed as  G[r] = max(P[r], up)  with the prefix maxima P[r] of the t[] off the chain): with one warp per scheduler
// the kernel is bound by that latency, not by issue slots, so the R-1 extra max instructions are free.  The operand
// vectors of the next 4 steps are loaded before the current 4 are computed (shared-memory latency off the chain too).
// LOWLAT = false (batch mode, several warps per scheduler hide the latency): the plain running max down the rows, 3R
// instructions per step and nothing more.
// TILE = true (pass 2 of full-table mode): every step also deposits the packed registers into this lane's slab of a
// shared-memory tile ([register][step], lane pitch odd => conflict-free), which the warp then writes out row by row.
template <int R, bool PRED, bool LOWLAT = true, bool TILE = false>
__device__ __forceinline__ void sweep16(uint32_t (&h)[R], uint32_t& dprev, const uint32_t (&sel)[R],
                                        const uint32_t upsel, const int src_lane, const uint32_t* __restrict__ ringm,
                                        const uint32_t* __restrict__ sin, uint32_t* sout, const int lane, const int cb,
                                        const int ncols, uint32_t& scar, uint32_t* tile_lane = nullptr)
{
    const int i0 = cb - lane + (lane & 3);      // ring index (before & 127) of this lane's low column at k = 0; 4 | i0
    uint4 clo = *reinterpret_cast<const uint4*>(ringm + (i0 & 127));
    uint4 chi = *reinterpret_cast<const uint4*>(ringm + ((i0 - 32) & 127));
    uint4 tin = *reinterpret_cast<const uint4*>(sin);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        const uint32_t cl[4] = {clo.x, clo.y, clo.z, clo.w};
        const uint32_t ch[4] = {chi.x, chi.y, chi.z, chi.w};
        const uint32_t tn[4] = {tin.x, tin.y, tin.z, tin.w};
        if (k4 < 7) {
            clo = *reinterpret_cast<const uint4*>(ringm + ((i0 + 4 * k4 + 4) & 127));
            chi = *reinterpret_cast<const uint4*>(ringm + ((i0 + 4 * k4 + 4 - 32) & 127));
            tin = *reinterpret_cast<const uint4*>(sin + 4 * k4 + 4);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = 4 * k4 + kk;
            const uint32_t s = scar;          // lane-1's last row, shuffled at the end of the previous step
            // off the chain: t[r] = max(G[i-1][j-1] + w, G[i][j-1]) and their prefix maxima, both halves at once
            uint32_t t[R];
            {
                uint32_t diag = dprev;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t w = prmt(cl[kk], ch[kk], sel[r]);
                    t[r] = __viaddmax_s16x2(diag, w, h[r]);
                    diag = h[r];
                }
            }
            const uint32_t up0 = prmt(s, tn[kk], upsel);       // {low: row above at col c, high: row above at col c-32}
            dprev = up0;
            uint32_t mask = 0xffffffffu;
            if (PRED) {
                const int col = cb + k - lane;
                mask = ((unsigned)col < (unsigned)ncols ? 0x0000ffffu : 0u) |
                       ((unsigned)(col - 32) < (unsigned)ncols ? 0xffff0000u : 0u);
            }
            if (LOWLAT) {
                // prefix maxima of t[0..r-1] (off the chain), then G[r] = max3(t[r], P[r-1], up): one instruction per row
                uint32_t P[R];
                P[0] = t[0];
#pragma unroll
                for (int r = 1; r + 1 < R; ++r) P[r] = __vmaxs2(t[r], P[r - 1]);
                {   // the last row first: it feeds the next step's shuffle
                    const uint32_t g = (R > 1) ? __vimax3_s16x2(t[R - 1], P[R > 1 ? R - 2 : 0], up0) : __vmaxs2(t[0], up0);
                    h[R - 1] = PRED ? ((g & mask) | (h[R - 1] & ~mask)) : g;
                    scar = __shfl_sync(FULL_MASK, h[R - 1], src_lane);     // next step's input: start it as early as possible
                }
#pragma unroll
                for (int r = 0; r + 1 < R; ++r) {
                    const uint32_t g = (r == 0) ? __vmaxs2(t[0], up0) : __vimax3_s16x2(t[r], P[r - 1], up0);
                    h[r] = PRED ? ((g & mask) | (h[r] & ~mask)) : g;
                }
            } else {
                uint32_t g = up0;                               // running max; every second link is a 3-input max
#pragma unroll
                for (int r = 0; r < R; r += 2) {
                    const uint32_t ga = __vmaxs2(t[r], g);
                    uint32_t gb = ga;
                    if (r + 1 < R) gb = __vimax3_s16x2(t[r + 1], t[r], g);
                    h[r] = PRED ? ((ga & mask) | (h[r] & ~mask)) : ga;
                    if (r + 1 < R) h[r + 1] = PRED ? ((gb & mask) | (h[r + 1] & ~mask)) : gb;
                    g = gb;
                }
                scar = __shfl_sync(FULL_MASK, h[R - 1], src_lane);
            }
            if (lane == 31) sout[k] = h[R - 1];
            if (TILE) {
#pragma unroll
                for (int r = 0; r < R; ++r) tile_lane[r * TILE_ROW_WORDS + 8 + k] = h[r];
            }
        }
    }
}

__device__ __forceinline__ int2 poll_tagged(const StripParams& sp, const int2* p, int epoch, int sys)
{
    int2 t;
    SpinGuard sg;
    for (;;) {
        t = sys ? ld_tagged_sys(p) : ld_tagged_gpu(p);
        if (t.x == epoch || sg.expired(sp)) return t;
        __nanosleep(100);
    }
}

// Column-strip parts: the left neighbour stores the halo words of a strip's rows all at once, when ITS strip of the same
// rows ends -- for most warps of a long chain milliseconds after the kernel started.  Wait for one representative word
// (the strip's last row) with growing sleeps, one request per warp per poll, before the lanes read their own words:
// hundreds of warps polling system-scope words at full speed slow the working warps down several times over.
__device__ __forceinline__ void wait_halo_politely(const StripParams& p, int i_rep)
{
    const int2* w = p.halo + max(i_rep, 1);
    SpinGuard sg;
    unsigned ns = 100;
    while ((p.halo_sys ? ld_tagged_sys(w) : ld_tagged_gpu(w)).x != p.epoch && !sg.expired(p)) {
        __nanosleep(ns);
        ns = min(ns * 2, 2000u);
    }
}

template <int R>
__device__ __forceinline__ void run_strip16(const StripParams& p, const int s, const int lane, uint32_t* smem)
{
    constexpr int SH = 64 * R;
    uint32_t* ring = smem;
    uint32_t* sin = smem + 4 * RING_COPY_WORDS;
    uint32_t* sout = sin + 32;
    const uint32_t* ringm = ring + (lane & 3) * RING_COPY_WORDS;
    const int ncols = p.ncols;
    const int q_lo = s * SH + lane * R;               // first padded row of the low half; the high half is 32*R below
    const int i_lo = q_lo - p.pad_top;                // table row just above the low half's first row (may be <= 0)
    const int i_hi = i_lo + 32 * R;

    uint32_t sel[R];
#pragma unroll
    for (int r = 0; r < R; ++r) sel[r] = p.rsel[(s * 32 + lane) * R + r];
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

    int2* tout = p.brow + (long long)s * p.pitch;
    const int2* tin = p.brow + (long long)(s - 1) * p.pitch;
    if (lane == 31) st_tagged_gpu(tout, p.epoch, ((int)h[R - 1] >> 16) + base);      // j = 0: the boundary column

    const uint32_t* wq = p.wq;
    chk_wq(p, lane);
    uint32_t wnext = wq[lane];
    int2 pre = make_int2(0, 0);
    if (s > 0 && lane < ncols) { chk_brow(p, s - 1, lane + 1); pre = ld_tagged_gpu(tin + lane + 1); }

    const int nblocks = (ncols + 63 + 31) >> 5;          // the high half of lane 31 reaches column ncols-1 at t = ncols+62
    uint32_t scar = __shfl_sync(FULL_MASK, h[R - 1], src_lane);
    for (int b = 0; b < nblocks; ++b) {
        const int cb = b << 5;
        // column operands of [cb, cb+32) into the four skewed ring copies; the ring then holds [cb-96, cb+32)
#pragma unroll
        for (int m = 0; m < 4; ++m) ring[m * RING_COPY_WORDS + ((cb + lane + m) & 127)] = wnext;
        {
            const int nc = cb + 32 + lane;
            wnext = wq[nc < ncols + WQ_PAD ? nc : ncols + WQ_PAD - 1];
        }
        // top boundary row of [cb, cb+32) -> stored form, low half
        if (cb < ncols) {
            int v = 0;
            if (s > 0) {
                const int col = cb + lane;
                const bool need = col < ncols;
                SpinGuard sg;
                while (!__all_sync(FULL_MASK, !need || pre.x == p.epoch)) {
                    if (need) chk_brow(p, s - 1, col + 1);
                    if (need && pre.x != p.epoch) pre = ld_tagged_gpu(tin + col + 1);
                    if (sg.expired_warp(p)) break;
                }
                v = pre.y;
                if (col + 32 < ncols) { chk_brow(p, s - 1, col + 33); pre = ld_tagged_gpu(tin + col + 33); }
            }
            sin[lane] = (uint32_t)(v - base) & 0xffffu;
            if (b == 0 && p.times != nullptr && lane == 0) { p.times[4 * s] = global_ns(); p.times[4 * s + 2] = (unsigned long long)clock64(); }
        }
        __syncwarp();
        if (cb >= 64 && cb + 31 < ncols)
            sweep16<R, false>(h, dprev, sel, upsel, src_lane, ringm, sin, sout, lane, cb, ncols, scar);
        else
            sweep16<R, true>(h, dprev, sel, upsel, src_lane, ringm, sin, sout, lane, cb, ncols, scar);
        __syncwarp();
        {
            const int oc = cb - 63 + lane;               // column finished by the high half of lane 31 at step k = lane
            if (oc >= 0 && oc < ncols) { chk_brow(p, s, oc + 1); st_tagged_gpu(tout + oc + 1, p.epoch, ((int)sout[lane] >> 16) + base); }
        }
        if ((b & 31) == 31) {                            // re-base: keep the stored values small
            uint32_t mm = dprev;
#pragma unroll
            for (int r = 0; r < R; ++r) mm = __vmins2(mm, h[r]);
            int m = min((int)(short)(mm & 0xffffu), (int)mm >> 16);
            m = __reduce_min_sync(FULL_MASK, m);
            const int D = m - p.margin;
            if (D > 0) {
                const uint32_t Dp = (uint32_t)D * 0x10001u;       // every half is >= D: no borrow between halves
#pragma unroll
                for (int r = 0; r < R; ++r) h[r] -= Dp;
                dprev -= Dp;
                scar -= Dp;
                base += D;
            }
        }
        // full-table mode, pass 1: snapshot of the warp's (skewed) register state every `tile_blocks` blocks, from which
        // pass 2 replays the blocks of a tile independently of every other tile
        if (p.snap != nullptr && ((b + 1) % p.tile_blocks) == 0 && b + 1 < nblocks) {
            uint32_t* sp = p.snap + ((long long)s * p.ntiles + (b + 1) / p.tile_blocks) * (long long)(32 * (R + 2));
#pragma unroll
            for (int r = 0; r < R; ++r) sp[r * 32 + lane] = h[r];
            sp[R * 32 + lane] = dprev;
            sp[(R + 1) * 32 + lane] = (uint32_t)base;
        }
    }

    if (p.times != nullptr && lane == 0) { p.times[4 * s + 1] = global_ns(); p.times[4 * s + 3] = (unsigned long long)clock64(); }
    // right boundary column of this lane's rows (absolute G)
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
__global__ void __launch_bounds__(512) nw_strip16_kernel(const StripParams p)
{
    extern __shared__ __align__(16) uint32_t nw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t* smem = nw_smem + warp * SMEM16_WORDS_PER_WARP;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    wait_mailbox_free(p);
    for (int s = slot; s < p.nstrips; s += nslots) run_strip16<R>(p, s, lane, smem);
}

// ---- full-table mode, pass 2 ------------------------------------------------------------------------------------------
// Tile (s, m) = blocks [m*tile_blocks, (m+1)*tile_blocks) of strip s, replayed from pass 1's snapshot (m > 0) or from the
// strip's initial state (m = 0) with the top boundary row read from brow (complete after pass 1).  Tiles are independent,
// so the whole GPU works on them at once and the kernel is bound by the table stores: each block's cells go through a
// shared-memory tile and leave as 128-byte row segments, H = G - i - j (reference layout: src/serial/serial.cpp:31).
__host__ __device__ constexpr int TILE_LANE_PITCH(int R) { return R * TILE_ROW_WORDS + 1; }
__host__ __device__ constexpr int SMEM16F_WORDS_PER_WARP(int R) { return SMEM16_WORDS_PER_WARP + 32 * TILE_LANE_PITCH(R) + 3; }

template <int R>
__global__ void __launch_bounds__(256) nw_full16_kernel(const StripParams p)
{
    extern __shared__ __align__(16) uint32_t nw_smem[];
    constexpr int SH = 64 * R;
    constexpr int WORDS = (SMEM16F_WORDS_PER_WARP(R) + 3) & ~3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t* ring = nw_smem + warp * WORDS;
    uint32_t* sin = ring + 4 * RING_COPY_WORDS;
    uint32_t* sout = sin + 32;
    uint32_t* tile = sout + 32;
    uint32_t* tile_lane = tile + lane * TILE_LANE_PITCH(R);
    const uint32_t* ringm = ring + (lane & 3) * RING_COPY_WORDS;
    const int ncols = p.ncols;
    const int nblocks = (ncols + 63 + 31) >> 5;
    const uint32_t upsel = (lane == 0) ? 0x1054u : 0x3210u;
    const int src_lane = (lane + 31) & 31;
    const long long ntasks = (long long)p.s_count * p.ntiles;

    for (long long task = (long long)blockIdx.x * nwarps + warp; task < ntasks; task += (long long)gridDim.x * nwarps) {
        const int s_rel = (int)(task / p.ntiles), m = (int)(task - (long long)s_rel * p.ntiles);
        const int s = p.s_begin + s_rel;
        const int i_lo = s * SH + lane * R - p.pad_top;      // table row just above the low half's first row
        const int i_hi = i_lo + 32 * R;
        uint32_t sel[R];
#pragma unroll
        for (int r = 0; r < R; ++r) sel[r] = p.rsel[(s * 32 + lane) * R + r];

        int base = 0;
        uint32_t h[R];
        uint32_t dprev = 0;
        if (m > 0) {
            const uint32_t* sp = p.snap + ((long long)s * p.ntiles + m) * (long long)(32 * (R + 2));
#pragma unroll
            for (int r = 0; r < R; ++r) h[r] = sp[r * 32 + lane];
            dprev = sp[R * 32 + lane];
            base = (int)sp[(R + 1) * 32 + lane];
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) h[r] = 0;
            if (p.halo != nullptr) {
                int lo[R + 1], hi[R + 1];
                int mn = 0x7fffffff;
#pragma unroll
                for (int r = -1; r < R; ++r) {
                    const int a = i_lo + 1 + r, b = i_hi + 1 + r;
                    lo[r + 1] = (a >= 1) ? p.halo[a].y : 0;
                    hi[r + 1] = (b >= 1) ? p.halo[b].y : 0;
                    mn = min(mn, min(lo[r + 1], hi[r + 1]));
                }
                base = max(__reduce_min_sync(FULL_MASK, mn) - p.margin, 0);
                dprev = ((uint32_t)(lo[0] - base) & 0xffffu) | ((uint32_t)(hi[0] - base) << 16);
#pragma unroll
                for (int r = 0; r < R; ++r)
                    h[r] = ((uint32_t)(lo[r + 1] - base) & 0xffffu) | ((uint32_t)(hi[r + 1] - base) << 16);
            }
            // table column 0 of this lane's rows: the left boundary (serial.cpp:17 / mpi-vert.cpp:57-59)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int a = i_lo + 1 + r, b = i_hi + 1 + r;
                if (a >= 1) p.table[(long long)a * p.tpitch] = (int)(short)(h[r] & 0xffffu) + base - a - p.jstart;
                if (b >= 1) p.table[(long long)b * p.tpitch] = ((int)h[r] >> 16) + base - b - p.jstart;
            }
        }

        const int2* tin = p.brow + (long long)(s - 1) * p.pitch;
        const int b0 = m * p.tile_blocks, b1 = min(nblocks, b0 + p.tile_blocks);
        const uint32_t* wq = p.wq;
        // ring contents a resumed tile still needs: columns [cb-96, cb) of its first block
        if (m > 0) {
            const int cb = b0 << 5;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int c = cb - 96 + 32 * q + lane;
                chk_wq(p, c);
                const uint32_t w = wq[c];               // c >= -WQ_PAD always holds here (cb >= 32 * tile_blocks >= 64)
#pragma unroll
                for (int mm = 0; mm < 4; ++mm) ring[mm * RING_COPY_WORDS + ((c + mm) & 127)] = w;
            }
        }
        uint32_t wnext = wq[min((b0 << 5) + lane, ncols + WQ_PAD - 1)];
        int pre = 0;                                        // top boundary row, one block ahead
        if (s > 0 && (b0 << 5) + lane < ncols) pre = tin[(b0 << 5) + lane + 1].y;
        const long long tpitch = p.tpitch;
        int32_t* const table = p.table;
        uint32_t scar = __shfl_sync(FULL_MASK, h[R - 1], src_lane);
        for (int b = b0; b < b1; ++b) {
            const int cb = b << 5;
#pragma unroll
            for (int mm = 0; mm < 4; ++mm) ring[mm * RING_COPY_WORDS + ((cb + lane + mm) & 127)] = wnext;
            {
                const int nc = cb + 32 + lane;
                wnext = wq[nc < ncols + WQ_PAD ? nc : ncols + WQ_PAD - 1];
            }
            if (cb < ncols) {
                sin[lane] = (uint32_t)(pre - base) & 0xffffu;
                pre = 0;
                if (s > 0 && cb + 32 + lane < ncols) pre = tin[cb + 32 + lane + 1].y;
            }
            __syncwarp();
            const bool interior = cb >= 64 && cb + 31 < ncols;
            if (interior)
                sweep16<R, false, false, true>(h, dprev, sel, upsel, src_lane, ringm, sin, sout, lane, cb, ncols, scar, tile_lane);
            else
                sweep16<R, true, false, true>(h, dprev, sel, upsel, src_lane, ringm, sin, sout, lane, cb, ncols, scar, tile_lane);
            __syncwarp();
            // write-out.  Slab row (L, r) holds, at position p, the packed cells of table column cb - L - 7 + p (low half;
            // the high half is 32 columns to the left and 32*R rows down): positions 0..7 are carried over from the
            // previous block, 8..39 are this block's steps.  Stores must be 32-byte-sector aligned to reach HBM speed
            // (misaligned 128-byte segments run at a third of it, tools/ubench/wr.cu), so in a run of interior blocks
            // every slab writes the aligned 32-column window [d, d+32), d = (L + 7) & 7; the few columns before the first
            // window / after the last one of a run are written as partial heads / tails.
            {
                const int row0 = s * SH - p.pad_top + 1;                 // table row of virtual lane 0, register 0
                const bool rows_real = (s > 0) || (p.pad_top == 0);
                const bool fast = interior && rows_real;
                if (fast) {
                    const bool first = (b == b0) || !(cb - 32 >= 64);
                    const bool last = (b + 1 == b1) || !(cb + 32 + 31 < ncols);
                    // H = stored + base - i - j_table, j_table = jstart + (cb - L - 7 + p)
                    const int o0 = base - p.jstart - cb + 7 - row0;
                    const long long hi_off = (long long)(32 * R) * tpitch - 32;
                    constexpr int UNR = (R >= 8) ? 2 : (R == 4 ? 4 : 8);
#pragma unroll 1
                    for (int L0 = 0; L0 < 32; L0 += UNR) {
                        uint32_t w[UNR][R];
                        int pos[UNR];
#pragma unroll
                        for (int u = 0; u < UNR; ++u) {
                            const int L = L0 + u, d = (L + 7) & 7;
                            pos[u] = d + lane;
                            if (first && pos[u] < 8) pos[u] = -1;                 // before this run's first own column
#pragma unroll
                            for (int r = 0; r < R; ++r)
                                w[u][r] = tile[L * TILE_LANE_PITCH(R) + r * TILE_ROW_WORDS + (pos[u] < 0 ? 0 : pos[u])];
                        }
#pragma unroll
                        for (int u = 0; u < UNR; ++u) {
                            const int L = L0 + u;
                            if (pos[u] >= 0) {
                                int32_t* q = table + (long long)(row0 + L * R) * tpitch + (cb - L - 7 + pos[u]);
                                NW_ASSERT(row0 + L * R >= 1 && row0 + L * R + R - 1 + 32 * R <= p.n2 && cb - L - 7 + pos[u] - 32 >= 0 &&
                                          cb - L - 7 + pos[u] < tpitch);
                                const int o = o0 + L - L * R - pos[u];
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    q[(long long)r * tpitch] = (int)(short)(w[u][r] & 0xffffu) + o - r;
                                    q[(long long)r * tpitch + hi_off] = ((int)w[u][r] >> 16) + o + (32 - 32 * R) - r;
                                }
                            }
                        }
                    }
                    if (last) {                                  // tail: positions [d + 32, 40) of every slab
                        for (int L = 0; L < 32; ++L) {
                            const int pp = ((L + 7) & 7) + 32 + lane;
                            if (pp < TILE_ROW_WORDS) {
                                int32_t* q = table + (long long)(row0 + L * R) * tpitch + (cb - L - 7 + pp);
                                const int o = o0 + L - L * R - pp;
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    const uint32_t ww = tile[L * TILE_LANE_PITCH(R) + r * TILE_ROW_WORDS + pp];
                                    q[(long long)r * tpitch] = (int)(short)(ww & 0xffffu) + o - r;
                                    q[(long long)r * tpitch + hi_off] = ((int)ww >> 16) + o + (32 - 32 * R) - r;
                                }
                            }
                        }
                    }
                } else {
                    for (int L = 0; L < 32; ++L) {
                        const uint32_t* tl = tile + L * TILE_LANE_PITCH(R) + 8 + lane;
                        const int clo = cb + lane - L, chi = clo - 32;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const uint32_t w = tl[r * TILE_ROW_WORDS];
                            const int ilo = row0 + L * R + r, ihi = ilo + 32 * R;
                            if (ilo >= 1 && clo >= 0 && clo < ncols)
                                table[(long long)ilo * tpitch + clo + 1] =
                                    (int)(short)(w & 0xffffu) + base - ilo - (p.jstart + clo + 1);
                            if (ihi >= 1 && chi >= 0 && chi < ncols)
                                table[(long long)ihi * tpitch + chi + 1] = ((int)w >> 16) + base - ihi - (p.jstart + chi + 1);
                        }
                    }
                }
                __syncwarp();
                // carry the last 8 columns of every slab row over to positions 0..7 for the next block
                if (fast) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        uint32_t c8[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) c8[q] = tile_lane[r * TILE_ROW_WORDS + 32 + q];
#pragma unroll
                        for (int q = 0; q < 8; ++q) tile_lane[r * TILE_ROW_WORDS + q] = c8[q];
                    }
                }
            }
            __syncwarp();
            if ((b & 31) == 31) {                            // the same re-basing decisions as pass 1
                uint32_t mm = dprev;
#pragma unroll
                for (int r = 0; r < R; ++r) mm = __vmins2(mm, h[r]);
                int mv = min((int)(short)(mm & 0xffffu), (int)mm >> 16);
                mv = __reduce_min_sync(FULL_MASK, mv);
                const int D = mv - p.margin;
                if (D > 0) {
                    const uint32_t Dp = (uint32_t)D * 0x10001u;
#pragma unroll
                    for (int r = 0; r < R; ++r) h[r] -= Dp;
                    dprev -= Dp;
                    scar -= Dp;
                    base += D;
                    // the carried-over columns in the tile are in the old base too; they may be smaller than D, so
                    // subtract half by half
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const uint32_t w = tile_lane[r * TILE_ROW_WORDS + q];
                            const int lo = (int)(short)(w & 0xffffu) - D, hi = ((int)w >> 16) - D;
                            tile_lane[r * TILE_ROW_WORDS + q] = ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16);
                        }
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace nw
