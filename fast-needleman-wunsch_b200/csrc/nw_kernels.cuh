// nw_kernels.cuh -- sm_100a device code of the Needleman-Wunsch wavefront fill.
//
// What is computed (reference: src/serial/serial.cpp:12-31, scoring src/common/needleman-wunsch.hpp:11-13):
//     H[i][j] = max(H[i-1][j-1] + (s1[j-1]==s2[i-1]), H[i-1][j] - 1, H[i][j-1] - 1),  H[0][j] = -j, H[i][0] = -i
// The kernels work on the shifted table  G[i][j] = H[i][j] + i + j  (exact integer change of variable):
//     G[i][j] = max(G[i-1][j-1] + w, G[i-1][j], G[i][j-1]),  w = 2 + (s1[j-1]==s2[i-1]),  G[0][j] = G[i][0] = 0
// so one cell is   w = PRMT(column word, row selector);  t = VIADDMNMX(diag, w, left);  G = VIMNMX(t, up)
// -- three integer-pipe instructions, two of them Blackwell DPX (__viaddmax_s32 / max).  H = G - i - j is applied
// only where cells leave the kernel (boundary rows/columns, table stores, the score).
//
// General linear-gap scoring (the reference's only "configuration" is the three macros of
// src/common/needleman-wunsch.hpp:11-13): with match M, mismatch X and gap g the same change of variable is
//     G[i][j] = H[i][j] - g*(i + j),   w = (M or X) - 2g,   H = G + g*(i + j),
// so the kernels only see other weight bytes (w_match, w_mis; a negative weight is clamped to 0, which is exact: G is
// monotone along rows and columns, so diag + w <= up whenever w <= 0) and the emitting code multiplies by g.
//
// Decomposition: the table is cut into horizontal STRIPS of 32*R rows.  One warp owns a strip: lane L keeps R
// consecutive rows in registers and sweeps the columns left to right, one column per step, lane L one column behind
// lane L-1 (so a warp advances one anti-diagonal of 32 lane-blocks per step); the bottom cell of lane L-1 reaches lane
// L with __shfl_up_sync.  The bottom row of a strip is written to HBM as "tagged" 64-bit words {epoch, G}; the warp
// that owns the strip below polls those words (no flags, no fences: an aligned 8-byte store is single-copy atomic)
// -- the GPU form of the reference's sentinel polling (src/sentinel/sentinel-otf-mt.cpp:44-51) and per-row progress
// counters (src/idxarray/idxarray-mod-mt.cpp:54-79).  Strips are dealt round-robin to the warps of a persistent grid.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#ifdef NW_CHECK
#include <cassert>
#endif

namespace nw {

constexpr int WQ_PAD = 64;          // the column operand array is addressable for col in [-WQ_PAD, ncols + WQ_PADR)
constexpr int WQ_PADR = 384;        // (the lag-2 kernel of nw_lag2.cuh prefetches up to ~260 columns past the last one)
constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int SMEM_WORDS_PER_WARP = 128;   // 64 column operands + 32 top-row inputs + 32 bottom-row outputs

// ---- tiny PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
// L1-bypassing loads/stores of tagged words {tag, value}.  Each is ONE scalar 64-bit access: PTX guarantees single-copy
// atomicity for an aligned scalar access of up to 64 bits (a .v2.s32 access is formally two 32-bit accesses), so a
// reader sees either the old pair or the new pair, never a mix -- the hand-off needs no flag and no fence.
__device__ __forceinline__ int2 unpack_tagged(unsigned long long v) { return make_int2((int)(uint32_t)v, (int)(uint32_t)(v >> 32)); }
__device__ __forceinline__ unsigned long long pack_tagged(int tag, int val)
{
    return (unsigned long long)(uint32_t)tag | ((unsigned long long)(uint32_t)val << 32);
}
__device__ __forceinline__ int2 ld_tagged_gpu(const int2* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return unpack_tagged(v);
}
__device__ __forceinline__ int2 ld_tagged_sys(const int2* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return unpack_tagged(v);
}
__device__ __forceinline__ void st_tagged_gpu(int2* p, int tag, int val)
{
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(pack_tagged(tag, val)) : "memory");
}
__device__ __forceinline__ void st_tagged_sys(int2* p, int tag, int val)
{
    asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(p), "l"(pack_tagged(tag, val)) : "memory");
}

// ---- kernel parameters ----------------------------------------------------------------------------------------
struct StripParams {
    const uint32_t* wq;     // column operand per interior column c (0-based), valid for c in [-WQ_PAD, ncols+WQ_PADR)
                            //   4-letter path: bytes b=0..3 hold 2 + (code(s1[c]) == b);  generic path: the raw byte
    const uint32_t* rsel;   // row operand per padded row q in [0, nstrips*32*R)
                            //   4-letter path: PRMT selector (0x5550|code real rows, 0xCCCC virtual rows)
                            //   generic path:  raw byte of s2 (real rows) or 0x200 (virtual rows)
    int2* brow;             // nstrips x pitch tagged words: brow[s][j] = {epoch, G[last row of strip s][j]}, j=0..ncols
    long long pitch;
    const int2* halo;       // tagged left boundary column, indexed by table row i (1..n2); nullptr => all zero
    int2* rcol;             // tagged right boundary column, indexed by table row i (1..n2); may be PEER memory
    int32_t* table;         // FULL mode: rows 0..n2 x tpitch, this part's columns (local column 0 = left boundary)
    long long tpitch;
    int32_t* dump;          // FULL mode: scratch row (tpitch ints) that swallows the stores of virtual rows
    int ncols;              // interior columns of this part
    int n2;                 // table rows - 1
    int nstrips;
    int pad_top;            // virtual rows above table row 1: nstrips*32*R - n2
    int jstart;             // global table column of this part's left boundary column
    int epoch;              // tag of this run (never 0)
    int halo_sys;           // halo written by a peer device: poll with system scope
    int rcol_sys;           // rcol is peer memory: store with system scope
    // back-pressure of a column-strip pipeline (mailboxes are double-buffered by epoch parity):
    // full-table mode with the packed kernels (nw_packed.cuh): pass 1 snapshots / pass 2 tiles
    uint32_t* snap;         // nstrips x ntiles x 32*(R+2) words, or nullptr
    int tile_blocks;        // 32-column blocks per tile
    int ntiles;             // tiles per strip
    int s_begin, s_count;   // pass 2: the strips this launch covers (streamed table delivery runs it band by band)
    const int* ack_in;      // producer side: the consumer's "finished epoch" word (in the consumer's mailbox
                            // allocation, possibly peer memory); the kernel waits for ack >= epoch - 2
    int* abort_flag;             // device word + mapped host word: a wait that exceeds spin_ns sets both and gives up (the
    int* abort_host;             // fill's results are then garbage and the host reports NW_ERR_CUDA) -- a failed peer must
    unsigned long long spin_ns;  // not hang us.  Waiting warps re-read only the DEVICE word (a host read costs microseconds).
    int w_match, w_mis;          // G-form weights: max(M - 2g, 0), max(X - 2g, 0)  (default scoring: 3, 2)
    int gap;                     // g (default -1): H = G + g*(i + j) where cells leave a kernel
    // score mode along a staircase (lag-2 kernel only): strip s sweeps only its first widths[s] columns (a multiple of 32, or
    // ncols) and is frozen beyond them, like every strip is beyond the table's last column; tails[s * 64 + k] is the selector
    // word of column widths[s] + k with the low half (column >= width) made virtual
    const int* widths;
    const uint32_t* tails;
    int4* local_best;            // local alignment (nw_local.cuh): per strip {best score, row, column, 0}
    int margin;                  // packed kernels: slack below the warp's minimum when re-basing (2 * max weight + 10)
    unsigned long long* times;   // nstrips x 4: %globaltimer (ns) when a strip has its first top-row block and when it
                                 // ends, then clock64 (SM cycles) at the same two points -- the trace behind the start-up
                                 // lag numbers and the SM clock actually seen (nw_plan_strip_times); or nullptr
};

// ---- bounds-check build (`make check`: -DNW_CHECK) --------------------------------------------------------------------------
// compute-sanitizer is closed on the GPU pool this was developed on, so the library carries its own index checks: every
// global-memory index of the strip kernels is asserted against the allocation as nw_cuda.cu sizes it.  A failed assert
// traps the kernel; the host call then fails with NW_ERR_CUDA.  tests/test_gpu_checked.py runs every kernel family on the
// check build.  Compiled out of the product build.
#ifdef NW_CHECK
#define NW_ASSERT(c) assert(c)
#else
#define NW_ASSERT(c) ((void)0)
#endif
__device__ __forceinline__ void chk_brow(const StripParams& p, int s, long long j)      // word j of boundary row s
{
    NW_ASSERT(s >= 0 && s < max(p.nstrips, 1) && j >= 0 && j < p.pitch);
}
__device__ __forceinline__ void chk_row(const StripParams& p, int i) { NW_ASSERT(i >= 1 && i <= p.n2); }      // halo / rcol index
__device__ __forceinline__ void chk_wq(const StripParams& p, int c) { NW_ASSERT(c >= -WQ_PAD && c < p.ncols + WQ_PADR); }
__device__ __forceinline__ void chk_table(const StripParams& p, long long i, long long j)
{
    NW_ASSERT(i >= 0 && i <= p.n2 && j >= 0 && j < p.tpitch);
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Bounded spinning: every wait loop of the strip kernels calls expired() once per failed poll; after spin_ns of waiting (or
// when another warp has already given up) it returns true and the caller leaves the loop.
struct SpinGuard {
    unsigned long long t0 = 0;
    unsigned n = 0;
    __device__ __forceinline__ bool expired(const StripParams& p)
    {
        if ((++n & 1023u) != 0 || p.abort_flag == nullptr) return false;
        int a;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(a) : "l"(p.abort_flag) : "memory");
        if (a != 0) return true;
        const unsigned long long t = global_ns();
        if (t0 == 0) {
            t0 = t;
            return false;
        }
        if (t - t0 < p.spin_ns) return false;
        asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p.abort_flag), "r"(1) : "memory");
        asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p.abort_host), "r"(1) : "memory");
        return true;
    }
    // the same for a loop that the whole warp runs in lockstep (every lane calls it in every iteration): one vote per
    // 1024 iterations instead of one per iteration
    __device__ __forceinline__ bool expired_warp(const StripParams& p)
    {
        if (((n + 1) & 1023u) != 0) {
            ++n;
            return false;
        }
        return __any_sync(FULL_MASK, expired(p));
    }
};

// producer side of a column-strip pipeline: do not overwrite a mailbox the consumer has not finished reading
__device__ __forceinline__ void wait_mailbox_free(const StripParams& p)
{
    if (p.ack_in != nullptr) {
        if (threadIdx.x == 0) {
            int a;
            SpinGuard sg;
            do {
                asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(a) : "l"(p.ack_in) : "memory");
                if (a < p.epoch - 2) __nanosleep(500);
            } while (a < p.epoch - 2 && !sg.expired(p));
        }
        __syncthreads();
    }
}

// ---- the strip sweep -------------------------------------------------------------------------------------------
template <int R, bool GENERIC>
struct RowOperands {
    uint32_t sel[R];
    int wx[GENERIC ? R : 1];
    int wm;
    __device__ __forceinline__ int weight(int r, uint32_t cop) const
    {
        if (GENERIC) return (sel[r] == cop) ? wm : wx[r];
        return (int)prmt(cop, 0x80u, sel[r]);
    }
};

template <int R, bool GENERIC, bool FULL, bool PRED>
__device__ __forceinline__ void sweep32(int (&h)[R], int& dprev, const RowOperands<R, GENERIC>& ro,
                                        const uint32_t* __restrict__ Wl, const int* __restrict__ sin, int* sout,
                                        const int lane, const int cb, const int ncols,
                                        int32_t* const (&trow)[FULL ? R : 1], const int (&hoff)[FULL ? R : 1], const int gap = -1)
{
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const uint32_t cop = Wl[k];
        int up = __shfl_up_sync(FULL_MASK, h[R - 1], 1);
        if (lane == 0) up = sin[k];
        const int col = cb + k - lane;
        if (!PRED || (col >= 0 && col < ncols)) {
            const int gcol = FULL ? gap * col : 0;
            int diag = dprev;
            dprev = up;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int w = ro.weight(r, cop);
                const int t = __viaddmax_s32(diag, w, h[r]);   // max(G[i-1][j-1] + w, G[i][j-1])
                diag = h[r];
                up = max(t, up);                               // ... , G[i-1][j])
                h[r] = up;
                if (FULL) {
                    NW_ASSERT(col + 1 >= 1 && col + 1 <= ncols);
                    trow[r][col + 1] = up + hoff[r] + gcol;   // H = G + g*(i + j)
                }
            }
        }
        if (lane == 31) sout[k] = h[R - 1];
    }
}

template <int R, bool GENERIC, bool FULL>
__device__ __forceinline__ void run_strip(const StripParams& p, const int s, const int lane, uint32_t* W, int* sin,
                                          int* sout)
{
    constexpr int SH = 32 * R;
    const int ncols = p.ncols;
    const int q0 = s * SH + lane * R;            // first padded row of this lane
    const int i0 = q0 - p.pad_top;               // table row just above this lane's first row (may be <= 0)

    RowOperands<R, GENERIC> ro;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t v = p.rsel[q0 + r];
        ro.sel[r] = v;
        if (GENERIC) ro.wx[r] = (v == 0x200u) ? -1 : p.w_mis;
    }
    ro.wm = p.w_match;

    // left boundary column (G form): zero for a whole table, the neighbour's right column for a column strip
    int h[R];
    int dprev = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) h[r] = 0;
    if (p.halo != nullptr) {
#pragma unroll
        for (int r = -1; r < R; ++r) {
            const int i = i0 + 1 + r;
            int v = 0;
            if (i >= 1) {
                int2 t;
                SpinGuard sg;
                do {
                    chk_row(p, i);
                    t = p.halo_sys ? ld_tagged_sys(p.halo + i) : ld_tagged_gpu(p.halo + i);
                    if (t.x != p.epoch) __nanosleep(100);
                } while (t.x != p.epoch && !sg.expired(p));
                v = t.y;
            }
            if (r < 0) dprev = v; else h[r] = v;
        }
    }

    int32_t* trow[FULL ? R : 1];
    int hoff[FULL ? R : 1];
    if (FULL) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = i0 + 1 + r;
            if (i >= 1) chk_table(p, i, ncols);
            trow[FULL ? r : 0] = (i >= 1) ? p.table + (long long)i * p.tpitch : p.dump;
            hoff[FULL ? r : 0] = p.gap * (i + p.jstart + 1);     // H = G + g*(i + jstart + col + 1)
            trow[FULL ? r : 0][0] = h[r] + p.gap * (i + p.jstart);   // left boundary column (serial.cpp:17 / mpi-vert.cpp:57-59)
        }
    }

    int2* tout = p.brow + (long long)s * p.pitch;
    const int2* tin = p.brow + (long long)(s - 1) * p.pitch;
    if (lane == 31) st_tagged_gpu(tout, p.epoch, h[R - 1]);

    // column operand window: W[0..63] holds columns [cb-32, cb+32) of the current 32-step block
    const uint32_t* wq = p.wq;
    chk_wq(p, lane - 32);
    chk_wq(p, 32 + lane);
    W[lane] = wq[lane - 32];
    W[lane + 32] = wq[lane];
    uint32_t wnext = wq[32 + lane];
    const uint32_t* Wl = W + 32 - lane;

    // top boundary row: prefetched one block ahead
    int2 pre = make_int2(0, 0);
    if (s > 0 && lane < ncols) { chk_brow(p, s - 1, lane + 1); pre = ld_tagged_gpu(tin + lane + 1); }
    if (s == 0) sin[lane] = 0;

    const int nblocks = (ncols + 31 + 31) >> 5;     // steps t = 0 .. ncols+30
    for (int b = 0; b < nblocks; ++b) {
        const int cb = b << 5;
        if (b > 0) {
            W[lane] = W[lane + 32];
            W[lane + 32] = wnext;
            int nc = cb + 32 + lane;
            chk_wq(p, nc < ncols + WQ_PAD ? nc : ncols + WQ_PAD - 1);
            wnext = wq[nc < ncols + WQ_PAD ? nc : ncols + WQ_PAD - 1];
        }
        if (s > 0 && cb < ncols) {
            const int col = cb + lane;
            const bool need = col < ncols;
            if (need) chk_brow(p, s - 1, col + 1);
            SpinGuard sg;
            while (!__all_sync(FULL_MASK, !need || pre.x == p.epoch)) {
                __nanosleep(100);
                if (need && pre.x != p.epoch) pre = ld_tagged_gpu(tin + col + 1);
                if (sg.expired_warp(p)) break;
            }
            sin[lane] = pre.y;
            if (col + 32 < ncols) { chk_brow(p, s - 1, col + 33); pre = ld_tagged_gpu(tin + col + 33); }
        }
        __syncwarp();
        if (cb >= 31 && cb + 31 < ncols)
            sweep32<R, GENERIC, FULL, false>(h, dprev, ro, Wl, sin, sout, lane, cb, ncols, trow, hoff, p.gap);
        else
            sweep32<R, GENERIC, FULL, true>(h, dprev, ro, Wl, sin, sout, lane, cb, ncols, trow, hoff, p.gap);
        __syncwarp();
        const int oc = cb - 31 + lane;               // column finished by lane 31 at step k = lane of this block
        const int ov = sout[lane];
        if (oc >= 0 && oc < ncols) { chk_brow(p, s, oc + 1); st_tagged_gpu(tout + oc + 1, p.epoch, ov); }
    }

    // right boundary column of this lane's rows
    if (p.rcol != nullptr) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = i0 + 1 + r;
            if (i >= 1) {
                chk_row(p, i);
                if (p.rcol_sys) st_tagged_sys(p.rcol + i, p.epoch, h[r]);
                else st_tagged_gpu(p.rcol + i, p.epoch, h[r]);
            }
        }
    }
    __syncwarp();
}

template <int R, bool GENERIC, bool FULL>
__global__ void __launch_bounds__(512) nw_strip_kernel(const StripParams p)
{
    extern __shared__ uint32_t nw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t* W = nw_smem + warp * SMEM_WORDS_PER_WARP;
    int* sin = (int*)(W + 64);
    int* sout = sin + 32;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    wait_mailbox_free(p);
    for (int s = slot; s < p.nstrips; s += nslots) run_strip<R, GENERIC, FULL>(p, s, lane, W, sin, sout);
}

// ---- preparation / finishing kernels (a few hundred KB of traffic; not on the critical path) --------------------
// presence bitmap of the byte values that occur in a sequence
__global__ void nw_presence_kernel(const uint8_t* s, int n, uint32_t* bitmap /*8 words*/)
{
    __shared__ uint32_t bm[8];
    if (threadIdx.x < 8) bm[threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t v = s[i];
        atomicOr(&bm[v >> 5], 1u << (v & 31));
    }
    __syncthreads();
    if (threadIdx.x < 8 && bm[threadIdx.x]) atomicOr(&bitmap[threadIdx.x], bm[threadIdx.x]);
}

struct EncodeParams {
    const uint8_t* s1;      // this part's slice: ncols bytes
    const uint8_t* s2;      // n2 bytes
    uint32_t* wq_base;      // ncols + WQ_PAD + WQ_PADR words (index c + WQ_PAD); the padding is the zero-weight word
    uint32_t* rsel;         // nrows_padded words
    int ncols, n2, nrows_padded, pad_top;
    int generic;
    int packed_regs;        // 0: 32-bit kernels; R > 0: packed kernel with R registers per lane (strip = 64*R rows)
    int lag2;               // packed lag-2 kernel (nw_lag2.cuh): the roles of row and column operands are swapped
    int w_match, w_mis;     // G-form weights (default 3, 2); the 4-letter paths need 0 <= w <= 127
    uint8_t code[256];      // 4-letter path: byte value -> 0..3
};

// profile word of a letter: byte b = weight of (this letter, letter with code b)
__device__ __forceinline__ uint32_t weight_word(uint32_t code, int w_match, int w_mis)
{
    return (uint32_t)w_mis * 0x01010101u + ((uint32_t)(w_match - w_mis) << (8 * code));
}

__global__ void nw_encode_kernel(const EncodeParams e)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int x = tid; x < e.ncols + WQ_PAD + WQ_PADR && !(e.packed_regs > 0 && e.lag2); x += nth) {
        const int c = x - WQ_PAD;
        uint32_t v;
        if (c >= 0 && c < e.ncols) v = e.generic ? (uint32_t)e.s1[c] : weight_word(e.code[e.s1[c]], e.w_match, e.w_mis);
        else v = e.generic ? 0x100u : 0u;      // weight 0: a virtual column repeats its left neighbour (nw_lag2.cuh)
        e.wq_base[x] = v;
    }
    if (e.packed_regs > 0 && e.lag2) {
        // lag-2 kernel (nw_lag2.cuh): w = PRMT(A, B, S).  Per (strip, lane, register) two row words A (low half's row) and B
        // (high half's row) with byte b = 2 + (code(row letter) == b), zero for a virtual row; per column c ONE selector
        // S[c] for the pair of columns (c, c - 64) the two halves of a lane work on: nibble 0 = code(s1[c]) picks A's byte,
        // nibble 2 = 4 + code(s1[c-64]) picks B's; nibbles 1, 3 and those of a virtual column replicate a sign bit = 0
        const int R = e.packed_regs;
        for (int x = tid; x < e.ncols + WQ_PAD + WQ_PADR; x += nth) {
            const int c = x - WQ_PAD, c2 = c - 64;
            const uint32_t n0 = (c >= 0 && c < e.ncols) ? e.code[e.s1[c]] : 0x8u;
            const uint32_t n2 = (c2 >= 0 && c2 < e.ncols) ? 4u + e.code[e.s1[c2]] : 0xCu;
            e.wq_base[x] = n0 | 0x80u | (n2 << 8) | 0xC000u;
        }
        for (int x = tid; x < e.nrows_padded / 2; x += nth) {
            const int s = x / (32 * R), rem = x - s * 32 * R, L = rem / R, r = rem - L * R;
            const int klo = s * 64 * R + L * R + r - e.pad_top, khi = klo + 32 * R;
            // (rows past n2 exist when the padding sits at the bottom: pad_top == 0 on the reversed half of a staircase)
            e.rsel[2 * x] = (klo >= 0 && klo < e.n2) ? weight_word(e.code[e.s2[klo]], e.w_match, e.w_mis) : 0u;
            e.rsel[2 * x + 1] = (khi >= 0 && khi < e.n2) ? weight_word(e.code[e.s2[khi]], e.w_match, e.w_mis) : 0u;
        }
        return;
    }
    if (e.packed_regs > 0) {
        // packed kernel (nw_packed.cuh): one PRMT selector per (strip, lane, register): nibble 0 picks the low half's
        // weight byte from the column word of column c, nibble 2 the high half's from the word of column c-32;
        // nibbles 1 and 3 (and both of a virtual row) replicate the sign of a small positive byte, i.e. give zero
        const int R = e.packed_regs;
        for (int x = tid; x < e.nrows_padded / 2; x += nth) {
            const int s = x / (32 * R), rem = x - s * 32 * R, L = rem / R, r = rem - L * R;
            const int klo = s * 64 * R + L * R + r - e.pad_top, khi = klo + 32 * R;
            const uint32_t n0 = (klo >= 0) ? e.code[e.s2[klo]] : 0x8u;
            const uint32_t n2 = (khi >= 0) ? 4u + e.code[e.s2[khi]] : 0xCu;
            e.rsel[x] = n0 | 0x80u | (n2 << 8) | 0xC000u;
        }
        return;
    }
    for (int q = tid; q < e.nrows_padded; q += nth) {
        const int k = q - e.pad_top;     // index into s2
        uint32_t v;
        if (k >= 0) v = e.generic ? (uint32_t)e.s2[k] : (0x5550u | e.code[e.s2[k]]);
        else v = e.generic ? 0x200u : 0xCCCCu;
        e.rsel[q] = v;
    }
}

// staircase score mode: selector words of the 64 columns behind every strip's own width (low half virtual, high half real)
__global__ void nw_encode_tails_kernel(const uint8_t* __restrict__ s1, const int* __restrict__ widths, uint32_t* tails,
                                       int nstrips, int ncols, const EncodeParams e)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int x = tid; x < nstrips * 64; x += nth) {
        const int s = x >> 6, k = x & 63;
        const int c2 = widths[s] + k - 64;      // the high half's column: inside the strip's width whenever it is >= 0
        const uint32_t n2 = (c2 >= 0 && c2 < ncols) ? 4u + e.code[s1[c2]] : 0xCu;
        tails[x] = 0x8u | 0x80u | (n2 << 8) | 0xC000u;
    }
}

// first row of a materialised table (reference: src/serial/serial.cpp:16, mpi-vert.cpp:20); the first column is
// written by the strip kernel itself because in a pipeline it is the halo, which arrives while the kernel runs.
__global__ void nw_table_row0_kernel(int32_t* table, int ncols, int jstart, int gap)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int j = tid; j <= ncols; j += nth) table[j] = gap * (jstart + j);
}

// boundary outputs in H form: last_row[j] = H[n2][jstart+j], last_col[i] = H[i][jstart+ncols].
// brow_last == nullptr  <=> there are no interior cells (n2 == 0: the last row is the init row; ncols == 0: the part is
// its boundary column only);  rcol == nullptr likewise: the last column is the left boundary (the halo, or H[i][0] = -i).
__global__ void nw_finish_kernel(const int2* brow_last, const int2* rcol, const int2* halo, int ncols, int n2,
                                 int jstart, int32_t* last_row, int32_t* last_col, int32_t* score, int* ack_out,
                                 int epoch, int gap)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int jend = jstart + ncols;
    for (int j = tid; j <= ncols; j += nth) {
        int g = 0;                                           // G of the init row / init column
        if (brow_last != nullptr) g = brow_last[j].y;
        else if (n2 > 0 && halo != nullptr) g = halo[n2].y;  // ncols == 0, j == 0
        last_row[j] = g + gap * (n2 + jstart + j);
    }
    for (int i = tid; i <= n2; i += nth) {
        int g = 0;
        if (i > 0) {
            if (rcol != nullptr) g = rcol[i].y;
            else if (halo != nullptr) g = halo[i].y;
        }
        last_col[i] = g + gap * (i + jend);
    }
    if (tid == 0) {
        int g = 0;
        if (n2 > 0) {
            if (rcol != nullptr) g = rcol[n2].y;
            else if (halo != nullptr) g = halo[n2].y;
        }
        *score = g + gap * (n2 + jend);
    }
    // the halo mailbox of this epoch has been consumed: let the producer (which polls this word, over NVLink when it
    // sits on another GPU) reuse it.  Stream order puts this kernel after the strip kernel.
    if (ack_out != nullptr && tid == 0)
        asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(ack_out), "r"(epoch) : "memory");
}

// dst[i] = src[n-1-i]: the bottom half of a score-mode plan runs forwards on both sequences reversed
__global__ void nw_reverse_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[n - 1 - i];
}

// NW_MODE_SCORE: F[j] = H of the top half's last row, B[j'] = the same for the reversed bottom half;
// H[n2][n1] = max_j F[j] + B[n1 - j]   (one block)
__global__ void nw_set_int_kernel(int32_t* p, int32_t v) { *p = v; }

__global__ void __launch_bounds__(256) nw_bidir_combine_kernel(const int32_t* __restrict__ F, const int32_t* __restrict__ B,
                                                               int n1, int32_t* score /* preset to INT_MIN */)
{
    __shared__ int red[8];
    int best = INT_MIN;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= n1; j += gridDim.x * blockDim.x) best = max(best, F[j] + B[n1 - j]);
    best = __reduce_max_sync(FULL_MASK, best);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : INT_MIN;
        v = __reduce_max_sync(FULL_MASK, v);
        if (threadIdx.x == 0 && v != INT_MIN) atomicMax(score, v);
    }
}

// NW_MODE_SCORE along a staircase: the forward plan filled, of strip s (table rows (r_s, r_s+1]), columns [0, x_s]; the
// reversed plan filled the rest.  Every monotone path crosses the 4-connected staircase
//   { (r_s+1, j) : x_s+1 <= j <= x_s }  and  { (i, x_s) : r_s < i < r_s+1 },   s = 0 .. S-1,  x_S = 0,
// in a vertex, so H[n2][n1] = max over those vertices of F + B.  F comes from the forward strips' bottom rows and right
// columns, B from the reversed plan's (strip t = S-1-s covers the same table rows; its predecessor's bottom row is table
// row r_s+1).  All values are in G form: F_H + B_H = F_G + B_G + gap * (n1 + n2).
struct StairParams {
    const int2* brow_f;   long long pitch_f;
    const int2* brow_b;   long long pitch_b;
    const int2* rcol_f;   // indexed by table row
    const int2* rcol_b;   // indexed by reversed table row
    const int* widths;    // x_s
    int nstrips, strip_rows, pad_top, n1, n2, gap;
    int32_t* score;       // preset to gap * (n1 + n2): the all-gap path through the corner
};
__global__ void __launch_bounds__(256) nw_stair_combine_kernel(const StairParams p)
{
    __shared__ int red[8];
    int best = 0;
    for (int s = blockIdx.x; s < p.nstrips; s += gridDim.x) {
        const int xs = p.widths[s], xn = (s + 1 < p.nstrips) ? p.widths[s + 1] : 0;
        const int r_lo = max(s * p.strip_rows - p.pad_top, 0), r_hi = (s + 1) * p.strip_rows - p.pad_top;
        const int t = p.nstrips - 1 - s;
        const int2* bf = p.brow_f + (long long)s * p.pitch_f;
        const int2* bb = (t >= 1) ? p.brow_b + (long long)(t - 1) * p.pitch_b : nullptr;
        for (int j = xn + threadIdx.x; j <= xs; j += blockDim.x) best = max(best, bf[j].y + (bb ? bb[p.n1 - j].y : 0));
        for (int i = r_lo + 1 + threadIdx.x; i < r_hi; i += blockDim.x) best = max(best, p.rcol_f[i].y + p.rcol_b[p.n2 - i].y);
    }
    best = __reduce_max_sync(FULL_MASK, best);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0;
        v = __reduce_max_sync(FULL_MASK, v);
        if (threadIdx.x == 0) atomicMax(p.score, v + p.gap * (p.n1 + p.n2));
    }
}

// strip boundary row k in H form (checkpoint rows kept in HBM)
__global__ void nw_strip_row_kernel(const int2* brow, int ncols, int row_i, int jstart, int32_t* out, int gap)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int j = tid; j <= ncols; j += nth) out[j] = brow[j].y + gap * (row_i + jstart + j);
}

// table column 0 of a part that has no interior column (n1 == 0): H[i][0] = -i (src/serial/serial.cpp:17)
__global__ void nw_table_col0_kernel(int32_t* table, long long tpitch, int n2, int gap)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid + 1; i <= n2; i += nth) table[(long long)i * tpitch] = gap * i;
}

// presence bitmap over a large byte array (batch inputs), 16 bytes per load where aligned
__global__ void nw_presence_kernel64(const uint8_t* s, long long n, uint32_t* bitmap /*8 words*/)
{
    __shared__ uint32_t bm[8];
    if (threadIdx.x < 8) bm[threadIdx.x] = 0;
    __syncthreads();
    uint32_t loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = gridDim.x * (long long)blockDim.x;
    const long long head = min(n, (long long)((16 - ((uintptr_t)s & 15)) & 15));
    for (long long i = tid; i < head; i += nth) loc[s[i] >> 5] |= 1u << (s[i] & 31);
    const uint4* v = (const uint4*)(s + head);
    const long long nv = (n - head) >> 4;
    for (long long i = tid; i < nv; i += nth) {
        const uint4 q = v[i];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t x = (w[k] >> (8 * b)) & 0xffu;
                loc[x >> 5] |= 1u << (x & 31);
            }
    }
    for (long long i = head + (nv << 4) + tid; i < n; i += nth) loc[s[i] >> 5] |= 1u << (s[i] & 31);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (loc[k]) atomicOr(&bm[k], loc[k]);
    __syncthreads();
    if (threadIdx.x < 8 && bm[threadIdx.x]) atomicOr(&bitmap[threadIdx.x], bm[threadIdx.x]);
}

// ---- traceback on a materialised table ----------------------------------------------------------------
// One CTA walks back from a start cell.  The table lives in global memory, so a naive walk would pay one dependent global
// load per step; instead the CTA stages a 64 x 64 window ending at the current cell in shared memory (coalesced row
// segments), thread 0 walks inside it until it leaves, and the window is re-staged.  Rule per cell (the reference
// defines none; this is the oracle's): diagonal if H[i][j] == H[i-1][j-1] + (s1[j-1]==s2[i-1] ? match : mismatch), else up
// if H[i][j] == H[i-1][j] + gap, else left.  Output: the two gapped sequences REVERSED (gap = 0, README.md:8).
// MODE WALK_TABLE: the table is the whole table; walk to (0, 0), moves along row 0 / column 0 are forced.
// MODE WALK_TILE:  the table is one tile of it (row 0 / column 0 = its top / left boundary); stop on reaching either.
// MODE WALK_LOCAL: a Smith-Waterman table; stop at the first cell whose value is 0 (pos[3] is set).
// pos[0..3] (shared) = i, j, emitted, finished; s1[j-1] / s2[i-1] are the letters of table column j / row i of THIS table.
constexpr int WALK_TABLE = 0, WALK_TILE = 1, WALK_LOCAL = 2;
constexpr int TB_W = 64;
struct WalkSmem {
    int win[TB_W][TB_W + 1];
    uint8_t c1[TB_W], c2[TB_W];
};
// cell accessors: a row-major table, and the phase-major scratch of the tile fill (see nw_tile_trace_kernel)
struct RowMajorCells {
    const int32_t* table;
    long long tpitch;
    static constexpr bool ROW_FASTEST = false;      // staging order that coalesces: columns fastest
    __device__ __forceinline__ int operator()(int i, int j) const { return table[(long long)i * tpitch + j]; }
};
template <int MODE, class Cells>
__device__ __forceinline__ void walk_table(const Cells cells, const uint8_t* __restrict__ s1,
                                           const uint8_t* __restrict__ s2, WalkSmem& w, volatile int* pos, uint8_t* out1,
                                           uint8_t* out2, int sc_match, int sc_mis, int sc_gap)
{
    for (;;) {
        const int i = pos[0], j = pos[1];
        if (MODE == WALK_TABLE ? (i == 0 && j == 0) : (i == 0 || j == 0)) break;
        if (MODE == WALK_LOCAL && pos[3] != 0) break;
        const int wi0 = max(i - (TB_W - 1), 0), wj0 = max(j - (TB_W - 1), 0);     // window = rows wi0..i, cols wj0..j
        const int nr = i - wi0 + 1, nc = j - wj0 + 1;
        if (Cells::ROW_FASTEST) {
            for (int x = threadIdx.x; x < nc * TB_W; x += blockDim.x) {
                const int c = x / TB_W, r = x - c * TB_W;
                if (r < nr) w.win[r][c] = cells(wi0 + r, wj0 + c);
            }
        } else {
            for (int x = threadIdx.x; x < nr * TB_W; x += blockDim.x) {
                const int r = x / TB_W, c = x - r * TB_W;
                if (c < nc) w.win[r][c] = cells(wi0 + r, wj0 + c);
            }
        }
        // s1[jj-1] for table columns jj = wj0+1..j  ->  c1[jj - wj0];  s2[ii-1] for rows ii = wi0+1..i -> c2[ii - wi0]
        for (int c = threadIdx.x; c < TB_W; c += blockDim.x) {
            if (c >= 1 && c < nc) w.c1[c] = s1[wj0 + c - 1];
            if (c >= 1 && c < nr) w.c2[c] = s2[wi0 + c - 1];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int a = nr - 1, b = nc - 1, k = pos[2];            // window coordinates of (i, j)
            // walk while the three neighbours are inside the window, or the move is forced by a table edge
            for (;;) {
                const int gi = wi0 + a, gj = wj0 + b;
                if (MODE == WALK_TABLE ? (gi == 0 && gj == 0) : (gi == 0 || gj == 0)) break;
                if (MODE == WALK_LOCAL && w.win[a][b] == 0) { pos[3] = 1; break; }
                int move;                                      // 0 diag, 1 up, 2 left
                if (gi == 0) move = 2;
                else if (gj == 0) move = 1;
                else {
                    if (a == 0 || b == 0) break;               // neighbours outside: re-stage the window
                    const int h = w.win[a][b];
                    if (h == w.win[a - 1][b - 1] + (w.c1[b] == w.c2[a] ? sc_match : sc_mis)) move = 0;
                    else if (h == w.win[a - 1][b] + sc_gap) move = 1;
                    else move = 2;
                }
                if (move == 2 && gi == 0 && b == 0) break;     // need the previous window for the sequence byte
                if (move == 1 && gj == 0 && a == 0) break;
                if (move == 0) { out1[k] = w.c1[b]; out2[k] = w.c2[a]; --a; --b; }
                else if (move == 1) { out1[k] = 0; out2[k] = w.c2[a]; --a; }
                else { out1[k] = w.c1[b]; out2[k] = 0; --b; }
                ++k;
            }
            pos[0] = wi0 + a;
            pos[1] = wj0 + b;
            pos[2] = k;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) nw_traceback_kernel(const int32_t* __restrict__ table, long long tpitch,
                                                           const uint8_t* __restrict__ s1, const uint8_t* __restrict__ s2,
                                                           int n1, int n2, uint8_t* out1, uint8_t* out2, int* out_len,
                                                           int sc_match, int sc_mis, int sc_gap)
{
    __shared__ WalkSmem w;
    __shared__ int pos[4];                 // i, j, emitted, finished
    if (threadIdx.x == 0) { pos[0] = n2; pos[1] = n1; pos[2] = 0; pos[3] = 0; }
    __syncthreads();
    walk_table<WALK_TABLE>(RowMajorCells{table, tpitch}, s1, s2, w, pos, out1, out2, sc_match, sc_mis, sc_gap);
    if (threadIdx.x == 0) *out_len = pos[2];
}

// the same on a Smith-Waterman table: from the best cell (best[1], best[2] -- what nw_local_finish_kernel left on the device)
// back to the first cell whose value is 0; out_len = number of alignment columns (0 when the best score is 0)
__global__ void __launch_bounds__(256) nw_traceback_local_kernel(const int32_t* __restrict__ table, long long tpitch,
                                                                 const uint8_t* __restrict__ s1, const uint8_t* __restrict__ s2,
                                                                 const int32_t* __restrict__ best, uint8_t* out1, uint8_t* out2,
                                                                 int* out_len, int sc_match, int sc_mis, int sc_gap)
{
    __shared__ WalkSmem w;
    __shared__ int pos[4];
    if (threadIdx.x == 0) { pos[0] = best[1]; pos[1] = best[2]; pos[2] = 0; pos[3] = 0; }
    __syncthreads();
    walk_table<WALK_LOCAL>(RowMajorCells{table, tpitch}, s1, s2, w, pos, out1, out2, sc_match, sc_mis, sc_gap);
    if (threadIdx.x == 0) *out_len = pos[2];
}

// ---- traceback WITHOUT the table: replay one tile at a time from checkpoint rows and columns -----------------------------
// After a boundary-mode fill of P column parts on one device, HBM holds a grid of checkpoints: the bottom row of every strip
// of every part (brow, every 32*R table rows) and the left boundary column of every part > 0 (its halo mailbox, = the right
// column of its neighbour).  The optimal path crosses every checkpoint row and column once.  One launch of this kernel
// handles the tile the current cell (i, j) lies in -- rows (i_top, i] of its strip, columns (j_left, j] of its part:
//   1. fill the tile in H form from its exact top row and left column (blocked wavefront: thread r owns tile row r+1 and
//      trails thread r-1 by one block of TT_B columns; neighbours exchange blocks through shared memory, one
//      __syncthreads per block step) into a scratch table -- any alphabet, any linear scoring.  The scratch is
//      PHASE-major (cell (r+1, c+1) at [((c / TT_B + r) * rpitch + r) * TT_B + c % TT_B]): at any instant the threads work on
//      different columns, and this is the layout in which their stores coalesce (row- or column-major scratch costs 32
//      cache lines per store instruction and was 7x slower);
//   2. walk back inside the scratch table (walk_table<true>) until the path reaches the tile's top row or left column;
//   3. leave the new (i, j) and the number of emitted columns in `state` for the next launch.
// Along table row 0 / column 0 the rest of the path is forced (gaps) and is emitted directly.  The host enqueues
// nstrips + nparts + 1 launches (a monotone path cannot visit more tiles); launches after the end do nothing.
struct TracePart {
    const int2* brow;       // this part's boundary rows: brow[s * pitch + c], c = 0..ncols (local table column)
    long long pitch;
    const int2* halo;       // tagged left boundary column (G form) indexed by table row, or nullptr for part 0
    int jstart, ncols;      // table column of the left boundary column; interior columns
};
struct TraceParams {
    const TracePart* parts;
    int nparts;
    const uint8_t* s1;      // the WHOLE s1 (n1 bytes) and s2
    const uint8_t* s2;
    int n1, n2, nstrips, strip_rows, pad_top;
    int sc_match, sc_mis, sc_gap;
    int32_t* scratch;       // top row (spitch ints), left column (spitch ints), then the phase-major cells
    long long spitch;       // >= widest part + 1 and >= strip_rows + 1
    int rpitch;             // threads per CTA (= strip_rows)
    int* state;             // i, j, emitted, error, then statistics: tiles, fill kcycles, walk kcycles, phases
    uint8_t* out1;
    uint8_t* out2;
};
#ifndef NW_TT_B
#define NW_TT_B 8
#endif
#ifndef NW_TT_DBG
#define NW_TT_DBG 0                // development only: bit mask of tile-fill ingredients to leave out (timing experiments)
#endif
constexpr int TT_B = NW_TT_B;      // columns per thread per block step
constexpr int TT_MAX_ROWS = 512;   // strip rows (threads)

struct TileCells {
    const int32_t* top;      // tile row 0
    const int32_t* leftc;    // tile column 0
    const int32_t* cells;
    int rpitch;
    static constexpr bool ROW_FASTEST = true;
    __device__ __forceinline__ int operator()(int a, int b) const
    {
        if (a == 0) return top[b];
        if (b == 0) return leftc[a];
        const int r = a - 1, c = b - 1;
        return cells[((long long)((c / TT_B) + r) * rpitch + r) * TT_B + (c % TT_B)];
    }
};

__global__ void __launch_bounds__(TT_MAX_ROWS) nw_tile_trace_kernel(const TraceParams p)
{
    __shared__ union {
        int up[2][TT_B][TT_MAX_ROWS];      // fill: the blocks handed from row to row (row index last: no bank conflicts)
        WalkSmem walk;                     // then the walker's window
    } sm;
    __shared__ int pos[4];
    const int tid = threadIdx.x;
    const int i = p.state[0], j = p.state[1], k0 = p.state[2];
    if ((i == 0 && j == 0) || p.state[3] != 0) return;
    const int g = p.sc_gap;
    if (i == 0 || j == 0) {      // forced: only gaps are left
        const int n = i + j;
        for (int x = tid; x < n; x += blockDim.x) {
            p.out1[k0 + x] = (i == 0) ? p.s1[j - 1 - x] : 0;
            p.out2[k0 + x] = (i == 0) ? 0 : p.s2[i - 1 - x];
        }
        __syncthreads();
        if (tid == 0) { p.state[0] = 0; p.state[1] = 0; p.state[2] = k0 + n; }
        return;
    }
    // the tile of (i, j)
    const int s = (i - 1 + p.pad_top) / p.strip_rows;
    int part = 0;
    while (part + 1 < p.nparts && p.parts[part + 1].jstart < j) ++part;
    const TracePart tp = p.parts[part];
    const int i_top = max(s * p.strip_rows - p.pad_top, 0);
    const int nr = i - i_top, jl = tp.jstart, wd = j - jl;       // tile rows 1..nr, columns 1..wd (0 = boundaries)
    if (nr > blockDim.x || wd + 1 > p.spitch || nr + 1 > p.spitch || wd > tp.ncols || (int)blockDim.x != p.rpitch) {
        if (tid == 0) p.state[3] = 1;
        return;
    }
    int32_t* const top = p.scratch;
    int32_t* const leftc = p.scratch + p.spitch;
    int32_t* const U = p.scratch + 2 * p.spitch;
    const int rp = p.rpitch;
    // top boundary row (H form): the init row of the table, or the checkpoint row of the strip above
    const int2* trow = (i_top > 0) ? tp.brow + (long long)(s - 1) * tp.pitch : nullptr;
    for (int c = tid; c <= wd; c += blockDim.x) top[c] = (trow ? trow[c].y : 0) + g * (i_top + jl + c);
    // left boundary column: the init column, or the neighbour's right column
    for (int a = tid; a <= nr; a += blockDim.x)
        leftc[a] = (a == 0) ? (trow ? trow[0].y : 0) + g * (i_top + jl)
                            : (tp.halo ? tp.halo[i_top + a].y : 0) + g * (i_top + a + jl);
    __syncthreads();

    const long long clk0 = clock64();
    const int r = tid;
    const bool rowok = r < nr;
    const int nq = (wd + TT_B - 1) / TT_B;
    const uint8_t b2 = rowok ? p.s2[i_top + r] : 0;
    int left = rowok ? leftc[r + 1] : 0;
    int diag = rowok ? leftc[r] : 0;
    const uint8_t* __restrict__ s1 = p.s1 + jl;                 // letter of tile column c (1-based) = s1[c - 1]
    const int nsteps = nq + nr - 1;
    int upn[TT_B];                      // row 0 reads the top boundary row from global memory: one block ahead
#pragma unroll
    for (int x = 0; x < TT_B; ++x) upn[x] = (r == 0) ? top[min(x + 1, wd)] : 0;
    for (int t = 0; t < nsteps; ++t) {
        const int q = t - r;
        if (rowok && q >= 0 && q < nq) {
            const int c0 = q * TT_B;
            int up[TT_B];
            if (r == 0) {
#pragma unroll
                for (int x = 0; x < TT_B; ++x) {
                    up[x] = upn[x];
                    upn[x] = (NW_TT_DBG & 8) ? 0 : top[min(c0 + TT_B + x + 1, wd)];
                }
            } else {
#pragma unroll
                for (int x = 0; x < TT_B; ++x) up[x] = (NW_TT_DBG & 4) ? x : sm.up[(t - 1) & 1][x][r - 1];
            }
            // the block's TT_B letters as ONE pair of aligned 8-byte loads + a funnel shift (byte loads get scheduled next to
            // their uses and the cell chain then pays a load latency per cell).  Reads up to 15 bytes past the block: the
            // caller pads s1; letters past the tile's last column only feed cells nobody reads.
            uint8_t cs[TT_B];
            {
                static_assert(TT_B <= 8, "one 64-bit word of letters per block");
                const uintptr_t addr = (uintptr_t)(s1 + c0);
                const unsigned long long* al = reinterpret_cast<const unsigned long long*>(addr & ~(uintptr_t)7);
                const unsigned long long lo = (NW_TT_DBG & 2) ? 0ull : al[0], hi = (NW_TT_DBG & 2) ? 0ull : al[1];
                const int sh = (int)(addr & 7) * 8;
                const unsigned long long w = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
#pragma unroll
                for (int x = 0; x < TT_B; ++x) cs[x] = (uint8_t)(w >> (8 * x));
            }
            int32_t* const Ut = U + ((long long)t * rp + r) * TT_B;    // this thread's cells of this phase: TT_B consecutive ints,
                                                                       // consecutive threads next to each other
            // No per-cell bound checks: the ragged last block computes up to TT_B - 1 cells past the tile's last column from
            // clamped letters.  A cell only depends on cells of its own or smaller columns, so what is computed there never
            // reaches a real cell, and the scratch has room for whole blocks.  (With a branch per cell the letter loads could
            // not be hoisted out of it: 157 cycles per cell instead of ~20.)
#pragma unroll
            for (int x = 0; x < TT_B; ++x) {
                const int sub = (cs[x] == b2) ? p.sc_match : p.sc_mis;
                const int h = max(max(diag + sub, up[x] + g), left + g);
                diag = up[x];
                left = h;
                up[x] = h;
            }
            if (!(NW_TT_DBG & 1)) {
                static_assert(TT_B % 4 == 0, "vector stores of the block");
#pragma unroll
                for (int x = 0; x < TT_B; x += 4) *reinterpret_cast<int4*>(Ut + x) = make_int4(up[x], up[x + 1], up[x + 2], up[x + 3]);
            }
#pragma unroll
            for (int x = 0; x < TT_B; ++x)
                if (!(NW_TT_DBG & 4)) sm.up[t & 1][x][r] = up[x];
        }
        if (!(NW_TT_DBG & 16)) __syncthreads();
    }
    // walk back inside the tile
    const long long clk1 = clock64();
    if (tid == 0) { pos[0] = nr; pos[1] = wd; pos[2] = k0; pos[3] = 0; }
    __syncthreads();
    walk_table<WALK_TILE>(TileCells{top, leftc, U, rp}, s1, p.s2 + i_top, sm.walk, pos, p.out1, p.out2, p.sc_match, p.sc_mis, p.sc_gap);
    if (tid == 0) {
        p.state[0] = i_top + pos[0];
        p.state[1] = jl + pos[1];
        p.state[2] = pos[2];
        p.state[4] += 1;
        p.state[5] += (int)((clk1 - clk0) >> 10);
        p.state[6] += (int)((clock64() - clk1) >> 10);
        p.state[7] += nsteps;
    }
}

// integer / DPX pipe rate: 8 independent VIADDMNMX chains per thread (roofline denominator)
constexpr int DPX_PEAK_OPS_PER_ITER = 8;
__global__ void nw_dpx_peak_kernel(int* out, unsigned long long* clk, int iters, int seed)
{
    int x[8], a[8], c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        x[k] = threadIdx.x + k + seed;
        a[k] = (seed & 1) + k - 3;
        c[k] = -(int)threadIdx.x - k;
    }
    unsigned long long t0 = 0, g0 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    }
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = __viaddmax_s32(x[k], a[k], c[k]);
    }
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) {
        unsigned long long t1 = clock64(), g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        clk[2 * blockIdx.x] = t1 - t0;
        clk[2 * blockIdx.x + 1] = g1 - g0;
    }
}

}  // namespace nw
