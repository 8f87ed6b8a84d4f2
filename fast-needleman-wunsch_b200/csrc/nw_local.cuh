// nw_local.cuh -- Smith-Waterman (local alignment) strip kernel: the wavefront fill of nw_kernels.cuh with the clamp at 0.
//
// SURVEY.md 8(f)-4; the reference names Smith-Waterman as a goal (README.md:2) and implements none, so this follows the
// textbook recurrence with the reference's conventions (s1 across, s2 down, linear gap, int32 -- src/serial/serial.cpp:4-36):
//     H[i][j] = max(0, H[i-1][j-1] + (s1[j-1]==s2[i-1] ? match : mismatch), H[i-1][j] + gap, H[i][j-1] + gap),
//     H[0][j] = H[i][0] = 0;   result: the best cell and its position (smallest column, then smallest row).
// The G = H - g(i+j) change of variable of the global kernels does not survive the clamp, so cells are kept in H form and a
// lane carries E = H + gap beside H.  Per cell: ISETP + SEL (substitution score, any byte alphabet),
//     t = VIADDMNMX.RELU(diag, s, E_left)     -- __viaddmax_s32_relu: max(diag + s, left + gap, 0)
//     H = VIADDMNMX(H_up, gap, t)             -- max(up + gap, t); the only op on the row-to-row chain
//     E = H + gap
// plus LEA + VIMNMX per cell for the running best as a key (H * 8 + 7 - row), and four selects per step -- no branch.
// Decomposition, hand-off through tagged boundary rows and the persistent cooperative grid are those of nw_strip_kernel;
// single part only (no halo / right column).  Virtual rows (padding above table row 1) get a hugely negative substitution
// score, so with gap <= 0 they stay 0 like the first row.
#pragma once
#include "nw_kernels.cuh"

namespace nw {

constexpr int LOCAL_NEG = -(1 << 28);

template <int R, bool FULL, bool PRED>
__device__ __forceinline__ void sweep_local(int (&h)[R], int (&e)[R], int& dprev, const uint32_t (&rowb)[R], const int (&wx)[R],
                                            const int wm, const int gap, const uint32_t* __restrict__ Wl,
                                            const int* __restrict__ sin, int* sout, const int lane, const int cb, const int ncols,
                                            int32_t* const (&trow)[FULL ? R : 1], int& best, int& bi, int& bj, const int i_first)
{
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const uint32_t cop = Wl[k];
        int up = __shfl_up_sync(FULL_MASK, h[R - 1], 1);
        if (lane == 0) up = sin[k];
        const int col = cb + k - lane;
        if (!PRED || (col >= 0 && col < ncols)) {
            int diag = dprev;
            dprev = up;
            // running best without a branch (a data-dependent branch inside the unrolled sweep makes the compiler fence every
            // later shuffle): key = H * 8 + (7 - r), so the maximum key is the largest H of the step with the smallest row
            int mk = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int w = (rowb[r] == cop) ? wm : wx[r];
                const int t = __viaddmax_s32_relu(diag, w, e[r]);
                diag = h[r];
                up = __viaddmax_s32(up, gap, t);
                h[r] = up;
                e[r] = up + gap;
                if (FULL) trow[r][col + 1] = up;
                mk = max(mk, up * 8 + (7 - r));
            }
            const int mh = mk >> 3;
            const bool upd = mh > best;       // strictly greater: the first column of a new maximum stays
            best = upd ? mh : best;
            bj = upd ? col + 1 : bj;
            bi = upd ? i_first + 7 - (mk & 7) : bi;
        }
        if (lane == 31) sout[k] = h[R - 1];
    }
}

template <int R, bool FULL>
__device__ __forceinline__ void run_strip_local(const StripParams& p, const int s, const int lane, uint32_t* W, int* sin, int* sout)
{
    constexpr int SH = 32 * R;
    const int ncols = p.ncols;
    const int q0 = s * SH + lane * R;            // first padded row of this lane
    const int i0 = q0 - p.pad_top;               // table row just above this lane's first row (may be <= 0)
    const int gap = p.gap;

    uint32_t rowb[R];
    int wx[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t v = p.rsel[q0 + r];       // raw byte of s2, or 0x200 for a virtual row (generic encoding)
        rowb[r] = v;
        wx[r] = (v == 0x200u) ? LOCAL_NEG : p.w_mis;      // (for this kernel w_match / w_mis carry the plain scores)
    }
    int h[R], e[R];
    int dprev = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) { h[r] = 0; e[r] = gap; }

    int32_t* trow[FULL ? R : 1];
    if (FULL) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = i0 + 1 + r;
            trow[FULL ? r : 0] = (i >= 1) ? p.table + (long long)i * p.tpitch : p.dump;
            trow[FULL ? r : 0][0] = 0;
        }
    }

    int2* tout = p.brow + (long long)s * p.pitch;
    const int2* tin = p.brow + (long long)(s - 1) * p.pitch;
    if (lane == 31) st_tagged_gpu(tout, p.epoch, 0);

    const uint32_t* wq = p.wq;
    W[lane] = wq[lane - 32];
    W[lane + 32] = wq[lane];
    uint32_t wnext = wq[32 + lane];
    const uint32_t* Wl = W + 32 - lane;

    int2 pre = make_int2(0, 0);
    if (s > 0 && lane < ncols) pre = ld_tagged_gpu(tin + lane + 1);
    if (s == 0) sin[lane] = 0;

    int best = 0, bi = 0, bj = 0;
    const int nblocks = (ncols + 31 + 31) >> 5;
    for (int b = 0; b < nblocks; ++b) {
        const int cb = b << 5;
        if (b > 0) {
            W[lane] = W[lane + 32];
            W[lane + 32] = wnext;
            int nc = cb + 32 + lane;
            wnext = wq[nc < ncols + WQ_PAD ? nc : ncols + WQ_PAD - 1];
        }
        if (s > 0 && cb < ncols) {
            const int col = cb + lane;
            const bool need = col < ncols;
            SpinGuard sg;
            while (!__all_sync(FULL_MASK, !need || pre.x == p.epoch)) {
                __nanosleep(100);
                if (need && pre.x != p.epoch) pre = ld_tagged_gpu(tin + col + 1);
                if (sg.expired_warp(p)) break;
            }
            sin[lane] = pre.y;
            if (col + 32 < ncols) pre = ld_tagged_gpu(tin + col + 33);
        }
        __syncwarp();
        if (cb >= 31 && cb + 31 < ncols)
            sweep_local<R, FULL, false>(h, e, dprev, rowb, wx, p.w_match, gap, Wl, sin, sout, lane, cb, ncols, trow, best, bi, bj, i0 + 1);
        else
            sweep_local<R, FULL, true>(h, e, dprev, rowb, wx, p.w_match, gap, Wl, sin, sout, lane, cb, ncols, trow, best, bi, bj, i0 + 1);
        __syncwarp();
        const int oc = cb - 31 + lane;               // column finished by lane 31 at step k = lane of this block
        const int ov = sout[lane];
        if (oc >= 0 && oc < ncols) st_tagged_gpu(tout + oc + 1, p.epoch, ov);
    }

    // best cell of the strip: highest score, then smallest column, then smallest row
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const int ob = __shfl_xor_sync(FULL_MASK, best, d), oi = __shfl_xor_sync(FULL_MASK, bi, d), oj = __shfl_xor_sync(FULL_MASK, bj, d);
        if (ob > best || (ob == best && (oj < bj || (oj == bj && oi < bi)))) { best = ob; bi = oi; bj = oj; }
    }
    if (lane == 0) p.local_best[s] = make_int4(best, bi, bj, 0);
    __syncwarp();
}

template <int R, bool FULL>
__global__ void __launch_bounds__(512) nw_local_kernel(const StripParams p)
{
    extern __shared__ uint32_t nw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t* W = nw_smem + warp * SMEM_WORDS_PER_WARP;
    int* sin = (int*)(W + 64);
    int* sout = sin + 32;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    for (int s = slot; s < p.nstrips; s += nslots) run_strip_local<R, FULL>(p, s, lane, W, sin, sout);
}

// best over the strips (same order) -> out[0..2] = score, i, j; a best score of 0 reports (0, 0)
__global__ void nw_local_finish_kernel(const int4* __restrict__ per_strip, int nstrips, int32_t* out)
{
    int best = 0, bi = 0, bj = 0;
    for (int s = threadIdx.x; s < nstrips; s += 32) {
        const int4 v = per_strip[s];
        if (v.x > best || (v.x == best && (v.z < bj || (v.z == bj && v.y < bi)))) { best = v.x; bi = v.y; bj = v.z; }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const int ob = __shfl_xor_sync(FULL_MASK, best, d), oi = __shfl_xor_sync(FULL_MASK, bi, d), oj = __shfl_xor_sync(FULL_MASK, bj, d);
        if (ob > best || (ob == best && (oj < bj || (oj == bj && oi < bi)))) { best = ob; bi = oi; bj = oj; }
    }
    if (threadIdx.x == 0) {
        out[0] = best;
        out[1] = best > 0 ? bi : 0;
        out[2] = best > 0 ? bj : 0;
    }
}

}  // namespace nw
