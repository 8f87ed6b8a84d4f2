// nw_cuda.cu -- libnw_cuda.so: the C ABI declared in include/nw_cuda.h over the sm_100a kernels of nw_kernels.cuh.
//
// Replaces, for the reference's callers, the body of
//     void needlemanWunsch(dnaArray s1, dnaArray s2, int* t)            (reference: src/serial/serial.cpp:4-36)
// and, for multi-GPU runs, the column-strip pipeline of src/mpi/mpi-vert.cpp:17-105 / mpi-vert-driver.cpp:35-38.
// There is no CPU fallback in this file: every compute entry point needs a CUDA device and fails with NW_ERR_CUDA
// (message in nw_cuda_last_error()) when there is none.
#include "../../include/nw_cuda.h"
#include "nw_kernels.cuh"
#include "nw_packed.cuh"
#include "nw_batch.cuh"
#include "nw_lag2.cuh"
#include "nw_ws.cuh"
#include "nw_local.cuh"

#include <cooperative_groups.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <emmintrin.h>
#include <new>
#include <string>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(NW_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

struct OccEntry {
    const void* fn;
    int threads;
    size_t smem;
    int per_sm;
};
struct DeviceState {
    bool inited = false, warmed = false;
    int sm_count = 0;
    cudaDeviceProp prop;
    // Every device buffer of a plan comes from this pool (cudaMallocFromPoolAsync).  Its release threshold is "never",
    // and nw_cuda_init pre-faults it, so a one-shot call -- which the reference driver times as a whole
    // (src/common/driver.cpp:26-30) -- sub-allocates in microseconds instead of paying cudaMalloc / cudaFree.
    cudaMemPool_t pool = nullptr;
    // One mapped pinned word per device: a strip kernel whose wait for a predecessor (another warp, another GPU, another
    // process) exceeds NW_CUDA_SPIN_TIMEOUT_MS sets it and gives up instead of hanging; the host then fails the call.
    volatile int* h_abort = nullptr;
    int* d_abort = nullptr;        // device view of h_abort
    int* d_abort_dev = nullptr;    // device-memory copy: what waiting warps re-read
    std::vector<OccEntry> occ;      // cudaOccupancyMaxActiveBlocksPerMultiprocessor results (and smem opt-in done)
};
std::mutex g_mu;
DeviceState g_dev[64];

int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// stream-ordered allocation from the device's pool / release back into it
cudaError_t dev_alloc(int device, cudaStream_t st, void** ptr, size_t bytes)
{
    return cudaMallocFromPoolAsync(ptr, bytes ? bytes : 1, g_dev[device].pool, st);
}
template <class T>
cudaError_t dev_alloc(int device, cudaStream_t st, T** ptr, size_t bytes)
{
    return dev_alloc(device, st, (void**)ptr, bytes);
}

// resident CTAs per SM of `fn` at this launch shape; raises the dynamic shared memory limit when needed.  Cached: a
// one-shot call must not repeat driver queries inside the reference driver's timed region.
int occupancy(int device, const void* fn, int threads, size_t smem, int* per_sm)
{
    DeviceState& d = g_dev[device];
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (const OccEntry& e : d.occ)
            if (e.fn == fn && e.threads == threads && e.smem == smem) {
                *per_sm = e.per_sm;
                return NW_OK;
            }
    }
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int n = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, threads, smem));
    std::lock_guard<std::mutex> lk(g_mu);
    d.occ.push_back({fn, threads, smem, n});
    *per_sm = n;
    return NW_OK;
}

// ---- kernel dispatch ---------------------------------------------------------------------------------------------
typedef void (*StripKernel)(const nw::StripParams);

template <int R>
StripKernel strip_kernel_r(bool generic, bool full)
{
    if (generic) return full ? nw::nw_strip_kernel<R, true, true> : nw::nw_strip_kernel<R, true, false>;
    return full ? nw::nw_strip_kernel<R, false, true> : nw::nw_strip_kernel<R, false, false>;
}
StripKernel strip_kernel(int R, bool generic, bool full)
{
    switch (R) {
    case 1: return strip_kernel_r<1>(generic, full);
    case 2: return strip_kernel_r<2>(generic, full);
    case 4: return strip_kernel_r<4>(generic, full);
    case 8: return strip_kernel_r<8>(generic, full);
    default: return nullptr;
    }
}

StripKernel strip16_kernel(int regs)
{
    switch (regs) {
    case 1: return nw::nw_strip16_kernel<1>;
    case 2: return nw::nw_strip16_kernel<2>;
    case 4: return nw::nw_strip16_kernel<4>;
    case 8: return nw::nw_strip16_kernel<8>;
    default: return nullptr;
    }
}

StripKernel strip16l2_kernel(int regs, bool stair = false)
{
    switch (regs) {
    case 1: return stair ? nw::nw_strip16l2_kernel<1, true> : nw::nw_strip16l2_kernel<1, false>;
    case 2: return stair ? nw::nw_strip16l2_kernel<2, true> : nw::nw_strip16l2_kernel<2, false>;
    case 4: return stair ? nw::nw_strip16l2_kernel<4, true> : nw::nw_strip16l2_kernel<4, false>;
    case 8: return stair ? nw::nw_strip16l2_kernel<8, true> : nw::nw_strip16l2_kernel<8, false>;
    default: return nullptr;
    }
}

StripKernel strip16ws_kernel(int regs)
{
    switch (regs) {
    case 1: return nw::nw_strip16ws_kernel<1>;
    case 2: return nw::nw_strip16ws_kernel<2>;
    case 4: return nw::nw_strip16ws_kernel<4>;
    case 8: return nw::nw_strip16ws_kernel<8>;
    default: return nullptr;
    }
}

StripKernel local_kernel(int R, bool full)
{
    switch (R) {
    case 1: return full ? nw::nw_local_kernel<1, true> : nw::nw_local_kernel<1, false>;
    case 2: return full ? nw::nw_local_kernel<2, true> : nw::nw_local_kernel<2, false>;
    case 4: return full ? nw::nw_local_kernel<4, true> : nw::nw_local_kernel<4, false>;
    case 8: return full ? nw::nw_local_kernel<8, true> : nw::nw_local_kernel<8, false>;
    default: return nullptr;
    }
}

StripKernel full16_kernel(int regs)
{
    switch (regs) {
    case 1: return nw::nw_full16_kernel<1>;
    case 2: return nw::nw_full16_kernel<2>;
    case 4: return nw::nw_full16_kernel<4>;
    case 8: return nw::nw_full16_kernel<8>;
    default: return nullptr;
    }
}
int full16_smem_words(int regs)
{
    switch (regs) {
    case 1: return (nw::SMEM16F_WORDS_PER_WARP(1) + 3) & ~3;
    case 2: return (nw::SMEM16F_WORDS_PER_WARP(2) + 3) & ~3;
    case 4: return (nw::SMEM16F_WORDS_PER_WARP(4) + 3) & ~3;
    default: return (nw::SMEM16F_WORDS_PER_WARP(8) + 3) & ~3;
    }
}

typedef void (*BatchKernel)(const nw::BatchParams);
BatchKernel batch_kernel(int R, bool generic)
{
    switch (R) {
    case 4: return generic ? nw::nw_batch_kernel<4, true> : nw::nw_batch_kernel<4, false>;
    case 8: return generic ? nw::nw_batch_kernel<8, true> : nw::nw_batch_kernel<8, false>;
    case 16: return generic ? nw::nw_batch_kernel<16, true> : nw::nw_batch_kernel<16, false>;
    case 32: return generic ? nw::nw_batch_kernel<32, true> : nw::nw_batch_kernel<32, false>;
    default: return nullptr;
    }
}

BatchKernel batch16_kernel(int regs)
{
    switch (regs) {
    case 2: return nw::nw_batch16_kernel<2>;
    case 4: return nw::nw_batch16_kernel<4>;
    case 8: return nw::nw_batch16_kernel<8>;
    case 16: return nw::nw_batch16_kernel<16>;
    default: return nullptr;
    }
}

// byte value -> 0..3 when at most four distinct byte values occur; returns false otherwise (generic path)
bool build_code(const bool seen[256], uint8_t code[256])
{
    int n = 0;
    memset(code, 0, 256);
    for (int v = 0; v < 256; ++v)
        if (seen[v]) {
            if (n == 4) return false;
            code[v] = (uint8_t)n++;
        }
    return true;
}

void bitmap_to_seen(const uint32_t bm[8], bool seen[256])
{
    for (int v = 0; v < 256; ++v) seen[v] = (bm[v >> 5] >> (v & 31)) & 1u;
}

// validated copy of an nw_scoring (NULL = the reference's macros, src/common/needleman-wunsch.hpp:11-13)
struct Scoring {
    int match = 1, mismatch = 0, gap = -1, local = 0;
    bool operator==(const Scoring& o) const { return match == o.match && mismatch == o.mismatch && gap == o.gap && local == o.local; }
};
int parse_scoring(const nw_scoring* in, int32_t n1, int32_t n2, Scoring* out)
{
    Scoring sc;
    if (in) {
        for (int k = 0; k < 4; ++k)
            if (in->reserved[k] != 0) return fail(NW_ERR_ARG, "nw_scoring.reserved must be zero");
        if (in->local != 0 && in->local != 1) return fail(NW_ERR_ARG, "nw_scoring.local must be 0 or 1 (got %d)", in->local);
        sc.match = in->match; sc.mismatch = in->mismatch; sc.gap = in->gap; sc.local = in->local;
    }
    if (sc.local && sc.gap > 0) return fail(NW_ERR_ARG, "local alignment needs gap <= 0 (got %d)", sc.gap);
    if (sc.local && (long long)std::max(sc.match, 0) * std::min(n1, n2) >= (1LL << 27))     // (the kernel packs H * 8 + row)
        return fail(NW_ERR_ARG, "scores (%d, %d, %d) overflow the local kernel's best-cell key on a %d x %d table", sc.match,
                    sc.mismatch, sc.gap, n1, n2);
    long long a = std::max({std::llabs((long long)sc.match), std::llabs((long long)sc.mismatch), std::llabs((long long)sc.gap)});
    a = std::max(a, 2 * std::llabs((long long)sc.gap) + std::max(std::llabs((long long)sc.match), std::llabs((long long)sc.mismatch)));
    if (a * ((long long)n1 + n2 + 2) >= (1LL << 30))
        return fail(NW_ERR_ARG, "scores (%d, %d, %d) overflow int32 on a %d x %d table", sc.match, sc.mismatch, sc.gap, n1, n2);
    *out = sc;
    return NW_OK;
}

}  // namespace

// =====================================================================================================================
// plan
// =====================================================================================================================
struct nw_plan {
    int device = 0;
    int n1 = 0, n2 = 0, mode = 0, part = 0, nparts = 1;
    int jstart = 0;       // global table column of this part's left boundary column
    int ncols = 0;        // interior columns of this part
    int R = 4;            // table rows per lane: 1, 2, 4, 8 (32-bit kernels) or 2, 4, 8, 16 (packed kernel)
    int R_req = 0, warps_req = 0, ctas_req = 0;      // what the caller asked for (0 = automatic)
    int warps = 8, ctas = 0, nstrips = 0, pad_top = 0;
    bool packed = false;  // nw_packed.cuh kernel (boundary mode, at most four distinct byte values)
    bool lag2 = false;    // nw_lag2.cuh: virtual lanes two columns apart (boundary mode; the default packed kernel)
    bool ws = false;      // nw_ws.cuh: the lag-2 sweep with a helper warp per strip (compute warp + helper warp per scheduler)
    int threads = 0;      // threads per CTA of the strip kernel (warps * 32, or warps * 64 with helper warps)
    bool generic = false, uploaded = false;
    size_t rsel_words = 0, brow_words = 0;
    int epoch = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    // device memory
    uint8_t *d_s1 = nullptr, *d_s2 = nullptr;
    uint32_t *d_wq = nullptr, *d_rsel = nullptr, *d_bitmap = nullptr;
    int2* d_brow = nullptr;                    // allocation; the kernels see brow() = d_brow + 1, so that the word of
    int2* brow() const { return d_brow + 1; }  // table column j = 2x (interior column 2x - 1) is 16-byte aligned
    long long pitch = 0;
    unsigned long long* d_times = nullptr;     // nstrips x 2 %globaltimer stamps of the most recent fill
    size_t times_strips = 0;
    uint8_t* d_rev = nullptr;                  // score mode, bottom half: staging of the sequences before reversal
    int4* d_local_best = nullptr;              // local alignment: best cell per strip
    size_t local_best_n = 0;

    bool mailbox_pooled = false;     // the mailbox comes from the device's pool (same-device pipelines: no IPC, no peers)
    int2* d_mailbox = nullptr;       // 2 x mpitch tagged words: halo of parts > 0 (double-buffered by epoch parity)
    int2* d_rcol_local = nullptr;    // 2 x mpitch: right column when nobody is connected on the right
    long long mpitch = 0;
    int2* rcol_target = nullptr;     // where the right column goes: d_rcol_local, or the right neighbour's mailbox
    bool rcol_peer = false, halo_peer = false;
    // the 64 bytes after a mailbox's two buffers hold its "consumed epoch" word, written by the consumer's finish
    // kernel and polled by the producer (over NVLink when the producer is another GPU)
    void* ipc_mailbox = nullptr;     // imported peer mappings (closed on destroy)
    int32_t *d_table = nullptr, *d_dump = nullptr;
    long long tpitch = 0;
    int32_t *d_last_row = nullptr, *d_last_col = nullptr, *d_score = nullptr, *d_tmp_row = nullptr;
    size_t smem = 0;
    StripKernel kernel = nullptr;
    // packed full-table mode: pass 1 (kernel above, with snapshots) + pass 2 (tile replay with coalesced table stores)
    StripKernel kernel2 = nullptr;
    uint32_t* d_snap = nullptr;
    size_t snap_words = 0, smem2 = 0;
    int tile_blocks = 32, ntiles = 1, ctas2 = 0, warps2 = 5;
    // streamed table delivery (one-shot calls with a HOST table): the device never holds the whole table; pass 2 runs
    // band by band into a two-band ring while the previous band is on its way to the host
    bool want_streamed = false, streamed = false;
    int band_strips = 0;
    size_t table_elems = 0;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t band_ev[2] = {nullptr, nullptr};
    nw::StripParams last_sp;         // parameters of the most recent fill (pass 2 of a streamed delivery reuses them)
    // NW_MODE_SCORE: this plan only owns two boundary-mode sub-plans (top half forwards, bottom half reversed), a join
    // event and the combined score
    nw_plan* sub[2] = {nullptr, nullptr};
    int split = 0;                   // rows of the top half
    bool swapped = false;            // score mode sweeps along the SHORTER sequence (the score is symmetric in s1, s2)
    cudaEvent_t join_ev = nullptr;
    // score mode along a staircase: both sub-plans span all rows; strip s of the forward one sweeps
    // widths[s] columns, strip t of the reversed one the rest.  stair_half: 0 = not a half, 1 = forward, 2 = reversed (padding
    // rows at the bottom, so that both halves share their strip boundaries)
    uint8_t code[256];               // letter codes of the most recent upload (four-letter paths)
    int stair_half = 0;
    int dev2 = -1;                   // (on the parent) device of the second half: the same one, or a second GPU
    bool stair = false;              // (on the parent) the sub-plans were built for the staircase
    int* d_widths = nullptr;
    uint32_t* d_tails = nullptr;
    size_t widths_n = 0;
    Scoring scoring_of() const { Scoring sc; sc.match = sc_match; sc.mismatch = sc_mis; sc.gap = sc_gap; sc.local = local ? 1 : 0; return sc; }
    // scoring (nw_scoring; default = the reference's macros, src/common/needleman-wunsch.hpp:11-13)
    int sc_match = 1, sc_mis = 0, sc_gap = -1;
    bool local = false;              // Smith-Waterman (nw_local.cuh)
    int w_match() const { return std::max(sc_match - 2 * sc_gap, 0); }     // G-form weights (nw_kernels.cuh)
    int w_mis() const { return std::max(sc_mis - 2 * sc_gap, 0); }
    int w_max() const { return std::max(w_match(), w_mis()); }
};

static int ensure_device(int device)
{
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n || device >= 64) return fail(NW_ERR_ARG, "device %d out of range (count %d)", device, n);
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceState& d = g_dev[device];
    if (!d.inited) {
        CK(cudaSetDevice(device));
        CK(cudaFree(0));
        CK(cudaGetDeviceProperties(&d.prop, device));
        d.sm_count = d.prop.multiProcessorCount;
        if (d.prop.major < 10)
            return fail(NW_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                        d.prop.major, d.prop.minor);
        cudaMemPoolProps pp;
        memset(&pp, 0, sizeof pp);
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = device;
        CK(cudaMemPoolCreate(&d.pool, &pp));
        unsigned long long never = ~0ULL;
        CK(cudaMemPoolSetAttribute(d.pool, cudaMemPoolAttrReleaseThreshold, &never));
        CK(cudaHostAlloc((void**)&d.h_abort, 64, cudaHostAllocMapped | cudaHostAllocPortable));
        *d.h_abort = 0;
        CK(cudaHostGetDevicePointer((void**)&d.d_abort, (void*)d.h_abort, 0));
        CK(cudaMalloc(&d.d_abort_dev, 64));
        CK(cudaMemset(d.d_abort_dev, 0, 64));
        d.inited = true;
    }
    return NW_OK;
}

extern "C" const char* nw_cuda_version(void) { return "nw_cuda 0.1 (sm_100a)"; }
extern "C" const char* nw_cuda_last_error(void) { return g_err; }

extern "C" int nw_cuda_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(NW_ERR_CUDA, "cudaGetDeviceCount -> %s", cudaGetErrorString(e));
    return n;
}

extern "C" int nw_cuda_device_info(int device, char* name, int name_len, int* sm_count, int* sm_clock_mhz)
{
    int rc = ensure_device(device);
    if (rc) return rc;
    const DeviceState& d = g_dev[device];
    if (name && name_len > 0) snprintf(name, name_len, "%s", d.prop.name);
    if (sm_count) *sm_count = d.sm_count;
    if (sm_clock_mhz) {
        int khz = 0;
        CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
        *sm_clock_mhz = khz / 1000;
    }
    return NW_OK;
}

// ---- geometry -------------------------------------------------------------------------------------------------------
// mpi-vert partition (src/mpi/mpi-vert-driver.cpp:35-36, src/mpi/mpi-vert.cpp:17):
//   q = (n1+1)/P, start = q*p - (p>0), table columns owned = q + (p>0) + (p==P-1 ? (n1+1)%P : 0), first one = boundary.
static void partition(int n1, int P, int p, int* jstart, int* ncols)
{
    const long long q = ((long long)n1 + 1) / P;
    const long long start = q * p - (p > 0);
    const long long owned = q + (p > 0) + ((p == P - 1) ? (((long long)n1 + 1) % P) : 0);
    *jstart = (int)start;
    *ncols = (int)(owned - 1);
}

static int choose_rows_per_lane(int n2, int ncols, int sm_count)
{
    // One warp keeps an SMSP's integer pipe busy once R >= ~4; below that the per-column overhead (shuffle, operand
    // load) dominates, above it the wavefront has too few strips.  Strip start-up lag is ~48 columns per strip.
    const int cand[4] = {8, 4, 2, 1};
    double best = 1e300;
    int bestR = 4;
    const double slots = sm_count * 8.0;
    for (int ci = 0; ci < 4; ++ci) {
        const int R = cand[ci];
        const double strips = (n2 + 32.0 * R - 1) / (32.0 * R);
        const double tcol = 3.0 * R + 7.0;                                   // issue slots per column per warp
        const double per_smsp = std::max(1.0, strips / (sm_count * 4.0));    // warps sharing one scheduler
        const double rounds = std::max(1.0, strips / slots);
        const double conc = std::min(strips, slots);
        const double t = (ncols * rounds + conc * 48.0) * std::max(tcol * std::min(per_smsp, 2.0), 4.0 * R + 30.0);
        if (t < best) { best = t; bestR = R; }
    }
    return bestR;
}

static int choose_rows_per_lane_packed(int n2, int ncols, int sm_count)
{
    // packed kernel: rows per lane = 2 * registers.  One warp per scheduler is enough; a strip lags its predecessor
    // by ~64 columns of skew plus the hand-off latency.
    const int cand[4] = {16, 8, 4, 2};
    double best = 1e300;
    int bestR = 8;
    const double slots = sm_count * 4.0;
    for (int ci = 0; ci < 4; ++ci) {
        const int R = cand[ci];
        const double strips = (n2 + 32.0 * R - 1) / (32.0 * R);
        const double issue = 2.0 * (1.5 * R + 5.0);                       // cycles per step, one warp
        const double chain = 27.0 + 2.0 * (R / 2) ;                       // shuffle + PRMT + max chain
        const double rounds = std::max(1.0, strips / slots);
        const double conc = std::min(strips, slots);
        const double t = (ncols * rounds + conc * 110.0) * std::max(issue, chain);
        if (t < best) { best = t; bestR = R; }
    }
    return bestR;
}

extern "C" int nw_plan_destroy(nw_plan* p)
{
    if (!p) return NW_OK;
    cudaSetDevice(p->device);
    if (p->sub[0]) nw_plan_destroy(p->sub[0]);
    if (p->sub[1]) nw_plan_destroy(p->sub[1]);
    if (p->join_ev) cudaEventDestroy(p->join_ev);
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->ipc_mailbox) cudaIpcCloseMemHandle(p->ipc_mailbox);
    void* bufs[] = {p->d_s1, p->d_s2, p->d_wq, p->d_rsel, p->d_bitmap, p->d_brow, p->d_rcol_local, p->d_table,
                    p->d_dump, p->d_snap, p->d_last_row, p->d_last_col, p->d_score, p->d_tmp_row, p->d_rev, p->d_times,
                    p->d_local_best, p->d_widths, p->d_tails};
    for (void* b : bufs)
        if (b) cudaFreeAsync(b, p->stream ? p->stream : (cudaStream_t)0);      // back into the device's pool
    if (p->d_mailbox && p->mailbox_pooled) cudaFreeAsync(p->d_mailbox, p->stream ? p->stream : (cudaStream_t)0);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->d_mailbox && !p->mailbox_pooled) cudaFree(p->d_mailbox);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->ev2) cudaEventDestroy(p->ev2);
    if (p->ev3) cudaEventDestroy(p->ev3);
    for (int i = 0; i < 2; ++i)
        if (p->band_ev[i]) cudaEventDestroy(p->band_ev[i]);
    if (p->stream2) cudaStreamDestroy(p->stream2);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return NW_OK;
}

static int plan_alloc(nw_plan* p, const nw_tuning* tuning)
{
    CK(cudaSetDevice(p->device));
    CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&p->ev0));
    CK(cudaEventCreate(&p->ev1));
    CK(cudaEventCreate(&p->ev2));
    CK(cudaEventCreate(&p->ev3));

    p->R_req = tuning ? tuning->rows_per_lane : 0;
    if (p->R_req == 0) p->R_req = env_int("NW_CUDA_R", 0);
    p->warps_req = tuning ? tuning->warps_per_cta : 0;
    if (p->warps_req == 0) p->warps_req = env_int("NW_CUDA_WARPS", 0);
    p->ctas_req = tuning ? tuning->ctas : 0;
    if (p->ctas_req == 0) p->ctas_req = env_int("NW_CUDA_CTAS", 0);
    if (p->R_req != 0 && p->R_req != 1 && p->R_req != 2 && p->R_req != 4 && p->R_req != 8 && p->R_req != 16)
        return fail(NW_ERR_ARG, "rows_per_lane must be 0 (auto), 1, 2, 4, 8 or 16 (got %d)", p->R_req);
    if (p->warps_req < 0 || p->warps_req > 16) return fail(NW_ERR_ARG, "warps_per_cta must be in 0..16 (got %d)", p->warps_req);

    const int nc = p->ncols, n2 = p->n2;
    CK(dev_alloc(p->device, p->stream, &p->d_s1, (size_t)std::max(nc, 1)));
    CK(dev_alloc(p->device, p->stream, &p->d_s2, (size_t)std::max(n2, 1)));
    CK(dev_alloc(p->device, p->stream, &p->d_wq, sizeof(uint32_t) * ((size_t)nc + nw::WQ_PAD + nw::WQ_PADR)));
    CK(dev_alloc(p->device, p->stream, &p->d_bitmap, 8 * sizeof(uint32_t)));
    // 48 spare words behind a row: the lag-2 kernel's out-of-range stores and its 32-column top-row prefetches
    p->pitch = ((long long)nc + 1 + 48 + 15) & ~15LL;
    p->mpitch = ((long long)n2 + 1 + 15) & ~15LL;
    if (p->part > 0) {
        // cudaMalloc, not the pool: the mailbox is exported to other processes as a CUDA IPC handle and written by peers
        if (p->mailbox_pooled) CK(dev_alloc(p->device, p->stream, &p->d_mailbox, sizeof(int2) * 2 * (size_t)p->mpitch + 64));
        else CK(cudaMalloc(&p->d_mailbox, sizeof(int2) * 2 * (size_t)p->mpitch + 64));
        CK(cudaMemsetAsync(p->d_mailbox, 0, sizeof(int2) * 2 * (size_t)p->mpitch + 64, p->stream));
    }
    CK(dev_alloc(p->device, p->stream, &p->d_rcol_local, sizeof(int2) * 2 * (size_t)p->mpitch));
    CK(cudaMemsetAsync(p->d_rcol_local, 0, sizeof(int2) * 2 * (size_t)p->mpitch, p->stream));
    p->rcol_target = p->d_rcol_local;
    if (p->mode == NW_MODE_FULL) {
        p->tpitch = ((long long)nc + 1 + 7) & ~7LL;      // rows start on a 32-byte sector: the table stores need it
        CK(dev_alloc(p->device, p->stream, &p->d_dump, sizeof(int32_t) * (size_t)p->tpitch));      // the table itself: plan_pick_kernel
    }
    CK(dev_alloc(p->device, p->stream, &p->d_last_row, sizeof(int32_t) * ((size_t)nc + 1)));
    CK(dev_alloc(p->device, p->stream, &p->d_last_col, sizeof(int32_t) * ((size_t)n2 + 1)));
    CK(dev_alloc(p->device, p->stream, &p->d_tmp_row, sizeof(int32_t) * ((size_t)nc + 1)));
    CK(dev_alloc(p->device, p->stream, &p->d_score, 64));
    // the mailbox (tags and ack word) must be zero before it is exported, connected or polled: producers run on other
    // streams, devices or processes, which nothing else orders after the clears above
    if (p->part > 0) CK(cudaStreamSynchronize(p->stream));
    return NW_OK;
}

// Strip geometry and kernel choice; needs the alphabet, so it runs at upload time.  (Re)allocates what depends on it.
static int plan_pick_kernel(nw_plan* p)
{
    const DeviceState& d = g_dev[p->device];
    CK(cudaSetDevice(p->device));
    // packed s16x2 kernels: four letters, small weights (the 16-bit window and its re-basing margin are sized for them),
    // and -- full-table mode, whose pass 2 emits H = G - i - j -- the reference's gap of -1
    p->packed = !p->generic && !p->local && !env_int("NW_CUDA_NO_PACKED", 0) && p->R_req != 1 && p->w_max() <= 16 &&
                (p->mode == NW_MODE_BOUNDARY || (!env_int("NW_CUDA_NO_PACKED_FULL", 0) && p->sc_gap == -1));
    int R = p->R_req;
    if (R == 0) R = p->packed ? choose_rows_per_lane_packed(p->n2, p->ncols, d.sm_count)
                              : choose_rows_per_lane(p->n2, p->ncols, d.sm_count);
    if (!p->packed && R > 8) R = 8;
    p->R = R;
    p->warps = p->warps_req ? p->warps_req : (p->packed ? 4 : 8);
    p->nstrips = (int)(((long long)p->n2 + 32LL * R - 1) / (32LL * R));
    p->pad_top = (p->stair_half == 2) ? 0 : p->nstrips * 32 * R - p->n2;
    const size_t rsel_words = (size_t)std::max(p->nstrips * 32 * R, 2);      // (packed: R/2 registers x 2 words each)
    const size_t brow_words = (size_t)p->pitch * (size_t)std::max(p->nstrips, 1) + 2;
    if (rsel_words > p->rsel_words) {
        if (p->d_rsel) CK(cudaFreeAsync(p->d_rsel, p->stream));
        p->d_rsel = nullptr;
        CK(dev_alloc(p->device, p->stream, &p->d_rsel, sizeof(uint32_t) * rsel_words));
        p->rsel_words = rsel_words;
    }
    if ((size_t)std::max(p->nstrips, 1) > p->times_strips) {
        if (p->d_times) CK(cudaFreeAsync(p->d_times, p->stream));
        p->d_times = nullptr;
        p->times_strips = (size_t)std::max(p->nstrips, 1);
        CK(dev_alloc(p->device, p->stream, &p->d_times, sizeof(unsigned long long) * 4 * p->times_strips));
    }
    if (brow_words > p->brow_words) {
        if (p->d_brow) CK(cudaFreeAsync(p->d_brow, p->stream));
        p->d_brow = nullptr;
        CK(dev_alloc(p->device, p->stream, &p->d_brow, sizeof(int2) * brow_words));
        CK(cudaMemsetAsync(p->d_brow, 0, sizeof(int2) * brow_words, p->stream));    // tags must not match any epoch
        p->brow_words = brow_words;
    }
    // boundary mode: the lag-2 schedule (nw_lag2.cuh); NW_CUDA_LAG2=0 selects the one-column skew of nw_packed.cuh, which
    // full-table mode always uses (its pass 2 replays tiles with the same schedule)
    p->lag2 = p->packed && p->mode == NW_MODE_BOUNDARY && env_int("NW_CUDA_LAG2", 1) != 0;
    p->ws = p->lag2 && env_int("NW_CUDA_WS", 0) != 0 && p->warps <= 8 && p->stair_half == 0;   // opt-in: measured slower in a chain
    p->threads = p->warps * (p->ws ? 64 : 32);
    if (p->ws) {
        p->kernel = strip16ws_kernel(R / 2);
        p->smem = sizeof(uint32_t) * nw::WS_SMEM_WORDS_PER_PAIR * (size_t)p->warps;
    } else if (p->lag2) {
        p->kernel = strip16l2_kernel(R / 2, p->stair_half != 0);
        p->smem = sizeof(uint32_t) * nw::L2_SMEM_WORDS_PER_WARP * (size_t)p->warps;
    } else if (p->packed) {
        p->kernel = strip16_kernel(R / 2);
        p->smem = sizeof(uint32_t) * nw::SMEM16_WORDS_PER_WARP * (size_t)p->warps;
    } else if (p->local) {
        p->kernel = local_kernel(R, p->mode == NW_MODE_FULL);
        p->smem = sizeof(uint32_t) * nw::SMEM_WORDS_PER_WARP * (size_t)p->warps;
        if ((size_t)std::max(p->nstrips, 1) > p->local_best_n) {
            if (p->d_local_best) CK(cudaFreeAsync(p->d_local_best, p->stream));
            p->d_local_best = nullptr;
            p->local_best_n = (size_t)std::max(p->nstrips, 1);
            CK(dev_alloc(p->device, p->stream, &p->d_local_best, sizeof(int4) * p->local_best_n));
        }
    } else {
        p->kernel = strip_kernel(R, p->generic, p->mode == NW_MODE_FULL);
        p->smem = sizeof(uint32_t) * nw::SMEM_WORDS_PER_WARP * (size_t)p->warps;
    }
    if (!p->kernel) return fail(NW_ERR_ARG, "no kernel for rows_per_lane=%d (%s)", R, p->packed ? "packed" : "32-bit");
    p->kernel2 = nullptr;
    if (p->packed && p->mode == NW_MODE_FULL) {
        const int regs = R / 2;
        const int nblocks = (p->ncols + 63 + 31) >> 5;
        p->tile_blocks = std::max(2, env_int("NW_CUDA_TILE_BLOCKS", 32));
        p->ntiles = std::max(1, (nblocks + p->tile_blocks - 1) / p->tile_blocks);
        const size_t need = (size_t)std::max(p->nstrips, 1) * (size_t)p->ntiles * 32u * (size_t)(regs + 2);
        if (need > p->snap_words) {
            if (p->d_snap) CK(cudaFreeAsync(p->d_snap, p->stream));
            p->d_snap = nullptr;
            CK(dev_alloc(p->device, p->stream, &p->d_snap, sizeof(uint32_t) * need));
            p->snap_words = need;
        }
        p->kernel2 = full16_kernel(regs);
        p->warps2 = std::max(1, std::min(8, env_int("NW_CUDA_FULL_WARPS", 5)));
        p->smem2 = sizeof(uint32_t) * (size_t)full16_smem_words(regs) * (size_t)p->warps2;
        int per_sm2 = 0;
        {
            const int rc2 = occupancy(p->device, (const void*)p->kernel2, p->warps2 * 32, p->smem2, &per_sm2);
            if (rc2) return rc2;
        }
        if (per_sm2 < 1) return fail(NW_ERR_CUDA, "full-table pass-2 kernel does not fit on an SM");
        const long long ntasks = (long long)p->nstrips * p->ntiles;
        p->ctas2 = (int)std::max<long long>(1, std::min<long long>((ntasks + p->warps2 - 1) / p->warps2,
                                                                     (long long)d.sm_count * per_sm2));
    }
    if (p->mode == NW_MODE_FULL) {
        // whole table, or (streamed delivery, packed kernels only) a ring of two bands of ~256 MB
        p->streamed = p->want_streamed && p->kernel2 != nullptr && p->nparts == 1 && p->ncols > 0 && p->n2 > 0 &&
                      !env_int("NW_CUDA_NO_STREAMED", 0);
        size_t elems = (size_t)p->tpitch * ((size_t)p->n2 + 1);
        if (p->streamed) {
            const size_t band_bytes = (size_t)std::max(16, env_int("NW_CUDA_BAND_MB", 256)) << 20;
            const size_t strip_bytes = sizeof(int32_t) * (size_t)p->tpitch * 32u * (size_t)R;
            p->band_strips = (int)std::max<size_t>(1, band_bytes / strip_bytes);
            if (p->band_strips >= p->nstrips) p->streamed = false;          // one band would hold everything anyway
            else elems = 2 * (size_t)p->tpitch * ((size_t)p->band_strips * 32u * (size_t)R + 1);
        }
        if (elems > p->table_elems) {
            if (p->d_table) CK(cudaFreeAsync(p->d_table, p->stream));
            p->d_table = nullptr;
            CK(dev_alloc(p->device, p->stream, &p->d_table, sizeof(int32_t) * elems));
            p->table_elems = elems;
        }
        if (p->streamed && !p->stream2) {
            CK(cudaStreamCreateWithFlags(&p->stream2, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&p->band_ev[0], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&p->band_ev[1], cudaEventDisableTiming));
        }
    }
    int per_sm = 0;
    {
        const int rc1 = occupancy(p->device, (const void*)p->kernel, p->threads, p->smem, &per_sm);
        if (rc1) return rc1;
    }
    if (per_sm < 1) return fail(NW_ERR_CUDA, "strip kernel does not fit on an SM (warps=%d)", p->warps);
    // 32-bit kernels: ~8 resident warps per SM (2 per scheduler) to cover shuffle latency; the packed kernel is
    // issue-bound with one warp per scheduler
    int target_warps_per_sm = env_int("NW_CUDA_WARPS_PER_SM", p->packed ? 4 : 8);
    int ctas_per_sm = std::max(1, std::min(per_sm, target_warps_per_sm / p->warps));
    int cap = d.sm_count * ctas_per_sm;
    int want = (p->nstrips + p->warps - 1) / p->warps;
    int ctas = p->ctas_req;
    if (ctas == 0) ctas = std::min(cap, want);
    ctas = std::max(1, std::min(ctas, d.sm_count * per_sm));    // never more than can be co-resident
    p->ctas = ctas;
    return NW_OK;
}

static int plan_create_internal(nw_plan** out, int device, int32_t n1, int32_t n2, int mode, int part, int nparts,
                                const nw_tuning* tuning, bool want_streamed, const Scoring& sc, bool same_device_pipeline = false,
                                int stair_half = 0);
static int score_build(nw_plan* q, bool stair);
static bool score_can_stair(const nw_plan* q, const bool seen[256]);

extern "C" int nw_plan_create(nw_plan** out, int device, int32_t n1, int32_t n2, int mode, int part, int nparts,
                              const nw_tuning* tuning)
{
    return plan_create_internal(out, device, n1, n2, mode, part, nparts, tuning, false, Scoring());
}

extern "C" int nw_plan_create_scored(nw_plan** out, int device, int32_t n1, int32_t n2, int mode, int part, int nparts,
                                     const nw_tuning* tuning, const nw_scoring* scoring)
{
    if (out) *out = nullptr;
    if (n1 < 0 || n2 < 0) return fail(NW_ERR_ARG, "negative sequence length (n1=%d, n2=%d)", n1, n2);
    Scoring sc;
    const int rc = parse_scoring(scoring, n1, n2, &sc);
    if (rc) return rc;
    return plan_create_internal(out, device, n1, n2, mode, part, nparts, tuning, false, sc);
}

static int plan_create_internal(nw_plan** out, int device, int32_t n1, int32_t n2, int mode, int part, int nparts,
                                const nw_tuning* tuning, bool want_streamed, const Scoring& sc, bool same_device_pipeline,
                                int stair_half)
{
    if (!out) return fail(NW_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (n1 < 0 || n2 < 0) return fail(NW_ERR_ARG, "negative sequence length (n1=%d, n2=%d)", n1, n2);
    if (sc.local) {
        if (nparts != 1 || part != 0) return fail(NW_ERR_UNSUPPORTED, "local alignment is a single-device mode");
        if (mode == NW_MODE_SCORE) mode = NW_MODE_BOUNDARY;      // the best cell can be anywhere: no meeting in the middle
    }
    if (mode == NW_MODE_SCORE) {
        // nparts = 2: the two halves on devices `device` and `device + 1` (their strips never talk to each other; the
        // combine kernel reads the second half's boundary rows and columns through peer access)
        if ((nparts != 1 && nparts != 2) || part != 0) return fail(NW_ERR_UNSUPPORTED, "score mode runs on one device, or on two (part 0 of 2)");
        int rc0 = ensure_device(device);
        if (rc0) return rc0;
        const int dev2 = (nparts == 2) ? device + 1 : device;
        if (dev2 != device) {
            rc0 = ensure_device(dev2);
            if (rc0) return rc0;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, device, dev2));
            if (!can) return fail(NW_ERR_UNSUPPORTED, "device %d cannot access device %d", device, dev2);
            CK(cudaSetDevice(device));
            cudaError_t e = cudaDeviceEnablePeerAccess(dev2, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            cudaGetLastError();
            cudaMemAccessDesc desc;
            memset(&desc, 0, sizeof desc);
            desc.location.type = cudaMemLocationTypeDevice;
            desc.location.id = device;
            desc.flags = cudaMemAccessFlagsProtReadWrite;
            CK(cudaMemPoolSetAccess(g_dev[dev2].pool, &desc, 1));       // the second half's buffers come from this pool
        }
        nw_plan* q = new (std::nothrow) nw_plan;
        if (!q) return fail(NW_ERR_CUDA, "out of host memory");
        // a column costs a full step of the critical path, a row only 1/256 of a strip's start-up lag: columns = shorter one
        q->swapped = n1 > n2 && !env_int("NW_CUDA_NO_SWAP", 0);
        if (q->swapped) std::swap(n1, n2);
        q->device = device; q->n1 = n1; q->n2 = n2; q->mode = mode;
        q->dev2 = dev2;
        q->sc_match = sc.match; q->sc_mis = sc.mismatch; q->sc_gap = sc.gap;
        q->R_req = tuning ? tuning->rows_per_lane : 0;
        q->warps_req = tuning ? tuning->warps_per_cta : 0;
        q->ctas_req = tuning ? tuning->ctas : 0;
        q->split = n2 / 2;
        // provisional (four-letter alphabet assumed, like every plan before its first upload): the staircase
        {
            bool four[256] = {false};
            four[1] = four[2] = four[3] = four[4] = true;
            rc0 = score_build(q, score_can_stair(q, four));
        }
        if (rc0 == NW_OK) {
            cudaSetDevice(dev2);
            if (cudaEventCreateWithFlags(&q->join_ev, cudaEventDisableTiming) != cudaSuccess) rc0 = fail(NW_ERR_CUDA, "cudaEventCreate failed");
            cudaSetDevice(device);
        }
        if (rc0 == NW_OK && cudaMalloc(&q->d_score, 64) != cudaSuccess) rc0 = fail(NW_ERR_CUDA, "device allocation failed");
        if (rc0 != NW_OK) {
            char keep[512];
            memcpy(keep, g_err, sizeof keep);
            nw_plan_destroy(q);
            memcpy(g_err, keep, sizeof keep);
            return rc0;
        }
        *out = q;
        return NW_OK;
    }
    if (mode != NW_MODE_BOUNDARY && mode != NW_MODE_FULL) return fail(NW_ERR_ARG, "unknown mode %d", mode);
    if (nparts < 1 || part < 0 || part >= nparts) return fail(NW_ERR_ARG, "bad part %d of %d", part, nparts);
    if (nparts > 1 && ((long long)n1 + 1) / nparts < 2)
        return fail(NW_ERR_ARG, "n1=%d is too short for %d column strips", n1, nparts);
    int rc = ensure_device(device);
    if (rc) return rc;
    nw_plan* p = new (std::nothrow) nw_plan;
    if (!p) return fail(NW_ERR_CUDA, "out of host memory");
    p->device = device;
    p->n1 = n1;
    p->n2 = n2;
    p->mode = mode;
    p->part = part;
    p->nparts = nparts;
    p->sc_match = sc.match; p->sc_mis = sc.mismatch; p->sc_gap = sc.gap; p->local = sc.local != 0;
    p->want_streamed = want_streamed;
    p->mailbox_pooled = same_device_pipeline;
    p->stair_half = stair_half;
    partition(n1, nparts, part, &p->jstart, &p->ncols);
    rc = plan_alloc(p, tuning);
    if (rc == NW_OK) rc = plan_pick_kernel(p);        // provisional (assumes the four-letter alphabet) so that
    if (rc != NW_OK) {                                // nw_plan_strip_info answers before the first upload
        nw_plan_destroy(p);
        return rc;
    }
    *out = p;
    return NW_OK;
}

// (Re)build the two sub-plans of a score-mode plan: halves of the table above / below row n2/2 (horizontal cut), or two
// plans over all rows whose strips share the columns along a staircase.
static int score_build(nw_plan* q, bool stair)
{
    for (int k = 0; k < 2; ++k) {
        if (q->sub[k]) nw_plan_destroy(q->sub[k]);
        q->sub[k] = nullptr;
    }
    nw_tuning tune;
    memset(&tune, 0, sizeof tune);
    tune.rows_per_lane = q->R_req; tune.warps_per_cta = q->warps_req; tune.ctas = q->ctas_req;
    const Scoring sc = q->scoring_of();
    int rc;
    if (stair) {
        rc = plan_create_internal(&q->sub[0], q->device, q->n1, q->n2, NW_MODE_BOUNDARY, 0, 1, &tune, false, sc, false, 1);
        if (rc == NW_OK) {
            tune.rows_per_lane = q->sub[0]->R_req = q->sub[0]->R;       // both halves: the same strip boundaries
            rc = plan_create_internal(&q->sub[1], q->dev2, q->n1, q->n2, NW_MODE_BOUNDARY, 0, 1, &tune, false, sc, false, 2);
        }
    } else {
        rc = plan_create_internal(&q->sub[0], q->device, q->n1, q->split, NW_MODE_BOUNDARY, 0, 1, &tune, false, sc);
        if (rc == NW_OK) rc = plan_create_internal(&q->sub[1], q->dev2, q->n1, q->n2 - q->split, NW_MODE_BOUNDARY, 0, 1, &tune, false, sc);
    }
    cudaSetDevice(q->device);
    if (rc != NW_OK) return rc;
    q->stair = stair;
    q->uploaded = false;
    q->R = q->sub[1]->R; q->warps = q->sub[1]->warps; q->nstrips = q->sub[0]->nstrips + q->sub[1]->nstrips;
    q->ctas = q->sub[0]->ctas + q->sub[1]->ctas;
    return NW_OK;
}

// Both halves of a staircase span all rows, i.e. twice as many strips are alive as with the horizontal cut.  On one GPU
// that pays only while every strip still has a scheduler to itself (measured on the 64gb pair: 2 x 498 strips on 592
// schedulers, 4.97 ms against 3.84 ms); on two GPUs each half has its own.
static bool stair_fits(const nw_plan* q)
{
    if (q->dev2 != q->device || env_int("NW_CUDA_FORCE_STAIR", 0)) return true;
    const int sms = g_dev[q->device].sm_count;
    int R = q->R_req ? q->R_req : env_int("NW_CUDA_R", 0);
    if (R == 0) R = choose_rows_per_lane_packed(q->n2, q->n1, sms);
    const long long S = ((long long)q->n2 + 32LL * R - 1) / (32LL * R);
    return 2 * S <= 4LL * sms;
}

// the staircase needs the lag-2 kernel on both halves: four letters, small weights, nothing forced through the environment
static bool score_can_stair(const nw_plan* q, const bool seen[256])
{
    uint8_t code[256];
    return build_code(seen, code) && !env_int("NW_CUDA_NO_STAIR", 0) && !env_int("NW_CUDA_GENERIC", 0) &&
           !env_int("NW_CUDA_NO_PACKED", 0) && env_int("NW_CUDA_LAG2", 1) != 0 && env_int("NW_CUDA_WS", 0) == 0 &&
           q->w_max() <= 16 && q->R_req != 1 && env_int("NW_CUDA_R", 0) != 1 && q->n1 > 0 && q->n2 > 0 && stair_fits(q);
}

// widths of the forward strips (x_0 = n1, then n1 * (S - s) / S) and of the reversed ones (the rest), their tail words
static int score_set_widths(nw_plan* q)
{
    nw_plan *a = q->sub[0], *b = q->sub[1];
    const int S = a->nstrips;
    if (!a->lag2 || !b->lag2 || a->ws || b->ws || b->nstrips != S || a->R != b->R || b->pad_top != 0)
        return fail(NW_ERR_STATE, "staircase score mode: the halves disagree (lag2 %d/%d, strips %d/%d, R %d/%d)", (int)a->lag2,
                    (int)b->lag2, S, b->nstrips, a->R, b->R);
    std::vector<int> x((size_t)S), y((size_t)S);
    for (int s = 0; s < S; ++s) x[s] = (s == 0) ? q->n1 : (int)((long long)q->n1 * (S - s) / S);
    for (int t = 0; t < S; ++t) y[t] = q->n1 - x[S - 1 - t];
    nw_plan* h[2] = {a, b};
    const std::vector<int>* w[2] = {&x, &y};
    for (int k = 0; k < 2; ++k) {
        nw_plan* p = h[k];
        CK(cudaSetDevice(p->device));
        if ((size_t)S > p->widths_n) {
            if (p->d_widths) CK(cudaFreeAsync(p->d_widths, p->stream));
            if (p->d_tails) CK(cudaFreeAsync(p->d_tails, p->stream));
            p->d_widths = nullptr; p->d_tails = nullptr;
            CK(dev_alloc(p->device, p->stream, &p->d_widths, sizeof(int) * (size_t)S));
            CK(dev_alloc(p->device, p->stream, &p->d_tails, sizeof(uint32_t) * 64 * (size_t)S));
            p->widths_n = (size_t)S;
        }
        CK(cudaMemcpyAsync(p->d_widths, w[k]->data(), sizeof(int) * (size_t)S, cudaMemcpyHostToDevice, p->stream));
        CK(cudaStreamSynchronize(p->stream));       // (the vector is a local)
        nw::EncodeParams e;
        memcpy(e.code, p->code, 256);
        nw::nw_encode_tails_kernel<<<32, 256, 0, p->stream>>>(p->d_s1, p->d_widths, p->d_tails, S, p->ncols, e);
        CK(cudaGetLastError());
    }
    CK(cudaSetDevice(q->device));
    return NW_OK;
}

static int plan_encode(nw_plan* p, const bool seen[256])
{
    nw::EncodeParams e;
    p->generic = !build_code(seen, e.code);
    // the four-letter paths carry weights as PRMT bytes; the local kernel compares raw bytes
    if (env_int("NW_CUDA_GENERIC", 0) || p->local || p->w_max() > 127) p->generic = true;
    int rc = plan_pick_kernel(p);
    if (rc) return rc;
    e.s1 = p->d_s1;
    e.s2 = p->d_s2;
    e.wq_base = p->d_wq;
    e.rsel = p->d_rsel;
    e.ncols = p->ncols;
    e.n2 = p->n2;
    e.nrows_padded = p->nstrips * 32 * p->R;
    e.pad_top = p->pad_top;
    e.generic = p->generic ? 1 : 0;
    e.packed_regs = p->packed ? p->R / 2 : 0;
    e.lag2 = p->lag2 ? 1 : 0;
    e.w_match = p->w_match();
    e.w_mis = p->w_mis();
    memcpy(p->code, e.code, 256);
    nw::nw_encode_kernel<<<64, 256, 0, p->stream>>>(e);
    CK(cudaGetLastError());
    p->uploaded = true;
    return NW_OK;
}

extern "C" int nw_plan_upload(nw_plan* p, const int8_t* s1, const int8_t* s2)
{
    if (!p) return fail(NW_ERR_ARG, "plan is NULL");
    if (p->mode == NW_MODE_SCORE && p->swapped) std::swap(s1, s2);
    if ((p->n1 > 0 && !s1) || (p->n2 > 0 && !s2)) return fail(NW_ERR_ARG, "sequence pointer is NULL");
    if (p->mode == NW_MODE_SCORE) {
        // first half forwards; second half backwards = forwards on both sequences reversed (reversal on the device: the host
        // copy loop used to cost ~0.4 ms per call).  Horizontal cut: s2[0, split) / s2[split, n2).  Staircase: all of s2 each.
        bool seen[256] = {false};
        const uint8_t* u1 = (const uint8_t*)s1;
        const uint8_t* u2 = (const uint8_t*)s2;
        for (int i = 0; i < p->n1; ++i) seen[u1[i]] = true;
        for (int i = 0; i < p->n2; ++i) seen[u2[i]] = true;
        CK(cudaSetDevice(p->device));
        int rc = NW_OK;
        if (score_can_stair(p, seen) != p->stair) {       // (another alphabet than assumed, or a knob changed: rebuild)
            rc = score_build(p, !p->stair);
            if (rc) return rc;
        }
        nw_plan *a = p->sub[0], *b = p->sub[1];
        const int na = p->stair ? p->n2 : p->split;       // rows of the first half
        const int nb = p->stair ? p->n2 : p->n2 - p->split, ob = p->stair ? 0 : p->split;
        if (p->n1 > 0) CK(cudaMemcpyAsync(a->d_s1, u1, (size_t)p->n1, cudaMemcpyHostToDevice, a->stream));
        if (na > 0) CK(cudaMemcpyAsync(a->d_s2, u2, (size_t)na, cudaMemcpyHostToDevice, a->stream));
        rc = plan_encode(a, seen);
        if (rc) return rc;
        CK(cudaSetDevice(b->device));
        if (!b->d_rev) CK(dev_alloc(b->device, b->stream, &b->d_rev, (size_t)p->n1 + (size_t)nb + 1));
        if (p->n1 > 0) {
            CK(cudaMemcpyAsync(b->d_rev, u1, (size_t)p->n1, cudaMemcpyHostToDevice, b->stream));
            nw::nw_reverse_kernel<<<64, 256, 0, b->stream>>>(b->d_rev, b->d_s1, p->n1);
        }
        if (nb > 0) {
            CK(cudaMemcpyAsync(b->d_rev + p->n1, u2 + ob, (size_t)nb, cudaMemcpyHostToDevice, b->stream));
            nw::nw_reverse_kernel<<<64, 256, 0, b->stream>>>(b->d_rev + p->n1, b->d_s2, nb);
        }
        CK(cudaGetLastError());
        rc = plan_encode(b, seen);
        if (rc == NW_OK && p->stair) rc = score_set_widths(p);
        cudaSetDevice(p->device);
        p->uploaded = rc == NW_OK;
        return rc;
    }
    CK(cudaSetDevice(p->device));
    // alphabet of BOTH full sequences, so that every part of a pipeline makes the same choice of path
    bool seen[256] = {false};
    const uint8_t* a = (const uint8_t*)s1;
    const uint8_t* b = (const uint8_t*)s2;
    for (int i = 0; i < p->n1; ++i) seen[a[i]] = true;
    for (int i = 0; i < p->n2; ++i) seen[b[i]] = true;
    if (p->ncols > 0) CK(cudaMemcpyAsync(p->d_s1, a + p->jstart, (size_t)p->ncols, cudaMemcpyHostToDevice, p->stream));
    if (p->n2 > 0) CK(cudaMemcpyAsync(p->d_s2, b, (size_t)p->n2, cudaMemcpyHostToDevice, p->stream));
    return plan_encode(p, seen);
}

extern "C" int nw_plan_upload_device(nw_plan* p, const int8_t* d_s1, const int8_t* d_s2)
{
    if (!p) return fail(NW_ERR_ARG, "plan is NULL");
    if (p->mode == NW_MODE_SCORE) return fail(NW_ERR_UNSUPPORTED, "score mode takes host sequences (nw_plan_upload)");
    if ((p->n1 > 0 && !d_s1) || (p->n2 > 0 && !d_s2)) return fail(NW_ERR_ARG, "sequence pointer is NULL");
    CK(cudaSetDevice(p->device));
    CK(cudaMemsetAsync(p->d_bitmap, 0, 8 * sizeof(uint32_t), p->stream));
    if (p->n1 > 0) nw::nw_presence_kernel<<<32, 256, 0, p->stream>>>((const uint8_t*)d_s1, p->n1, p->d_bitmap);
    if (p->n2 > 0) nw::nw_presence_kernel<<<32, 256, 0, p->stream>>>((const uint8_t*)d_s2, p->n2, p->d_bitmap);
    CK(cudaGetLastError());
    uint32_t bm[8];
    CK(cudaMemcpyAsync(bm, p->d_bitmap, sizeof bm, cudaMemcpyDeviceToHost, p->stream));
    if (p->ncols > 0)
        CK(cudaMemcpyAsync(p->d_s1, (const uint8_t*)d_s1 + p->jstart, (size_t)p->ncols, cudaMemcpyDeviceToDevice, p->stream));
    if (p->n2 > 0) CK(cudaMemcpyAsync(p->d_s2, d_s2, (size_t)p->n2, cudaMemcpyDeviceToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    bool seen[256];
    bitmap_to_seen(bm, seen);
    return plan_encode(p, seen);
}

// ---- pipeline wiring ----------------------------------------------------------------------------------------------------
extern "C" int nw_plan_connect(nw_plan* left, nw_plan* right)
{
    if (!left || !right) return fail(NW_ERR_ARG, "plan is NULL");
    if (left->mode == NW_MODE_SCORE || right->mode == NW_MODE_SCORE) return fail(NW_ERR_UNSUPPORTED, "score mode is a single-device mode");
    if (left->nparts != right->nparts || right->part != left->part + 1 || left->n1 != right->n1 || left->n2 != right->n2)
        return fail(NW_ERR_ARG, "plans are not adjacent parts of the same pipeline");
    if (left->device != right->device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, left->device, right->device));
        if (!can) return fail(NW_ERR_UNSUPPORTED, "device %d cannot access device %d", left->device, right->device);
        CK(cudaSetDevice(left->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(right->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
        CK(cudaSetDevice(right->device));
        e = cudaDeviceEnablePeerAccess(left->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
        left->rcol_peer = true;
        right->halo_peer = true;
    }
    left->rcol_target = right->d_mailbox;
    return NW_OK;
}

// handle64 layout: [0..63] cudaIpcMemHandle_t of the consumer's mailbox
extern "C" int nw_plan_export_mailbox(nw_plan* p, void* handle64)
{
    if (!p || !handle64) return fail(NW_ERR_ARG, "NULL argument");
    if (p->mailbox_pooled) return fail(NW_ERR_STATE, "this plan's mailbox is pool memory (same-device pipeline)");
    if (p->mode == NW_MODE_SCORE) return fail(NW_ERR_UNSUPPORTED, "score mode is a single-device mode");
    if (p->part == 0) return fail(NW_ERR_STATE, "part 0 has no halo mailbox");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaSetDevice(p->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, p->d_mailbox));
    memcpy(handle64, &h, 64);
    p->halo_peer = true;
    return NW_OK;
}

extern "C" int nw_plan_import_mailbox(nw_plan* p, const void* handle64, int consumer_device)
{
    if (!p || !handle64) return fail(NW_ERR_ARG, "NULL argument");
    if (p->mode == NW_MODE_SCORE) return fail(NW_ERR_UNSUPPORTED, "score mode is a single-device mode");
    if (p->part == p->nparts - 1) return fail(NW_ERR_STATE, "the last part has no right neighbour");
    (void)consumer_device;
    CK(cudaSetDevice(p->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->ipc_mailbox = ptr;
    p->rcol_target = (int2*)ptr;
    p->rcol_peer = true;
    return NW_OK;
}

// ---- running ----------------------------------------------------------------------------------------------------------
static int plan_enqueue(nw_plan* p)
{
    if (!p->uploaded) return fail(NW_ERR_STATE, "nw_plan_upload has not been called");
    CK(cudaSetDevice(p->device));
    p->epoch += 1;
    const int par = p->epoch & 1;
    const int2* halo = (p->part > 0) ? p->d_mailbox + (long long)par * p->mpitch : nullptr;
    int2* rcol = p->rcol_target + (long long)par * p->mpitch;
    const bool have_cells = p->ncols > 0 && p->n2 > 0;
    if (p->mode == NW_MODE_FULL && !p->streamed) {
        const int bgap = p->local ? 0 : p->sc_gap;      // local alignment: zero first row and column
        nw::nw_table_row0_kernel<<<64, 256, 0, p->stream>>>(p->d_table, p->ncols, p->jstart, bgap);
        CK(cudaGetLastError());
        if (!have_cells && p->n2 > 0) {   // no interior column: the table is just the boundary column
            if (halo) return fail(NW_ERR_UNSUPPORTED, "full-table part without interior columns");
            nw::nw_table_col0_kernel<<<64, 256, 0, p->stream>>>(p->d_table, p->tpitch, p->n2, bgap);
            CK(cudaGetLastError());
        }
    }
    if (have_cells) {
        nw::StripParams sp;
        sp.wq = p->d_wq + nw::WQ_PAD;
        sp.rsel = p->d_rsel;
        sp.brow = p->brow();
        sp.pitch = p->pitch;
        sp.halo = halo;
        sp.rcol = rcol;
        sp.table = p->d_table;
        sp.tpitch = p->tpitch;
        sp.dump = p->d_dump;
        sp.ncols = p->ncols;
        sp.n2 = p->n2;
        sp.nstrips = p->nstrips;
        sp.pad_top = p->pad_top;
        sp.jstart = p->jstart;
        sp.epoch = p->epoch;
        sp.halo_sys = p->halo_peer ? 1 : 0;
        sp.rcol_sys = p->rcol_peer ? 1 : 0;
        sp.ack_in = (p->rcol_target != p->d_rcol_local) ? (const int*)(p->rcol_target + 2 * p->mpitch) : nullptr;
        sp.times = p->d_times;
        sp.w_match = p->local ? p->sc_match : p->w_match();      // (the local kernel works in H form with the plain scores)
        sp.w_mis = p->local ? p->sc_mis : p->w_mis();
        sp.local_best = p->d_local_best;
        sp.widths = p->stair_half ? p->d_widths : nullptr;
        sp.tails = p->stair_half ? p->d_tails : nullptr;
        sp.gap = p->sc_gap;
        sp.margin = 2 * p->w_max() + 10;
        {   // bounded waits (NW_CUDA_SPIN_TIMEOUT_MS, default 20 s; 0 = wait for ever)
            const int ms = env_int("NW_CUDA_SPIN_TIMEOUT_MS", 20000);
            sp.abort_flag = ms > 0 ? g_dev[p->device].d_abort_dev : nullptr;
            sp.abort_host = g_dev[p->device].d_abort;
            sp.spin_ns = (unsigned long long)std::max(1, ms) * 1000000ULL;
        }
        sp.snap = p->kernel2 ? p->d_snap : nullptr;
        sp.tile_blocks = p->tile_blocks;
        sp.ntiles = p->ntiles;
        sp.s_begin = 0;
        sp.s_count = p->nstrips;
        void* args[] = {&sp};
        CK(cudaLaunchCooperativeKernel((const void*)p->kernel, dim3(p->ctas), dim3(p->threads), args, p->smem, p->stream));
        p->last_sp = sp;
        if (p->kernel2 && !p->streamed && !env_int("NW_CUDA_DBG_SKIP_PASS2", 0)) {           // pass 2: every tile of every strip at once, table stores as 128-byte row segments
            sp.ack_in = nullptr;
            p->kernel2<<<p->ctas2, p->warps2 * 32, p->smem2, p->stream>>>(sp);
            CK(cudaGetLastError());
        }
    }
    if (p->local) {
        nw::nw_local_finish_kernel<<<1, 32, 0, p->stream>>>(p->d_local_best, have_cells ? p->nstrips : 0, p->d_score);
        CK(cudaGetLastError());
    } else {
        const int2* brow_last = (p->n2 > 0 && have_cells) ? p->brow() + (long long)(p->nstrips - 1) * p->pitch : nullptr;
        // n2 > 0 but no interior column: the last row is the single boundary cell; handled through rcol/halo == nullptr
        const int2* rc = have_cells ? rcol : nullptr;
        nw::nw_finish_kernel<<<64, 256, 0, p->stream>>>(brow_last, rc, halo, p->ncols, p->n2, p->jstart, p->d_last_row,
                                                         p->d_last_col, p->d_score,
                                                         p->d_mailbox ? (int*)(p->d_mailbox + 2 * p->mpitch) : nullptr, p->epoch,
                                                         p->sc_gap);
        CK(cudaGetLastError());
    }
    return NW_OK;
}

// NW_MODE_SCORE: both halves on their own streams, then the combine kernel on the top half's stream
static int score_enqueue(nw_plan* p)
{
    if (!p->uploaded) return fail(NW_ERR_STATE, "nw_plan_upload has not been called");
    nw_plan *a = p->sub[0], *b = p->sub[1];
    int rc = plan_enqueue(b);
    if (rc == NW_OK) rc = plan_enqueue(a);
    if (rc) return rc;
    CK(cudaSetDevice(b->device));
    CK(cudaEventRecord(p->join_ev, b->stream));
    CK(cudaSetDevice(a->device));
    CK(cudaStreamWaitEvent(a->stream, p->join_ev, 0));
    if (p->stair) {
        nw::StairParams sp;
        sp.brow_f = a->brow(); sp.pitch_f = a->pitch;
        sp.brow_b = b->brow(); sp.pitch_b = b->pitch;
        sp.rcol_f = a->rcol_target + (long long)(a->epoch & 1) * a->mpitch;
        sp.rcol_b = b->rcol_target + (long long)(b->epoch & 1) * b->mpitch;
        sp.widths = a->d_widths;
        sp.nstrips = a->nstrips; sp.strip_rows = 32 * a->R; sp.pad_top = a->pad_top;
        sp.n1 = p->n1; sp.n2 = p->n2; sp.gap = p->sc_gap;
        sp.score = p->d_score;
        nw::nw_set_int_kernel<<<1, 1, 0, a->stream>>>(p->d_score, p->sc_gap * (p->n1 + p->n2));      // the all-gap path
        nw::nw_stair_combine_kernel<<<std::max(1, std::min(148, a->nstrips)), 256, 0, a->stream>>>(sp);
    } else {
        nw::nw_set_int_kernel<<<1, 1, 0, a->stream>>>(p->d_score, INT_MIN);
        nw::nw_bidir_combine_kernel<<<std::max(1, std::min(148, (p->n1 + 256) / 256)), 256, 0, a->stream>>>(a->d_last_row, b->d_last_row,
                                                                                                      p->n1, p->d_score);
    }
    CK(cudaGetLastError());
    p->epoch += 1;
    return NW_OK;
}

extern "C" int nw_plan_run(nw_plan* p)
{
    if (!p) return fail(NW_ERR_ARG, "plan is NULL");
    CK(cudaSetDevice(p->device));
    if (p->mode == NW_MODE_SCORE) {
        // the bottom half's stream must not start before the timing event of the top half's stream
        CK(cudaEventRecord(p->sub[0]->ev0, p->sub[0]->stream));
        CK(cudaSetDevice(p->sub[1]->device));
        CK(cudaStreamWaitEvent(p->sub[1]->stream, p->sub[0]->ev0, 0));
        CK(cudaSetDevice(p->device));
        int rc = score_enqueue(p);
        if (rc) return rc;
        CK(cudaEventRecord(p->sub[0]->ev1, p->sub[0]->stream));
        return NW_OK;
    }
    CK(cudaEventRecord(p->ev0, p->stream));
    int rc = plan_enqueue(p);
    if (rc) return rc;
    CK(cudaEventRecord(p->ev1, p->stream));
    return NW_OK;
}

// after a stream synchronisation: did a kernel of this device give up waiting?
static int check_abort(int device)
{
    DeviceState& d = g_dev[device];
    if (d.h_abort && *d.h_abort) {
        *d.h_abort = 0;
        cudaMemset(d.d_abort_dev, 0, 64);
        return fail(NW_ERR_CUDA, "fill aborted on device %d: a strip waited longer than NW_CUDA_SPIN_TIMEOUT_MS for its "
                                 "predecessor (a neighbouring part failed or was never run); results are invalid", device);
    }
    return NW_OK;
}

extern "C" int nw_plan_sync(nw_plan* p)
{
    if (!p) return fail(NW_ERR_ARG, "plan is NULL");
    CK(cudaSetDevice(p->device));
    if (p->mode == NW_MODE_SCORE) {
        CK(cudaStreamSynchronize(p->sub[1]->stream));
        CK(cudaStreamSynchronize(p->sub[0]->stream));
        const int rc2 = (p->dev2 != p->device) ? check_abort(p->dev2) : NW_OK;
        const int rc1 = check_abort(p->device);
        return rc1 ? rc1 : rc2;
    }
    CK(cudaStreamSynchronize(p->stream));
    return check_abort(p->device);
}

extern "C" int nw_plan_time(nw_plan* p, int iters, float* ms_per_fill)
{
    if (!p || !ms_per_fill || iters < 1) return fail(NW_ERR_ARG, "bad argument");
    if (p->nparts > 1) return fail(NW_ERR_STATE, "nw_plan_time is for single-part plans; time pipelines with nw_plan_run");
    CK(cudaSetDevice(p->device));
    if (p->mode == NW_MODE_SCORE) {
        nw_plan *a = p->sub[0], *b = p->sub[1];
        CK(cudaStreamSynchronize(a->stream));
        CK(cudaStreamSynchronize(b->stream));
        CK(cudaEventRecord(a->ev2, a->stream));
        CK(cudaSetDevice(b->device));
        CK(cudaStreamWaitEvent(b->stream, a->ev2, 0));
        CK(cudaSetDevice(a->device));
        for (int i = 0; i < iters; ++i) {
            int rc = score_enqueue(p);        // every combine joins the two streams on a's
            if (rc) return rc;
        }
        CK(cudaEventRecord(a->ev3, a->stream));
        CK(cudaEventSynchronize(a->ev3));
        CK(cudaStreamSynchronize(b->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a->ev2, a->ev3));
        *ms_per_fill = ms / iters;
        return NW_OK;
    }
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaEventRecord(p->ev0, p->stream));
    for (int i = 0; i < iters; ++i) {
        int rc = plan_enqueue(p);
        if (rc) return rc;
    }
    CK(cudaEventRecord(p->ev1, p->stream));
    CK(cudaEventSynchronize(p->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    *ms_per_fill = ms / iters;
    return NW_OK;
}

extern "C" int nw_plan_timer_start(nw_plan* p)
{
    if (!p) return fail(NW_ERR_ARG, "plan is NULL");
    CK(cudaSetDevice(p->device));
    if (p->mode == NW_MODE_SCORE) {
        CK(cudaEventRecord(p->sub[0]->ev2, p->sub[0]->stream));
        CK(cudaSetDevice(p->sub[1]->device));
        CK(cudaStreamWaitEvent(p->sub[1]->stream, p->sub[0]->ev2, 0));
        CK(cudaSetDevice(p->device));
        return NW_OK;
    }
    CK(cudaEventRecord(p->ev2, p->stream));
    return NW_OK;
}

extern "C" int nw_plan_timer_stop(nw_plan* p, float* ms)
{
    if (!p || !ms) return fail(NW_ERR_ARG, "bad argument");
    CK(cudaSetDevice(p->device));
    if (p->mode == NW_MODE_SCORE) {          // the last combine kernel already joined both streams on sub[0]'s
        nw_plan* a = p->sub[0];
        CK(cudaEventRecord(a->ev3, a->stream));
        CK(cudaEventSynchronize(a->ev3));
        CK(cudaEventElapsedTime(ms, a->ev2, a->ev3));
        return NW_OK;
    }
    CK(cudaEventRecord(p->ev3, p->stream));
    CK(cudaEventSynchronize(p->ev3));
    CK(cudaEventElapsedTime(ms, p->ev2, p->ev3));
    return NW_OK;
}

extern "C" int nw_plan_last_ms(nw_plan* p, float* ms)
{
    if (!p || !ms) return fail(NW_ERR_ARG, "bad argument");
    CK(cudaSetDevice(p->device));
    if (p->mode == NW_MODE_SCORE) p = p->sub[0];
    CK(cudaEventSynchronize(p->ev1));
    CK(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return NW_OK;
}

extern "C" int nw_plan_launches_per_run(nw_plan* p, int* n)
{
    if (!p || !n) return fail(NW_ERR_ARG, "bad argument");
    if (p->mode == NW_MODE_SCORE) {
        int a = 0, b = 0;
        nw_plan_launches_per_run(p->sub[0], &a);
        nw_plan_launches_per_run(p->sub[1], &b);
        *n = a + b + 2;
        return NW_OK;
    }
    const bool have_cells = p->ncols > 0 && p->n2 > 0;
    *n = (have_cells ? 1 : 0) + 1 + (p->mode == NW_MODE_FULL ? 1 + ((!have_cells && p->n2 > 0) ? 1 : 0) : 0) +
         ((have_cells && p->kernel2) ? 1 : 0);
    return NW_OK;
}

// ---- results ------------------------------------------------------------------------------------------------------------
extern "C" int nw_plan_score(nw_plan* p, int32_t* score)
{
    if (!p || !score) return fail(NW_ERR_ARG, "bad argument");
    if (p->epoch == 0) return fail(NW_ERR_STATE, "no fill has been run");
    CK(cudaSetDevice(p->device));
    cudaStream_t st = (p->mode == NW_MODE_SCORE) ? p->sub[0]->stream : p->stream;
    CK(cudaMemcpyAsync(score, p->d_score, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return check_abort(p->device);
}

// the reported cell: (n2, n1) for global alignment, the best cell for local alignment
extern "C" int nw_plan_best(nw_plan* p, int32_t* score, int32_t* end_i, int32_t* end_j)
{
    if (!p) return fail(NW_ERR_ARG, "bad argument");
    if (p->epoch == 0) return fail(NW_ERR_STATE, "no fill has been run");
    CK(cudaSetDevice(p->device));
    cudaStream_t st = (p->mode == NW_MODE_SCORE) ? p->sub[0]->stream : p->stream;
    int32_t v[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(v, p->d_score, sizeof(int32_t) * (p->local ? 3 : 1), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (!p->local) {
        const bool sw = p->mode == NW_MODE_SCORE && p->swapped;      // a score plan may hold the sequences swapped
        v[1] = sw ? p->n1 : p->n2;
        v[2] = sw ? p->n2 : p->n1;
    }
    if (score) *score = v[0];
    if (end_i) *end_i = v[1];
    if (end_j) *end_j = v[2];
    return check_abort(p->device);
}

extern "C" int nw_plan_last_row(nw_plan* p, int32_t* last_row)
{
    if (!p || !last_row) return fail(NW_ERR_ARG, "bad argument");
    if (p->local) return fail(NW_ERR_UNSUPPORTED, "a local-alignment plan reports its best cell (nw_plan_best), not boundaries");
    if (p->mode == NW_MODE_SCORE) return fail(NW_ERR_STATE, "a score-mode plan only has a score");
    if (p->epoch == 0) return fail(NW_ERR_STATE, "no fill has been run");
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(last_row, p->d_last_row, sizeof(int32_t) * ((size_t)p->ncols + 1), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return check_abort(p->device);
}

extern "C" int nw_plan_last_col(nw_plan* p, int32_t* last_col)
{
    if (!p || !last_col) return fail(NW_ERR_ARG, "bad argument");
    if (p->local) return fail(NW_ERR_UNSUPPORTED, "a local-alignment plan reports its best cell (nw_plan_best), not boundaries");
    if (p->mode == NW_MODE_SCORE) return fail(NW_ERR_STATE, "a score-mode plan only has a score");
    if (p->epoch == 0) return fail(NW_ERR_STATE, "no fill has been run");
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(last_col, p->d_last_col, sizeof(int32_t) * ((size_t)p->n2 + 1), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return check_abort(p->device);
}

// ---- table delivery to a PAGEABLE host buffer -------------------------------------------------------------------------
// The reference's driver owns the table as `new int[size]` (src/common/driver.cpp:22): a plain cudaMemcpy into it runs
// at a few GB/s.  Instead the table goes device -> pinned staging (DMA at PCIe rate) -> destination (several host
// threads), double-buffered so that the DMA of chunk k+1 overlaps the host copy of chunk k.
namespace {
// Persistent host threads for the staging -> table copies.  (Spawning and joining 16 std::threads per 32 MB chunk cost more
// than the copy itself: 2 GB through 64 chunks took 64 ms, of which the DMA needs 40.)
class CopyPool {
public:
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // runs job(0) .. job(n-1), job(0) on the calling thread; returns when all are done
    void run(int n, const std::function<void(int)>& job)
    {
        if (n <= 1) {
            job(0);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            while ((int)workers_.size() < n - 1) {
                const int id = (int)workers_.size() + 1;
                workers_.emplace_back([this, id] { loop(id); });
            }
            job_ = &job;
            njobs_ = n;
            remaining_ = n - 1;
            ++generation_;
        }
        cv_work_.notify_all();
        job(0);
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return remaining_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int id)
    {
        unsigned seen = 0;
        for (;;) {
            const std::function<void(int)>* job = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                if (id < njobs_) job = job_;
            }
            if (job) {
                (*job)(id);
                std::lock_guard<std::mutex> lk(mu_);
                if (--remaining_ == 0) cv_done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)>* job_ = nullptr;
    int njobs_ = 0, remaining_ = 0;
    unsigned generation_ = 0;
    bool stop_ = false;
};

constexpr int STAGE_SLOTS = 4;     // pinned staging buffers per device: up to three DMAs in flight behind the host copy
struct Staging {
    void* buf[STAGE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[STAGE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    size_t bytes = 0;
    int device = -1;
    CopyPool* pool = nullptr;      // (leaked at exit on purpose: joining threads in a static destructor is fragile)
};
Staging g_stages[64];              // one pair of pinned buffers per device: parts of a pipeline deliver concurrently
std::mutex g_stage_mus[64];

// Copy out of a pinned staging buffer into the caller's table with non-temporal stores: the destination is written once
// and not read back here, so streaming stores save the read-for-ownership traffic of a plain memcpy (SSE2: baseline x86-64).
void stream_copy(char* dst, const char* src, size_t n)
{
    if (env_int("NW_CUDA_NO_NT", 0)) { memcpy(dst, src, n); return; }
    size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    if (head > n) head = n;
    if (head) memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    const size_t blocks = n / 64;
    for (size_t b = 0; b < blocks; ++b) {
        const __m128i a0 = _mm_loadu_si128((const __m128i*)(src + 64 * b));
        const __m128i a1 = _mm_loadu_si128((const __m128i*)(src + 64 * b + 16));
        const __m128i a2 = _mm_loadu_si128((const __m128i*)(src + 64 * b + 32));
        const __m128i a3 = _mm_loadu_si128((const __m128i*)(src + 64 * b + 48));
        _mm_stream_si128((__m128i*)(dst + 64 * b), a0);
        _mm_stream_si128((__m128i*)(dst + 64 * b + 16), a1);
        _mm_stream_si128((__m128i*)(dst + 64 * b + 32), a2);
        _mm_stream_si128((__m128i*)(dst + 64 * b + 48), a3);
    }
    _mm_sfence();
    if (n - 64 * blocks) memcpy(dst + 64 * blocks, src + 64 * blocks, n - 64 * blocks);
}

void parallel_memcpy(CopyPool* pool, char* dst, const char* src, size_t n, int nthreads)
{
    if (n < (4u << 20) || nthreads <= 1 || pool == nullptr) {
        stream_copy(dst, src, n);
        return;
    }
    const size_t per = ((n / nthreads) + 4095) & ~(size_t)4095;
    const int njobs = (int)std::min<size_t>((size_t)nthreads, (n + per - 1) / per);
    pool->run(njobs, [=](int t) {
        const size_t off = per * (size_t)t;
        if (off < n) stream_copy(dst + off, src + off, std::min(per, n - off));
    });
}

bool host_pointer_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
}  // namespace

static int ensure_staging(int device)       // caller holds g_stage_mus[device]
{
    Staging& g_stage = g_stages[device];
    const size_t chunk = (size_t)std::max(1, env_int("NW_CUDA_STAGE_MB", 32)) << 20;
    if (g_stage.bytes == chunk && g_stage.device == device) return NW_OK;
    CK(cudaSetDevice(device));
    for (int i = 0; i < STAGE_SLOTS; ++i) {
        if (g_stage.buf[i]) cudaFreeHost(g_stage.buf[i]);
        if (g_stage.ev[i]) cudaEventDestroy(g_stage.ev[i]);
        g_stage.buf[i] = nullptr;
        g_stage.ev[i] = nullptr;
    }
    g_stage.bytes = 0;
    for (int i = 0; i < STAGE_SLOTS; ++i) {
        CK(cudaHostAlloc(&g_stage.buf[i], chunk, cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&g_stage.ev[i], cudaEventDisableTiming));
    }
    if (!g_stage.pool) g_stage.pool = new CopyPool;
    g_stage.bytes = chunk;
    g_stage.device = device;
    return NW_OK;
}

// rows x width bytes from device (pitch dpitch) to host (pitch hpitch), host pageable
static int staged_d2h_2d_locked(nw_plan* p, char* dst, size_t hpitch, const char* src, size_t dpitch, size_t width,
                                size_t rows, cudaStream_t cs);

static int staged_d2h_2d(nw_plan* p, char* dst, size_t hpitch, const char* src, size_t dpitch, size_t width, size_t rows,
                         cudaStream_t cs = nullptr)
{
    if (cs == nullptr) cs = p->stream;
    std::lock_guard<std::mutex> lk(g_stage_mus[p->device]);
    int rc0 = ensure_staging(p->device);
    if (rc0) return rc0;
    const size_t chunk = g_stages[p->device].bytes;
    const bool flat = (width == hpitch && width == dpitch);
    if (flat || width <= chunk) return staged_d2h_2d_locked(p, dst, hpitch, src, dpitch, width, rows, cs);
    // a row is wider than a staging buffer (n1 > ~8.4 M columns at the default 32 MB, less with a small
    // NW_CUDA_STAGE_MB): deliver the table in column segments that fit
    for (size_t off = 0; off < width; off += chunk) {
        const int rc = staged_d2h_2d_locked(p, dst + off, hpitch, src + off, dpitch, std::min(chunk, width - off), rows, cs);
        if (rc) return rc;
    }
    return NW_OK;
}

static int staged_d2h_2d_locked(nw_plan* p, char* dst, size_t hpitch, const char* src, size_t dpitch, size_t width,
                                size_t rows, cudaStream_t cs)       // caller holds g_stage_mus[device]; width <= chunk or flat
{
    Staging& g_stage = g_stages[p->device];
    const size_t chunk = g_stage.bytes;
    const int nthreads = std::max(1, std::min(env_int("NW_CUDA_COPY_THREADS", 16), (int)std::thread::hardware_concurrency()) /
                                      std::max(1, p->nparts > 1 ? std::min(p->nparts, 4) : 1));
    const bool flat = (width == hpitch && width == dpitch);
    // unit of work: a run of whole rows (or a byte range when the table is contiguous on both sides)
    const size_t total = flat ? width * rows : rows;
    const size_t step = flat ? chunk : std::max<size_t>(1, chunk / width);
    // a ring of STAGE_SLOTS pinned buffers: DMAs are issued as far ahead as there are free slots, the host retires them in
    // order with the persistent copy threads
    size_t done_issue = 0, done_copy = 0;
    size_t pend_off[STAGE_SLOTS], pend_n[STAGE_SLOTS];
    int head = 0, tail = 0, inflight = 0;           // next slot to issue into / oldest slot in flight
    while (done_copy < total) {
        while (done_issue < total && inflight < STAGE_SLOTS) {
            const size_t n = std::min(step, total - done_issue);
            if (flat) CK(cudaMemcpyAsync(g_stage.buf[head], src + done_issue, n, cudaMemcpyDeviceToHost, cs));
            else CK(cudaMemcpy2DAsync(g_stage.buf[head], width, src + done_issue * dpitch, dpitch, width, n,
                                      cudaMemcpyDeviceToHost, cs));
            CK(cudaEventRecord(g_stage.ev[head], cs));
            pend_off[head] = done_issue;
            pend_n[head] = n;
            done_issue += n;
            head = (head + 1) % STAGE_SLOTS;
            ++inflight;
        }
        const int o = tail;
        CK(cudaEventSynchronize(g_stage.ev[o]));
        if (flat) parallel_memcpy(g_stage.pool, dst + pend_off[o], (const char*)g_stage.buf[o], pend_n[o], nthreads);
        else {
            const char* sb = (const char*)g_stage.buf[o];
            const size_t r0 = pend_off[o], nr = pend_n[o];
            if (nr * width < (4u << 20) || nthreads <= 1) {
                for (size_t r = 0; r < nr; ++r) stream_copy(dst + (r0 + r) * hpitch, sb + r * width, width);
            } else {
                const size_t per = (nr + nthreads - 1) / nthreads;
                const int njobs = (int)((nr + per - 1) / per);
                g_stage.pool->run(njobs, [=](int t) {
                    const size_t a = per * (size_t)t, b = std::min(nr, a + per);
                    for (size_t r = a; r < b; ++r) stream_copy(dst + (r0 + r) * hpitch, sb + r * width, width);
                });
            }
        }
        done_copy += pend_n[o];
        tail = (tail + 1) % STAGE_SLOTS;
        --inflight;
    }
    return NW_OK;
}

// Streamed delivery: pass 2 of band k+1 (kernel, stream) overlaps the device->host copy of band k (stream2 + host threads).
static int plan_stream_table_to_host(nw_plan* p, int32_t* table)
{
    CK(cudaSetDevice(p->device));
    const int SH = 32 * p->R;
    const size_t band_rows = (size_t)p->band_strips * SH + 1;                 // + table row 0 in band 0
    const size_t band_elems = (size_t)p->tpitch * band_rows;
    const size_t host_pitch = sizeof(int32_t) * ((size_t)p->n1 + 1), dpitch = sizeof(int32_t) * (size_t)p->tpitch;
    const int nbands = (p->nstrips + p->band_strips - 1) / p->band_strips;
    auto first_row = [&](int k) { return k == 0 ? 0LL : (long long)k * p->band_strips * SH - p->pad_top + 1; };
    auto last_row = [&](int k) { return std::min<long long>(p->n2, (long long)(k + 1) * p->band_strips * SH - p->pad_top); };
    auto launch = [&](int k) -> int {
        nw::StripParams sp = p->last_sp;
        int32_t* buf = p->d_table + (size_t)(k & 1) * band_elems;
        sp.table = buf - first_row(k) * p->tpitch;                            // the kernel indexes by table row
        sp.s_begin = k * p->band_strips;
        sp.s_count = std::min(p->band_strips, p->nstrips - sp.s_begin);
        sp.ack_in = nullptr;
        if (k == 0) nw::nw_table_row0_kernel<<<64, 256, 0, p->stream>>>(buf, p->ncols, p->jstart, p->sc_gap);
        const long long ntasks = (long long)sp.s_count * p->ntiles;
        const int ctas = (int)std::max<long long>(1, std::min<long long>((ntasks + p->warps2 - 1) / p->warps2, p->ctas2));
        p->kernel2<<<ctas, p->warps2 * 32, p->smem2, p->stream>>>(sp);
        CK(cudaGetLastError());
        CK(cudaEventRecord(p->band_ev[k & 1], p->stream));
        return NW_OK;
    };
    int rc = launch(0);
    for (int k = 0; k < nbands && rc == NW_OK; ++k) {
        // band k+1 goes into the other half of the ring, whose previous content (band k-1) has already been delivered
        if (k + 1 < nbands) rc = launch(k + 1);
        if (rc != NW_OK) break;
        CK(cudaStreamWaitEvent(p->stream2, p->band_ev[k & 1], 0));
        const long long r0 = first_row(k), r1 = last_row(k);
        rc = staged_d2h_2d(p, (char*)table + (size_t)r0 * host_pitch, host_pitch,
                           (const char*)(p->d_table + (size_t)(k & 1) * band_elems), dpitch, host_pitch,
                           (size_t)(r1 - r0 + 1), p->stream2);
    }
    return rc;
}

extern "C" int nw_plan_table_to_host(nw_plan* p, int32_t* table)
{
    if (!p || !table) return fail(NW_ERR_ARG, "bad argument");
    if (p->mode != NW_MODE_FULL) return fail(NW_ERR_STATE, "plan is not in full-table mode");
    if (p->epoch == 0) return fail(NW_ERR_STATE, "no fill has been run");
    CK(cudaSetDevice(p->device));
    if (p->streamed) return plan_stream_table_to_host(p, table);
    const size_t host_pitch = sizeof(int32_t) * ((size_t)p->n1 + 1);
    // parts > 0 own their halo column too, but it is the left neighbour's last column: copy interior columns only
    const int skip = (p->part > 0) ? 1 : 0;
    const size_t width = sizeof(int32_t) * ((size_t)p->ncols + 1 - skip);
    if (width == 0) return NW_OK;
    const size_t rows = (size_t)p->n2 + 1;
    const bool pinned = host_pointer_is_pinned(table);
    if (!pinned && width * rows >= (8u << 20) && !env_int("NW_CUDA_NO_STAGING", 0)) {
        CK(cudaStreamSynchronize(p->stream));
        if (p->nparts == 1)
            return staged_d2h_2d(p, (char*)table, host_pitch, (const char*)p->d_table,
                                 sizeof(int32_t) * (size_t)p->tpitch, host_pitch, rows);
        return staged_d2h_2d(p, (char*)(table + p->jstart + skip), host_pitch, (const char*)(p->d_table + skip),
                             sizeof(int32_t) * (size_t)p->tpitch, width, rows);
    }
    if (p->nparts == 1) {
        CK(cudaMemcpy2DAsync(table, host_pitch, p->d_table, sizeof(int32_t) * (size_t)p->tpitch, host_pitch, rows,
                             cudaMemcpyDeviceToHost, p->stream));
    } else {
        CK(cudaMemcpy2DAsync(table + p->jstart + skip, host_pitch, p->d_table + skip, sizeof(int32_t) * (size_t)p->tpitch,
                             width, rows, cudaMemcpyDeviceToHost, p->stream));
    }
    CK(cudaStreamSynchronize(p->stream));
    return NW_OK;
}

extern "C" int nw_plan_table_device(nw_plan* p, int32_t** d_table, int64_t* pitch)
{
    if (!p || !d_table || !pitch) return fail(NW_ERR_ARG, "bad argument");
    if (p->mode != NW_MODE_FULL) return fail(NW_ERR_STATE, "plan is not in full-table mode");
    if (p->streamed) return fail(NW_ERR_STATE, "the table of a streamed one-shot call is not resident");
    *d_table = p->d_table;
    *pitch = p->tpitch;
    return NW_OK;
}

#include <chrono>
namespace {
struct Trace {      // NW_CUDA_TRACE=1: wall-clock phases of a one-shot call on stderr (stdout belongs to the driver)
    bool on = env_int("NW_CUDA_TRACE", 0) != 0;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char* what)
    {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[nw_cuda] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
}  // namespace

// the kernels emit the path backwards into d_out[0..cap) / d_out[cap..2cap); *d_n = its length
static int fetch_reversed_path(cudaStream_t st, const uint8_t* d_out, size_t cap, const int* d_n, int8_t* a1, int8_t* a2, int32_t* len)
{
    int n = 0;
    std::vector<uint8_t> r1(cap), r2(cap);
    CK(cudaMemcpyAsync(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(r1.data(), d_out, cap, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(r2.data(), d_out + cap, cap, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n < 0 || (size_t)n > cap) return fail(NW_ERR_CUDA, "traceback: bad path length %d", n);
    for (int k = 0; k < n; ++k) {
        a1[k] = (int8_t)r1[(size_t)n - 1 - k];
        a2[k] = (int8_t)r2[(size_t)n - 1 - k];
    }
    *len = n;
    return NW_OK;
}

// Traceback from the checkpoint rows and columns of boundary-mode parts (nw_tile_trace_kernel): no table anywhere.
extern "C" int nw_plans_traceback(nw_plan* const* parts, int nparts, int8_t* a1, int8_t* a2, int32_t* len)
{
    if (!parts || nparts < 1 || !a1 || !a2 || !len) return fail(NW_ERR_ARG, "bad argument");
    nw_plan* p0 = parts[0];
    for (int g = 0; g < nparts; ++g) {
        nw_plan* q = parts[g];
        if (!q) return fail(NW_ERR_ARG, "part %d is NULL", g);
        if (q->mode != NW_MODE_BOUNDARY || q->local) return fail(NW_ERR_STATE, "checkpoint traceback needs boundary-mode global-alignment plans");
        if (q->nparts != nparts || q->part != g || q->n1 != p0->n1 || q->n2 != p0->n2 || q->device != p0->device)
            return fail(NW_ERR_ARG, "plans are not the %d parts of one pipeline on one device", nparts);
        if (q->epoch == 0 || q->epoch != p0->epoch) return fail(NW_ERR_STATE, "the parts have not all run the same fill");
        if (q->R != p0->R || q->nstrips != p0->nstrips || q->pad_top != p0->pad_top) return fail(NW_ERR_STATE, "the parts differ in strip height");
        if (q->sc_match != p0->sc_match || q->sc_mis != p0->sc_mis || q->sc_gap != p0->sc_gap) return fail(NW_ERR_ARG, "the parts differ in scoring");
        if (g > 0 && parts[g - 1]->rcol_target != q->d_mailbox) return fail(NW_ERR_STATE, "part %d is not connected to part %d", g - 1, g);
    }
    const int SH = 32 * p0->R;
    if (SH > nw::TT_MAX_ROWS) return fail(NW_ERR_UNSUPPORTED, "strip height %d exceeds %d", SH, nw::TT_MAX_ROWS);
    CK(cudaSetDevice(p0->device));
    for (int g = 0; g < nparts; ++g) {
        CK(cudaStreamSynchronize(parts[g]->stream));
        int rc = check_abort(parts[g]->device);
        if (rc) return rc;
    }
    nw_plan* last = parts[nparts - 1];
    cudaStream_t st = last->stream;
    const int dev = p0->device;
    const size_t cap = (size_t)p0->n1 + (size_t)p0->n2 + 1;
    int maxcols = 0;
    for (int g = 0; g < nparts; ++g) maxcols = std::max(maxcols, parts[g]->ncols);
    const long long spitch = ((long long)std::max(maxcols, SH) + 1 + 7) & ~7LL;
    // phase-major cells: (blocks of the widest part + rows) phases x TT_B x SH
    const size_t scratch_ints = 2 * (size_t)spitch + ((size_t)(maxcols + nw::TT_B - 1) / nw::TT_B + (size_t)SH) * nw::TT_B * (size_t)SH;
    uint8_t *d_out = nullptr, *d_s1 = nullptr;
    int32_t* d_scratch = nullptr;
    int* d_state = nullptr;
    nw::TracePart* d_parts = nullptr;
    std::vector<nw::TracePart> hp((size_t)nparts);
    for (int g = 0; g < nparts; ++g) {
        nw_plan* q = parts[g];
        hp[g].brow = q->brow();
        hp[g].pitch = q->pitch;
        hp[g].halo = (g > 0) ? q->d_mailbox + (long long)(q->epoch & 1) * q->mpitch : nullptr;
        hp[g].jstart = q->jstart;
        hp[g].ncols = q->ncols;
    }
    int rc = NW_OK;
    auto body = [&]() -> int {
        CK(dev_alloc(dev, st, &d_out, 2 * cap));
        CK(dev_alloc(dev, st, &d_s1, (size_t)p0->n1 + 32));      // (+ padding: the tile kernel reads whole aligned words)
        CK(dev_alloc(dev, st, &d_scratch, sizeof(int32_t) * scratch_ints));
        CK(dev_alloc(dev, st, &d_state, 8 * sizeof(int)));
        CK(dev_alloc(dev, st, &d_parts, sizeof(nw::TracePart) * (size_t)nparts));
        for (int g = 0; g < nparts; ++g)       // the whole s1 back together from the parts' slices
            if (parts[g]->ncols > 0)
                CK(cudaMemcpyAsync(d_s1 + parts[g]->jstart, parts[g]->d_s1, (size_t)parts[g]->ncols, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(d_parts, hp.data(), sizeof(nw::TracePart) * (size_t)nparts, cudaMemcpyHostToDevice, st));
        const int state0[8] = {p0->n2, p0->n1, 0, 0, 0, 0, 0, 0};
        CK(cudaMemcpyAsync(d_state, state0, sizeof state0, cudaMemcpyHostToDevice, st));
        nw::TraceParams tp;
        tp.parts = d_parts;
        tp.nparts = nparts;
        tp.s1 = d_s1;
        tp.s2 = last->d_s2;
        tp.n1 = p0->n1; tp.n2 = p0->n2; tp.nstrips = p0->nstrips; tp.strip_rows = SH; tp.pad_top = p0->pad_top;
        tp.sc_match = p0->sc_match; tp.sc_mis = p0->sc_mis; tp.sc_gap = p0->sc_gap;
        tp.scratch = d_scratch;
        tp.spitch = spitch;
        tp.rpitch = SH;
        tp.state = d_state;
        tp.out1 = d_out;
        tp.out2 = d_out + cap;
        // a monotone path visits at most one tile per strip plus one per part, then one launch for the forced tail
        const int launches = std::max(p0->nstrips, 0) + nparts + 2;
        for (int k = 0; k < launches; ++k) nw::nw_tile_trace_kernel<<<1, SH, 0, st>>>(tp);
        CK(cudaGetLastError());
        int state[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        CK(cudaMemcpyAsync(state, d_state, sizeof state, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (env_int("NW_CUDA_TRACE", 0))
            fprintf(stderr, "[nw_cuda] tile traceback: %d tiles of %d launches, fill %.1f Mcycles in %d block steps, walk %.1f Mcycles\n",
                    state[4], launches, state[5] * 1024e-6, state[7], state[6] * 1024e-6);
        if (state[3] != 0 || state[0] != 0 || state[1] != 0)
            return fail(NW_ERR_CUDA, "checkpoint traceback did not finish (stopped at %d, %d, flag %d)", state[0], state[1], state[3]);
        return fetch_reversed_path(st, d_out, cap, d_state + 2, a1, a2, len);
    };
    rc = body();
    void* bufs[] = {d_out, d_s1, d_scratch, d_state, d_parts};
    for (void* b : bufs)
        if (b) cudaFreeAsync(b, st);
    cudaStreamSynchronize(st);
    return rc;
}

extern "C" int nw_plan_traceback(nw_plan* p, int8_t* a1, int8_t* a2, int32_t* len)
{
    if (!p || !a1 || !a2 || !len) return fail(NW_ERR_ARG, "bad argument");
    if (p->mode == NW_MODE_BOUNDARY && !p->local) {          // from the checkpoint rows: no table
        if (p->nparts != 1) return fail(NW_ERR_ARG, "a pipeline is traced back through nw_plans_traceback (all its parts)");
        nw_plan* one[1] = {p};
        return nw_plans_traceback(one, 1, a1, a2, len);
    }
    if (p->mode != NW_MODE_FULL) return fail(NW_ERR_STATE, "traceback needs a boundary-mode or full-table plan");
    if (p->nparts != 1 || p->streamed) return fail(NW_ERR_UNSUPPORTED, "table traceback needs the whole table on one device");
    if (p->epoch == 0) return fail(NW_ERR_STATE, "no fill has been run");
    CK(cudaSetDevice(p->device));
    const size_t cap = (size_t)p->n1 + (size_t)p->n2 + 1;
    uint8_t* d_out = nullptr;
    int* d_len = nullptr;
    CK(dev_alloc(p->device, p->stream, &d_out, 2 * cap));
    CK(dev_alloc(p->device, p->stream, &d_len, sizeof(int)));
    if (p->local)       // from the best cell back to the first zero
        nw::nw_traceback_local_kernel<<<1, 256, 0, p->stream>>>(p->d_table, p->tpitch, p->d_s1, p->d_s2, p->d_score, d_out,
                                                                d_out + cap, d_len, p->sc_match, p->sc_mis, p->sc_gap);
    else
        nw::nw_traceback_kernel<<<1, 256, 0, p->stream>>>(p->d_table, p->tpitch, p->d_s1, p->d_s2, p->n1, p->n2, d_out,
                                                          d_out + cap, d_len, p->sc_match, p->sc_mis, p->sc_gap);
    int rc = (cudaGetLastError() == cudaSuccess) ? NW_OK : fail(NW_ERR_CUDA, "traceback launch failed");
    if (rc == NW_OK) rc = fetch_reversed_path(p->stream, d_out, cap, d_len, a1, a2, len);
    cudaFreeAsync(d_out, p->stream);
    cudaFreeAsync(d_len, p->stream);
    cudaStreamSynchronize(p->stream);
    return rc;
}

// Global alignment of one pair with NO table: P column parts on one device (a checkpoint column every ~NW_CUDA_ALIGN_TILE
// table columns, a checkpoint row per strip), filled one part after the other, then the tile-by-tile traceback.
extern "C" int nw_cuda_align(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, const nw_scoring* scoring,
                             int8_t* a1, int8_t* a2, int32_t* len, int32_t* score)
{
    if (!a1 || !a2 || !len) return fail(NW_ERR_ARG, "bad argument");
    if (n1 < 0 || n2 < 0) return fail(NW_ERR_ARG, "negative sequence length");
    Scoring sc;
    int rc = parse_scoring(scoring, n1, n2, &sc);
    if (rc) return rc;
    if (sc.local) return fail(NW_ERR_UNSUPPORTED, "nw_cuda_align is a global alignment");
    const int tile = std::max(64, env_int("NW_CUDA_ALIGN_TILE", 4096));
    int P = (int)std::min<long long>(64, std::max<long long>(1, ((long long)n1 + tile - 1) / tile));
    while (P > 1 && ((long long)n1 + 1) / P < 2) --P;
    std::vector<nw_plan*> parts((size_t)P, nullptr);
    nw_tuning tune;
    memset(&tune, 0, sizeof tune);
    Trace tr;
    for (int g = 0; g < P && rc == NW_OK; ++g) {
        rc = plan_create_internal(&parts[g], 0, n1, n2, NW_MODE_BOUNDARY, g, P, g == 0 ? nullptr : &tune, false, sc, true);
        if (rc == NW_OK && g == 0) tune.rows_per_lane = parts[0]->R_req = parts[0]->R;     // every part: the same strip height
        if (g < 3) tr.mark(g == 0 ? "align:   part 0 created" : "align:   next part created");
    }
    for (int g = 0; g + 1 < P && rc == NW_OK; ++g) rc = nw_plan_connect(parts[g], parts[g + 1]);
    tr.mark("align: create parts");
    for (int g = 0; g < P && rc == NW_OK; ++g) rc = nw_plan_upload(parts[g], s1, s2);
    tr.mark("align: upload");
    // One device, P cooperative kernels: part g polls part g-1's right column, exactly like a part on another GPU would, so
    // consecutive parts may be resident together -- part g then trails part g-1 by one part width instead of by a whole
    // start-up ramp of the strip chain.  A window of W parts is in flight (part g waits for part g-W to finish); parts
    // finish in order, so the oldest unfinished part is always resident and nothing can deadlock.
    const int W = std::max(1, env_int("NW_CUDA_ALIGN_WINDOW", 8));
    for (int g = 0; g < P && rc == NW_OK; ++g) {
        if (g >= W && cudaStreamWaitEvent(parts[g]->stream, parts[g - W]->ev1, 0) != cudaSuccess) rc = fail(NW_ERR_CUDA, "cudaStreamWaitEvent failed");
        if (rc == NW_OK) rc = nw_plan_run(parts[g]);
    }
    if (rc == NW_OK && score) rc = nw_plan_score(parts[(size_t)P - 1], score);
    tr.mark("align: fill (all parts)");
    if (rc == NW_OK) rc = nw_plans_traceback(parts.data(), P, a1, a2, len);
    tr.mark("align: tile traceback");
    char keep[512];
    memcpy(keep, g_err, sizeof keep);
    for (nw_plan* q : parts) nw_plan_destroy(q);
    memcpy(g_err, keep, sizeof keep);
    tr.mark("align: destroy");
    return rc;
}

extern "C" int nw_plan_strip_info(nw_plan* p, int* nstrips, int* strip_rows, int* rows_per_lane, int* warps, int* ctas)
{
    if (!p) return fail(NW_ERR_ARG, "plan is NULL");
    if (nstrips) *nstrips = p->nstrips;
    if (strip_rows) *strip_rows = 32 * p->R;
    if (rows_per_lane) *rows_per_lane = p->R;
    if (warps) *warps = p->warps;
    if (ctas) *ctas = p->ctas;
    return NW_OK;
}

extern "C" int nw_plan_strip_times(nw_plan* p, int64_t* start_ns, int64_t* end_ns, int64_t* sm_cycles)
{
    if (!p || !start_ns || !end_ns) return fail(NW_ERR_ARG, "bad argument");
    if (p->mode == NW_MODE_SCORE) return fail(NW_ERR_STATE, "ask the halves of a score-mode plan");
    if (p->epoch == 0) return fail(NW_ERR_STATE, "no fill has been run");
    CK(cudaSetDevice(p->device));
    std::vector<unsigned long long> t(4 * (size_t)std::max(p->nstrips, 1));
    CK(cudaMemcpyAsync(t.data(), p->d_times, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    for (int s = 0; s < p->nstrips; ++s) {
        start_ns[s] = (int64_t)t[4 * (size_t)s];
        end_ns[s] = (int64_t)t[4 * (size_t)s + 1];
        if (sm_cycles) sm_cycles[s] = (int64_t)(t[4 * (size_t)s + 3] - t[4 * (size_t)s + 2]);
    }
    return NW_OK;
}

extern "C" int nw_plan_strip_row(nw_plan* p, int strip, int32_t* row)
{
    if (!p || !row) return fail(NW_ERR_ARG, "bad argument");
    if (p->mode == NW_MODE_SCORE) return fail(NW_ERR_STATE, "a score-mode plan only has a score");
    if (strip < 0 || strip >= p->nstrips) return fail(NW_ERR_ARG, "strip %d out of range (%d strips)", strip, p->nstrips);
    if (p->epoch == 0) return fail(NW_ERR_STATE, "no fill has been run");
    if (p->ncols == 0) return fail(NW_ERR_STATE, "part has no interior columns");
    CK(cudaSetDevice(p->device));
    const int row_i = p->n2 - (p->nstrips - 1 - strip) * 32 * p->R;
    nw::nw_strip_row_kernel<<<64, 256, 0, p->stream>>>(p->brow() + (long long)strip * p->pitch, p->ncols, row_i, p->jstart,
                                                        p->d_tmp_row, p->local ? 0 : p->sc_gap);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(row, p->d_tmp_row, sizeof(int32_t) * ((size_t)p->ncols + 1), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return NW_OK;
}

// =====================================================================================================================
// one-shot entry points (host buffers)
// =====================================================================================================================
// every kernel the automatic choices can launch: touched once in nw_cuda_init so that (lazy) module loading never lands
// inside a timed one-shot call
static std::vector<const void*> all_kernels()
{
    std::vector<const void*> v;
    for (int regs : {1, 2, 4, 8}) {
        v.push_back((const void*)strip16l2_kernel(regs));
        v.push_back((const void*)strip16l2_kernel(regs, true));
        v.push_back((const void*)strip16ws_kernel(regs));
        v.push_back((const void*)strip16_kernel(regs));
        v.push_back((const void*)full16_kernel(regs));
    }
    for (int R : {1, 2, 4, 8})
        for (int g = 0; g < 2; ++g)
            for (int f = 0; f < 2; ++f) v.push_back((const void*)strip_kernel(R, g != 0, f != 0));
    for (int R : {1, 2, 4, 8})
        for (int f = 0; f < 2; ++f) v.push_back((const void*)local_kernel(R, f != 0));
    v.push_back((const void*)nw::nw_local_finish_kernel);
    for (int R : {4, 8, 16, 32})
        for (int g = 0; g < 2; ++g) v.push_back((const void*)batch_kernel(R, g != 0));
    for (int regs : {2, 4, 8, 16}) v.push_back((const void*)batch16_kernel(regs));
    v.push_back((const void*)nw::nw_encode_kernel);
    v.push_back((const void*)nw::nw_finish_kernel);
    v.push_back((const void*)nw::nw_presence_kernel);
    v.push_back((const void*)nw::nw_presence_kernel64);
    v.push_back((const void*)nw::nw_table_row0_kernel);
    v.push_back((const void*)nw::nw_table_col0_kernel);
    v.push_back((const void*)nw::nw_strip_row_kernel);
    v.push_back((const void*)nw::nw_set_int_kernel);
    v.push_back((const void*)nw::nw_bidir_combine_kernel);
    v.push_back((const void*)nw::nw_stair_combine_kernel);
    v.push_back((const void*)nw::nw_encode_tails_kernel);
    v.push_back((const void*)nw::nw_reverse_kernel);
    v.push_back((const void*)nw::nw_traceback_kernel);
    v.push_back((const void*)nw::nw_tile_trace_kernel);
    v.push_back((const void*)nw::nw_traceback_local_kernel);
    return v;
}

extern "C" int nw_cuda_init(int device)
{
    int rc = ensure_device(device);
    if (rc) return rc;
    DeviceState& d = g_dev[device];
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (d.warmed) return NW_OK;
    }
#if NW_L2_DBG
    return NW_OK;      // timing-experiment builds compute garbage: no self-test
#endif
    CK(cudaSetDevice(device));
    // (1) load every kernel
    for (const void* fn : all_kernels()) {
        cudaFuncAttributes a;
        if (fn) CK(cudaFuncGetAttributes(&a, fn));
    }
    // (2) pre-fault the memory pool (NW_CUDA_POOL_MB, default 1536: the boundary rows of the reference's largest pair are
    //     0.5 GB): allocate, free, keep -- the release threshold of the pool is "never"
    {
        const size_t bytes = (size_t)std::max(0, env_int("NW_CUDA_POOL_MB", 1536)) << 20;
        if (bytes) {
            void* tmp = nullptr;
            cudaError_t e = dev_alloc(device, (cudaStream_t)0, &tmp, bytes);
            if (e == cudaSuccess) e = cudaFreeAsync(tmp, (cudaStream_t)0);
            if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)0);
            if (e != cudaSuccess) cudaGetLastError();      // a smaller device: the pool grows on demand instead
        }
    }
    // (3) self-test = warm-up: small fills with the kernels a real call will use (boundary, score and both passes of the
    //     packed full-table mode), so that the first cooperative launches happen here and not inside the reference driver's
    //     timed call (src/common/driver.cpp:26-30)
    std::vector<int8_t> a(700);
    for (size_t i = 0; i < a.size(); ++i) a[i] = (int8_t)(1 + ((i * 7 + i / 5) & 3));
    nw_tuning tune;
    memset(&tune, 0, sizeof tune);
    tune.rows_per_lane = 8;
    for (int mode : {NW_MODE_BOUNDARY, NW_MODE_FULL, NW_MODE_SCORE}) {
        if (rc != NW_OK) break;
        int32_t score = 0;
        nw_plan* p = nullptr;
        rc = nw_plan_create(&p, device, (int32_t)a.size(), (int32_t)a.size(), mode, 0, 1, &tune);
        if (rc == NW_OK) rc = nw_plan_upload(p, a.data(), a.data());
        if (rc == NW_OK) rc = nw_plan_run(p);
        if (rc == NW_OK) rc = nw_plan_score(p, &score);
        char keep[512];
        memcpy(keep, g_err, sizeof keep);
        nw_plan_destroy(p);
        memcpy(g_err, keep, sizeof keep);
        if (rc == NW_OK && score != (int32_t)a.size())
            return fail(NW_ERR_CUDA, "self-test failed: score %d, expected %d", score, (int)a.size());
    }
    if (rc == NW_OK) {      // pinned staging for table delivery, so the first timed call does not allocate it
        std::lock_guard<std::mutex> lk(g_stage_mus[device]);
        rc = ensure_staging(device);
    }
    // (4) column strips over NW_CUDA_GPUS devices of this process: contexts, pools and peer access to the neighbours now,
    //     not inside the timed call
    const int ng = env_int("NW_CUDA_GPUS", 1);
    if (rc == NW_OK && ng > 1 && device < ng) {
        for (int peer : {device - 1, device + 1}) {
            if (peer < 0 || peer >= ng) continue;
            rc = ensure_device(peer);
            if (rc != NW_OK) break;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, device, peer));
            if (!can) continue;      // nw_plan_connect reports it
            CK(cudaSetDevice(device));
            cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            cudaGetLastError();
        }
        // score mode with one half per GPU: device 0's combine kernel reads device 1's boundary rows, which live in
        // device 1's pool -- mapping a pre-faulted pool into a peer costs ~0.4 s, so it happens here, not in the timed call
        if (rc == NW_OK && device == 1) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, 0, 1));
            if (can) {
                cudaMemAccessDesc desc;
                memset(&desc, 0, sizeof desc);
                desc.location.type = cudaMemLocationTypeDevice;
                desc.location.id = 0;
                desc.flags = cudaMemAccessFlagsProtReadWrite;
                CK(cudaMemPoolSetAccess(d.pool, &desc, 1));
            }
        }
    }
    if (rc == NW_OK) {
        std::lock_guard<std::mutex> lk(g_mu);
        d.warmed = true;
    }
    return rc;
}


static int run_pipeline(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int mode, int ngpus,
                        int32_t* table, int32_t* last_row, int32_t* last_col, int32_t* score, const Scoring& sc = Scoring(),
                        int32_t* end_i = nullptr, int32_t* end_j = nullptr)
{
    Trace tr;
    if (ngpus < 1) return fail(NW_ERR_ARG, "ngpus must be >= 1");
    if (ngpus > 1 && ((long long)n1 + 1) / ngpus < 2) ngpus = 1;     // too narrow to split
    int ndev = nw_cuda_device_count();
    if (ndev < 0) return ndev;
    if (ndev < ngpus) return fail(NW_ERR_ARG, "%d GPUs requested, %d visible", ngpus, ndev);
    // One-shot calls keep their plans alive in a one-entry cache: a repeated call with the same shape reuses the device
    // buffers, and -- what matters for the reference driver, which calls exactly once -- the call never pays for
    // cudaFree (synchronising, tens of ms for a 2 GB table).  The cache is dropped when the shape changes.
    static std::mutex cache_mu;
    static std::vector<nw_plan*> cache;
    static int c_n1 = -1, c_n2 = -1, c_mode = -1, c_ngpus = -1;
    static Scoring c_sc;
    std::lock_guard<std::mutex> cache_lock(cache_mu);
    int rc = NW_OK;
    if (sc.local && ngpus != 1) return fail(NW_ERR_UNSUPPORTED, "local alignment is a single-device mode");
    if (!(c_n1 == n1 && c_n2 == n2 && c_mode == mode && c_ngpus == ngpus && c_sc == sc && (int)cache.size() == ngpus)) {
        for (nw_plan* p : cache) nw_plan_destroy(p);
        cache.assign((size_t)ngpus, nullptr);
        nw_tuning tune;
        memset(&tune, 0, sizeof tune);
        for (int g = 0; g < ngpus && rc == NW_OK; ++g) {
            rc = plan_create_internal(&cache[g], g, n1, n2, mode, g, ngpus, g == 0 ? nullptr : &tune,
                                      /*want_streamed=*/ngpus == 1 && mode == NW_MODE_FULL && table != nullptr, sc);
            if (rc == NW_OK && g == 0 && ngpus > 1) tune.rows_per_lane = cache[0]->R_req = cache[0]->R;     // all parts share the strip height
        }
        for (int g = 0; g + 1 < ngpus && rc == NW_OK; ++g) rc = nw_plan_connect(cache[g], cache[g + 1]);
        if (rc != NW_OK) {
            char keep[512];
            memcpy(keep, g_err, sizeof keep);
            for (nw_plan* p : cache) nw_plan_destroy(p);
            memcpy(g_err, keep, sizeof keep);
            cache.clear();
            c_n1 = -1;
            return rc;
        }
        c_n1 = n1; c_n2 = n2; c_mode = mode; c_ngpus = ngpus; c_sc = sc;
    }
    std::vector<nw_plan*>& plans = cache;
    tr.mark("plan_create");
    for (int g = 0; g < ngpus && rc == NW_OK; ++g) rc = nw_plan_upload(plans[g], s1, s2);
    tr.mark("upload (async)");
    for (int g = 0; g < ngpus && rc == NW_OK; ++g) rc = nw_plan_run(plans[g]);
    tr.mark("run (enqueue)");
    for (int g = 0; g < ngpus && rc == NW_OK; ++g) rc = nw_plan_sync(plans[g]);
    tr.mark("sync (kernels done)");
    nw_plan* last = plans[(size_t)ngpus - 1];
    if (rc == NW_OK && table && mode == NW_MODE_FULL) {
        if (ngpus == 1) rc = nw_plan_table_to_host(plans[0], table);
        else {
            // every part delivers its own columns over its own PCIe link, concurrently
            std::vector<int> rcs((size_t)ngpus, NW_OK);
            std::vector<std::string> msgs((size_t)ngpus);
            std::vector<std::thread> th;
            for (int g = 0; g < ngpus; ++g)
                th.emplace_back([&, g] {
                    rcs[g] = nw_plan_table_to_host(plans[g], table);
                    if (rcs[g] != NW_OK) msgs[g] = g_err;          // g_err is thread-local
                });
            for (auto& x : th) x.join();
            for (int g = 0; g < ngpus; ++g)
                if (rcs[g] != NW_OK && rc == NW_OK) rc = fail(rcs[g], "%s", msgs[g].c_str());
        }
    }
    tr.mark("table_to_host");
    if (rc == NW_OK && (score || end_i || end_j)) rc = nw_plan_best(last, score, end_i, end_j);
    if (rc == NW_OK && last_col) rc = nw_plan_last_col(last, last_col);
    if (rc == NW_OK && last_row)
        for (int g = 0; g < ngpus && rc == NW_OK; ++g) {
            // part g's row covers global columns jstart .. jstart+ncols; column jstart of parts > 0 repeats the
            // left neighbour's last column, so overlapping writes agree
            std::vector<int32_t> tmp((size_t)plans[g]->ncols + 1);
            rc = nw_plan_last_row(plans[g], tmp.data());
            if (rc == NW_OK) memcpy(last_row + plans[g]->jstart, tmp.data(), sizeof(int32_t) * tmp.size());
        }
    tr.mark("results");
    if (rc != NW_OK) {
        // A failure part-way leaves the parts with different epochs, and the tagged hand-off needs them in lockstep: a
        // later call that reused these plans would wait for tags nobody writes.  Drop them.
        char keep[512];
        memcpy(keep, g_err, sizeof keep);
        for (nw_plan* p : cache) nw_plan_destroy(p);
        memcpy(g_err, keep, sizeof keep);
        cache.clear();
        c_n1 = c_n2 = c_mode = c_ngpus = -1;
    }
    return rc;
}

// score only, through a cached NW_MODE_SCORE plan (what the reference driver reads in boundary mode: driver.cpp:35)
static int run_score_oneshot(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* score,
                             const Scoring& sc = Scoring(), int gpus = 1)
{
    Trace tr;
    static std::mutex mu;
    static nw_plan* cached = nullptr;
    static int c_n1 = -1, c_n2 = -1, c_gpus = -1;         // the caller's sizes (a score plan may hold them swapped)
    static Scoring c_sc;
    std::lock_guard<std::mutex> lk(mu);
    int rc = NW_OK;
    if (!cached || c_n1 != n1 || c_n2 != n2 || c_gpus != gpus || !(c_sc == sc)) {
        if (cached) nw_plan_destroy(cached);
        cached = nullptr;
        c_n1 = c_n2 = -1;
        rc = plan_create_internal(&cached, 0, n1, n2, NW_MODE_SCORE, 0, gpus, nullptr, false, sc);
        if (rc) return rc;
        c_n1 = n1;
        c_n2 = n2;
        c_sc = sc;
        c_gpus = gpus;
    }
    tr.mark("plan_create");
    rc = nw_plan_upload(cached, s1, s2);
    tr.mark("upload");
    if (rc == NW_OK) rc = nw_plan_run(cached);
    if (rc == NW_OK) rc = nw_plan_score(cached, score);
    tr.mark("run + score");
    if (rc != NW_OK) {
        char keep[512];
        memcpy(keep, g_err, sizeof keep);
        nw_plan_destroy(cached);
        memcpy(g_err, keep, sizeof keep);
        cached = nullptr;
        c_n1 = c_n2 = -1;
    }
    return rc;
}

static int fill_scored(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* table, int mode, int ngpus,
                       const Scoring& sc)
{
    if (mode == NW_MODE_FULL) return run_pipeline(s1, n1, s2, n2, mode, ngpus, table, nullptr, nullptr, nullptr, sc);
    if (mode != NW_MODE_BOUNDARY) return fail(NW_ERR_ARG, "unknown mode %d", mode);
    int32_t score = 0;
    // one GPU: score mode (two half-length chains); more: score mode with one half on each of the first two GPUs (along a
    // staircase their chains are shorter still, and there is nothing a third GPU could shorten: a single fill is bound by
    // its critical path).  NW_CUDA_NO_BIDIR=1: one forward fill, column strips over all GPUs.
    if (ngpus < 1) return fail(NW_ERR_ARG, "ngpus must be >= 1");
    int rc = (!sc.local && !env_int("NW_CUDA_NO_BIDIR", 0))
                 ? run_score_oneshot(s1, n1, s2, n2, &score, sc, std::min(ngpus, 2))
                 : run_pipeline(s1, n1, s2, n2, mode, ngpus, nullptr, nullptr, nullptr, &score, sc);
    if (rc == NW_OK) table[((long long)n1 + 1) * ((long long)n2 + 1) - 1] = score;     // what driver.cpp:35 reads
    return rc;
}

extern "C" int nw_cuda_fill_ex(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* table, int mode,
                               int ngpus)
{
    if (!table) return fail(NW_ERR_ARG, "table is NULL");
    if (n1 < 0 || n2 < 0) return fail(NW_ERR_ARG, "negative sequence length");
    return fill_scored(s1, n1, s2, n2, table, mode, ngpus, Scoring());
}

// mode < 0 / ngpus < 1: taken from NW_CUDA_MODE / NW_CUDA_GPUS like nw_cuda_fill (the reference driver's argv is fixed)
extern "C" int nw_cuda_fill_scored(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* table, int mode,
                                   int ngpus, const nw_scoring* scoring)
{
    if (!table) return fail(NW_ERR_ARG, "table is NULL");
    if (n1 < 0 || n2 < 0) return fail(NW_ERR_ARG, "negative sequence length");
    Scoring sc;
    int rc = parse_scoring(scoring, n1, n2, &sc);
    if (rc) return rc;
    if (mode < 0) {
        mode = NW_MODE_FULL;
        const char* m = getenv("NW_CUDA_MODE");
        if (m && *m) {
            if (!strcmp(m, "boundary")) mode = NW_MODE_BOUNDARY;
            else if (strcmp(m, "full")) return fail(NW_ERR_ARG, "NW_CUDA_MODE must be 'full' or 'boundary' (got '%s')", m);
        }
    }
    if (ngpus < 1) ngpus = env_int("NW_CUDA_GPUS", 1);
    return fill_scored(s1, n1, s2, n2, table, mode, ngpus, sc);
}

extern "C" int nw_cuda_score_scored(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, const nw_scoring* scoring,
                                    int32_t* score, int32_t* end_i, int32_t* end_j)
{
    if (!score) return fail(NW_ERR_ARG, "score is NULL");
    if (n1 < 0 || n2 < 0) return fail(NW_ERR_ARG, "negative sequence length");
    Scoring sc;
    int rc = parse_scoring(scoring, n1, n2, &sc);
    if (rc) return rc;
    if (!sc.local && !env_int("NW_CUDA_NO_BIDIR", 0)) {
        rc = run_score_oneshot(s1, n1, s2, n2, score, sc);
        if (rc == NW_OK && end_i) *end_i = n2;
        if (rc == NW_OK && end_j) *end_j = n1;
        return rc;
    }
    return run_pipeline(s1, n1, s2, n2, NW_MODE_BOUNDARY, 1, nullptr, nullptr, nullptr, score, sc, end_i, end_j);
}

extern "C" int nw_cuda_fill(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* table)
{
    int mode = NW_MODE_FULL;
    const char* m = getenv("NW_CUDA_MODE");
    if (m && *m) {
        if (!strcmp(m, "boundary")) mode = NW_MODE_BOUNDARY;
        else if (!strcmp(m, "full")) mode = NW_MODE_FULL;
        else return fail(NW_ERR_ARG, "NW_CUDA_MODE must be 'full' or 'boundary' (got '%s')", m);
    }
    const int ngpus = env_int("NW_CUDA_GPUS", 1);
    return nw_cuda_fill_ex(s1, n1, s2, n2, table, mode, ngpus);
}

extern "C" int nw_cuda_score(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* score)
{
    if (!score) return fail(NW_ERR_ARG, "score is NULL");
    if (n1 < 0 || n2 < 0) return fail(NW_ERR_ARG, "negative sequence length");
    if (!env_int("NW_CUDA_NO_BIDIR", 0)) return run_score_oneshot(s1, n1, s2, n2, score);
    return run_pipeline(s1, n1, s2, n2, NW_MODE_BOUNDARY, 1, nullptr, nullptr, nullptr, score);
}

extern "C" int nw_cuda_boundaries(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* last_row,
                                  int32_t* last_col, int32_t* score)
{
    if (n1 < 0 || n2 < 0) return fail(NW_ERR_ARG, "negative sequence length");
    return run_pipeline(s1, n1, s2, n2, NW_MODE_BOUNDARY, 1, nullptr, last_row, last_col, score);
}

// =====================================================================================================================
// batch plans
// =====================================================================================================================
struct nw_batch {
    int device = 0;
    long long npairs = 0;
    int len1 = 0, len2 = 0;
    int R = 32, nstrips = 0, pad_top = 0, warps = 8, ctas = 0;
    bool generic = false, uploaded = false, ran = false;
    bool packed = false;      // nw_batch16_kernel: at most four letters and G fits 15 bits
    size_t scratch_words = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint8_t *d_S1 = nullptr, *d_S2 = nullptr;
    const uint8_t *S1 = nullptr, *S2 = nullptr;    // what the kernel reads (own copies or caller's device arrays)
    int32_t *d_scores = nullptr, *d_scratch = nullptr;
    uint32_t* d_bitmap = nullptr;
    long long scratch_pitch = 0;
    uint8_t code[256];
    BatchKernel kernel = nullptr;
    size_t smem = 0;
    cudaStream_t copy_stream = nullptr;            // nw_batch_run_host: H2D of chunk c+1 while chunk c computes
    std::vector<cudaEvent_t> chunk_ev;
    int sc_match = 1, sc_mis = 0, sc_gap = -1;     // nw_scoring (default: the reference's macros)
    int w_match() const { return std::max(sc_match - 2 * sc_gap, 0); }
    int w_mis() const { return std::max(sc_mis - 2 * sc_gap, 0); }
};

static int batch_scan(nw_batch* b);

extern "C" int nw_batch_destroy(nw_batch* b)
{
    if (!b) return NW_OK;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    void* bufs[] = {b->d_S1, b->d_S2, b->d_scores, b->d_scratch, b->d_bitmap};
    for (void* x : bufs)
        if (x) cudaFree(x);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    for (cudaEvent_t e : b->chunk_ev) cudaEventDestroy(e);
    if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
    return NW_OK;
}

// geometry + kernel choice; needs the alphabet, so it is called again after every upload
static int batch_pick_kernel(nw_batch* b)
{
    const DeviceState& d = g_dev[b->device];
    CK(cudaSetDevice(b->device));
    const int wmax = std::max(b->w_match(), b->w_mis());
    const long long gmax = (long long)wmax * std::min(b->len1, b->len2) + 64;
    b->packed = !b->generic && gmax < 32000 && wmax <= 127 && !env_int("NW_CUDA_NO_PACKED", 0);
    if (!b->packed && !b->generic && wmax > 127) b->generic = true;        // the PRMT paths carry weights as bytes
    int R = env_int("NW_CUDA_BATCH_R", 0);       // table rows per lane
    if (R == 0) {
        // the smallest strip that covers the pair in one pass (profiles/r01_batch_sweep.log: with the edge blocks gone,
        // one pass of 32 rows per lane beats two passes of 16)
        R = 32;
        while (R > 4 && b->len2 <= 32 * (R / 2)) R /= 2;
    }
    if (R != 4 && R != 8 && R != 16 && R != 32) return fail(NW_ERR_ARG, "batch rows_per_lane must be 4, 8, 16 or 32");
    b->R = R;
    b->nstrips = (int)(((long long)b->len2 + 32LL * R - 1) / (32LL * R));
    b->pad_top = b->nstrips * 32 * R - b->len2;
    b->warps = b->packed ? 4 : 8;
    b->kernel = b->packed ? batch16_kernel(R / 2) : batch_kernel(R, b->generic);
    if (!b->kernel) return fail(NW_ERR_ARG, "no batch kernel for rows_per_lane=%d", R);
    b->smem = sizeof(uint32_t) * (size_t)(b->packed ? nw::SMEM16_WORDS_PER_WARP : nw::SMEM_WORDS_PER_WARP) * (size_t)b->warps;
    int per_sm = 0;
    {
        const int rc1 = occupancy(b->device, (const void*)b->kernel, b->warps * 32, b->smem, &per_sm);
        if (rc1) return rc1;
    }
    if (per_sm < 1) return fail(NW_ERR_CUDA, "batch kernel does not fit on an SM");
    const int want_per_sm = env_int("NW_CUDA_BATCH_CTAS_PER_SM", b->packed ? 7 : 2);
    per_sm = std::min(per_sm, std::max(1, want_per_sm));
    long long want = (b->npairs + b->warps - 1) / b->warps;
    b->ctas = (int)std::max<long long>(1, std::min<long long>(want, (long long)d.sm_count * per_sm));
    b->scratch_pitch = ((long long)b->len1 + 63) & ~31LL;
    const size_t need = (b->nstrips > 1) ? (size_t)b->scratch_pitch * (size_t)b->ctas * (size_t)b->warps : 0;
    if (need > b->scratch_words) {
        if (b->d_scratch) CK(cudaFree(b->d_scratch));
        b->d_scratch = nullptr;
        CK(cudaMalloc(&b->d_scratch, sizeof(int32_t) * need));
        b->scratch_words = need;
    }
    return NW_OK;
}

static int batch_setup(nw_batch* b)
{
    CK(cudaSetDevice(b->device));
    CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&b->ev0));
    CK(cudaEventCreate(&b->ev1));
    CK(cudaMalloc(&b->d_scores, sizeof(int32_t) * (size_t)std::max<long long>(b->npairs, 1)));
    CK(cudaMalloc(&b->d_bitmap, 8 * sizeof(uint32_t)));
    return batch_pick_kernel(b);
}

extern "C" int nw_batch_create(nw_batch** out, int device, int64_t npairs, int32_t len1, int32_t len2)
{
    if (!out) return fail(NW_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (npairs < 0 || len1 < 0 || len2 < 0) return fail(NW_ERR_ARG, "negative size");
    int rc = ensure_device(device);
    if (rc) return rc;
    nw_batch* b = new (std::nothrow) nw_batch;
    if (!b) return fail(NW_ERR_CUDA, "out of host memory");
    b->device = device;
    b->npairs = npairs;
    b->len1 = len1;
    b->len2 = len2;
    rc = batch_setup(b);
    if (rc) {
        nw_batch_destroy(b);
        return rc;
    }
    *out = b;
    return NW_OK;
}

static int batch_scan(nw_batch* b)
{
    // alphabet of the whole batch, on the device
    CK(cudaMemsetAsync(b->d_bitmap, 0, 8 * sizeof(uint32_t), b->stream));
    const long long t1 = b->npairs * b->len1, t2 = b->npairs * b->len2;
    if (t1 > 0) nw::nw_presence_kernel64<<<296, 256, 0, b->stream>>>(b->S1, t1, b->d_bitmap);
    if (t2 > 0) nw::nw_presence_kernel64<<<296, 256, 0, b->stream>>>(b->S2, t2, b->d_bitmap);
    CK(cudaGetLastError());
    uint32_t bm[8];
    CK(cudaMemcpyAsync(bm, b->d_bitmap, sizeof bm, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    bool seen[256];
    bitmap_to_seen(bm, seen);
    b->generic = !build_code(seen, b->code);
    if (env_int("NW_CUDA_GENERIC", 0)) b->generic = true;
    int rc = batch_pick_kernel(b);
    if (rc) return rc;
    b->uploaded = true;
    return NW_OK;
}

extern "C" int nw_batch_set_scoring(nw_batch* b, const nw_scoring* scoring)
{
    if (!b) return fail(NW_ERR_ARG, "batch is NULL");
    Scoring sc;
    int rc = parse_scoring(scoring, b->len1, b->len2, &sc);
    if (rc) return rc;
    if (sc.local) return fail(NW_ERR_UNSUPPORTED, "batches are global alignments");
    b->sc_match = sc.match; b->sc_mis = sc.mismatch; b->sc_gap = sc.gap;
    b->uploaded = false;       // the kernel choice depends on the weights: upload again
    return NW_OK;
}

extern "C" int nw_batch_upload(nw_batch* b, const int8_t* S1, const int8_t* S2)
{
    if (!b) return fail(NW_ERR_ARG, "batch is NULL");
    const size_t t1 = (size_t)b->npairs * (size_t)b->len1, t2 = (size_t)b->npairs * (size_t)b->len2;
    if ((t1 && !S1) || (t2 && !S2)) return fail(NW_ERR_ARG, "sequence pointer is NULL");
    CK(cudaSetDevice(b->device));
    if (!b->d_S1) CK(cudaMalloc(&b->d_S1, std::max<size_t>(t1, 1)));
    if (!b->d_S2) CK(cudaMalloc(&b->d_S2, std::max<size_t>(t2, 1)));
    if (t1) CK(cudaMemcpyAsync(b->d_S1, S1, t1, cudaMemcpyHostToDevice, b->stream));
    if (t2) CK(cudaMemcpyAsync(b->d_S2, S2, t2, cudaMemcpyHostToDevice, b->stream));
    b->S1 = b->d_S1;
    b->S2 = b->d_S2;
    return batch_scan(b);
}

extern "C" int nw_batch_upload_device(nw_batch* b, const int8_t* d_S1, const int8_t* d_S2)
{
    if (!b) return fail(NW_ERR_ARG, "batch is NULL");
    CK(cudaSetDevice(b->device));
    b->S1 = (const uint8_t*)d_S1;      // borrowed: the caller keeps them alive
    b->S2 = (const uint8_t*)d_S2;
    return batch_scan(b);
}

static int batch_enqueue(nw_batch* b, long long first = 0, long long count = -1)
{
    if (!b->uploaded) return fail(NW_ERR_STATE, "nw_batch_upload has not been called");
    if (count < 0) count = b->npairs - first;
    if (count == 0) return NW_OK;
    nw::BatchParams bp;
    bp.S1 = b->S1 + first * b->len1;
    bp.S2 = b->S2 + first * b->len2;
    bp.scores = b->d_scores + first;
    bp.scratch = b->d_scratch;
    bp.npairs = count;
    bp.scratch_pitch = b->scratch_pitch;
    bp.len1 = b->len1;
    bp.len2 = b->len2;
    bp.nstrips = b->nstrips;
    bp.pad_top = b->pad_top;
    bp.generic = b->generic ? 1 : 0;
    bp.w_match = b->w_match();
    bp.w_mis = b->w_mis();
    bp.gap = b->sc_gap;
    memcpy(bp.code, b->code, 256);
    const int ctas = (int)std::max<long long>(1, std::min<long long>(b->ctas, (count + b->warps - 1) / b->warps));
    b->kernel<<<ctas, b->warps * 32, b->smem, b->stream>>>(bp);
    CK(cudaGetLastError());
    b->ran = true;
    return NW_OK;
}

// Host arrays in, host scores out, in chunks: the H2D copy of chunk c+1 (copy stream) overlaps the kernel of chunk c, and
// every chunk's scores go home as soon as they exist.  The kernel is chosen from the alphabet of chunk 0; the alphabet of
// every later chunk is scanned on the device as it arrives, and if one turns out to need another kernel (a fifth letter)
// the whole batch is simply run again the plain way.
extern "C" int nw_batch_run_host(nw_batch* b, const int8_t* S1, const int8_t* S2, int32_t* scores, int nchunks)
{
    if (!b) return fail(NW_ERR_ARG, "batch is NULL");
    const size_t t1 = (size_t)b->npairs * (size_t)b->len1, t2 = (size_t)b->npairs * (size_t)b->len2;
    if ((t1 && !S1) || (t2 && !S2) || (b->npairs && !scores)) return fail(NW_ERR_ARG, "NULL pointer");
    if (b->npairs == 0) return NW_OK;
    CK(cudaSetDevice(b->device));
    if (nchunks < 1) nchunks = std::max(1, env_int("NW_CUDA_BATCH_CHUNKS", 8));
    nchunks = (int)std::min<long long>(nchunks, b->npairs);
    if (!b->d_S1) CK(cudaMalloc(&b->d_S1, std::max<size_t>(t1, 1)));
    if (!b->d_S2) CK(cudaMalloc(&b->d_S2, std::max<size_t>(t2, 1)));
    b->S1 = b->d_S1;
    b->S2 = b->d_S2;
    if (!b->copy_stream) CK(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    while ((int)b->chunk_ev.size() < nchunks) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        b->chunk_ev.push_back(e);
    }
    CK(cudaStreamSynchronize(b->stream));      // (an earlier run may still read the buffers)
    const long long per = (b->npairs + nchunks - 1) / nchunks;
    auto range = [&](int c, long long* first, long long* count) {
        *first = (long long)c * per;
        *count = std::max<long long>(0, std::min<long long>(per, b->npairs - *first));
    };
    // every copy is enqueued now; the copy stream runs ahead of the kernels
    for (int c = 0; c < nchunks; ++c) {
        long long f, n;
        range(c, &f, &n);
        if (n > 0 && b->len1 > 0)
            CK(cudaMemcpyAsync(b->d_S1 + f * b->len1, (const uint8_t*)S1 + f * b->len1, (size_t)(n * b->len1), cudaMemcpyHostToDevice, b->copy_stream));
        if (n > 0 && b->len2 > 0)
            CK(cudaMemcpyAsync(b->d_S2 + f * b->len2, (const uint8_t*)S2 + f * b->len2, (size_t)(n * b->len2), cudaMemcpyHostToDevice, b->copy_stream));
        CK(cudaEventRecord(b->chunk_ev[c], b->copy_stream));
    }
    // chunk 0 decides the kernel
    uint32_t bm0[8];
    {
        long long f, n;
        range(0, &f, &n);
        CK(cudaStreamWaitEvent(b->stream, b->chunk_ev[0], 0));
        CK(cudaMemsetAsync(b->d_bitmap, 0, 8 * sizeof(uint32_t), b->stream));
        if (n * b->len1 > 0) nw::nw_presence_kernel64<<<296, 256, 0, b->stream>>>(b->S1, n * b->len1, b->d_bitmap);
        if (n * b->len2 > 0) nw::nw_presence_kernel64<<<296, 256, 0, b->stream>>>(b->S2, n * b->len2, b->d_bitmap);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(bm0, b->d_bitmap, sizeof bm0, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
        bool seen[256];
        bitmap_to_seen(bm0, seen);
        b->generic = !build_code(seen, b->code);
        if (env_int("NW_CUDA_GENERIC", 0)) b->generic = true;
        int rc = batch_pick_kernel(b);
        if (rc) return rc;
        b->uploaded = true;
    }
    for (int c = 0; c < nchunks; ++c) {
        long long f, n;
        range(c, &f, &n);
        if (n == 0) continue;
        if (c > 0) {
            CK(cudaStreamWaitEvent(b->stream, b->chunk_ev[c], 0));
            if (n * b->len1 > 0) nw::nw_presence_kernel64<<<296, 256, 0, b->stream>>>(b->S1 + f * b->len1, n * b->len1, b->d_bitmap);
            if (n * b->len2 > 0) nw::nw_presence_kernel64<<<296, 256, 0, b->stream>>>(b->S2 + f * b->len2, n * b->len2, b->d_bitmap);
        }
        int rc = batch_enqueue(b, f, n);
        if (rc) return rc;
        CK(cudaMemcpyAsync(scores + f, b->d_scores + f, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, b->stream));
    }
    uint32_t bm[8];
    CK(cudaMemcpyAsync(bm, b->d_bitmap, sizeof bm, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    bool same = true;
    for (int k = 0; k < 8; ++k) same = same && ((bm[k] & ~bm0[k]) == 0);
    if (!same) {
        // a later chunk brought letters chunk 0 did not have: the letter codes (or the kernel) may not fit them
        int rc = batch_scan(b);       // the whole batch is on the device by now
        if (rc == NW_OK) rc = batch_enqueue(b);
        if (rc == NW_OK) rc = nw_batch_scores(b, scores);
        return rc;
    }
    return NW_OK;
}

extern "C" int nw_batch_run(nw_batch* b)
{
    if (!b) return fail(NW_ERR_ARG, "batch is NULL");
    CK(cudaSetDevice(b->device));
    CK(cudaEventRecord(b->ev0, b->stream));
    int rc = batch_enqueue(b);
    if (rc) return rc;
    CK(cudaEventRecord(b->ev1, b->stream));
    return NW_OK;
}

extern "C" int nw_batch_sync(nw_batch* b)
{
    if (!b) return fail(NW_ERR_ARG, "batch is NULL");
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    return NW_OK;
}

extern "C" int nw_batch_time(nw_batch* b, int iters, float* ms_per_run)
{
    if (!b || !ms_per_run || iters < 1) return fail(NW_ERR_ARG, "bad argument");
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    CK(cudaEventRecord(b->ev0, b->stream));
    for (int i = 0; i < iters; ++i) {
        int rc = batch_enqueue(b);
        if (rc) return rc;
    }
    CK(cudaEventRecord(b->ev1, b->stream));
    CK(cudaEventSynchronize(b->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
    *ms_per_run = ms / iters;
    return NW_OK;
}

extern "C" int nw_batch_scores(nw_batch* b, int32_t* scores)
{
    if (!b || (!scores && b->npairs)) return fail(NW_ERR_ARG, "bad argument");
    if (!b->ran && b->npairs) return fail(NW_ERR_STATE, "no batch has been run");
    CK(cudaSetDevice(b->device));
    if (b->npairs)
        CK(cudaMemcpyAsync(scores, b->d_scores, sizeof(int32_t) * (size_t)b->npairs, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return NW_OK;
}

extern "C" int nw_cuda_batch_scores(const int8_t* S1, const int8_t* S2, int64_t npairs, int32_t len1, int32_t len2,
                                    int32_t* scores, int device)
{
    return nw_cuda_batch_scores_scored(S1, S2, npairs, len1, len2, nullptr, scores, device);
}

extern "C" int nw_cuda_batch_scores_scored(const int8_t* S1, const int8_t* S2, int64_t npairs, int32_t len1, int32_t len2,
                                           const nw_scoring* scoring, int32_t* scores, int device)
{
    nw_batch* b = nullptr;
    int rc = nw_batch_create(&b, device, npairs, len1, len2);
    if (rc == NW_OK && scoring) rc = nw_batch_set_scoring(b, scoring);
    if (rc == NW_OK && npairs >= 4096 && !env_int("NW_CUDA_BATCH_NO_CHUNKS", 0)) {      // overlapped: copies, kernels, scores
        rc = nw_batch_run_host(b, S1, S2, scores, 0);
        char keep2[512];
        memcpy(keep2, g_err, sizeof keep2);
        nw_batch_destroy(b);
        memcpy(g_err, keep2, sizeof keep2);
        return rc;
    }
    if (rc == NW_OK) rc = nw_batch_upload(b, S1, S2);
    if (rc == NW_OK) rc = nw_batch_run(b);
    if (rc == NW_OK) rc = nw_batch_scores(b, scores);
    char keep[512];
    memcpy(keep, g_err, sizeof keep);
    nw_batch_destroy(b);
    memcpy(g_err, keep, sizeof keep);
    return rc;
}

// =====================================================================================================================
// roofline support: integer / DPX pipe rate
// =====================================================================================================================
extern "C" int nw_cuda_dpx_peak(int device, double* giga_lane_ops_per_s, double* sm_clock_mhz)
{
    int rc = ensure_device(device);
    if (rc) return rc;
    CK(cudaSetDevice(device));
    const DeviceState& d = g_dev[device];
    const int ctas = d.sm_count * 2, threads = 512, iters = 4096;
    int* d_out = nullptr;
    unsigned long long* d_clk = nullptr;
    CK(cudaMalloc(&d_out, sizeof(int) * (size_t)ctas * threads));
    CK(cudaMalloc(&d_clk, sizeof(unsigned long long) * 2 * (size_t)ctas));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        nw::nw_dpx_peak_kernel<<<ctas, threads>>>(d_out, d_clk, iters, rep);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best_ms = std::min(best_ms, ms);
    }
    std::vector<unsigned long long> clk(2 * (size_t)ctas);
    CK(cudaMemcpy(clk.data(), d_clk, sizeof(unsigned long long) * clk.size(), cudaMemcpyDeviceToHost));
    double cyc = 0, ns = 0;
    for (int i = 0; i < ctas; ++i) {
        cyc += (double)clk[2 * i];
        ns += (double)clk[2 * i + 1];
    }
    const double ops = (double)ctas * threads * (double)iters * nw::DPX_PEAK_OPS_PER_ITER;
    if (giga_lane_ops_per_s) *giga_lane_ops_per_s = ops / (best_ms * 1e-3) / 1e9;
    if (sm_clock_mhz) *sm_clock_mhz = (ns > 0) ? cyc / ns * 1e3 : 0.0;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    cudaFree(d_clk);
    return NW_OK;
}
