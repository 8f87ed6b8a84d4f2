/*
 * nw_cuda.h -- C ABI of libnw_cuda.so: the B200 (sm_100a) Needleman-Wunsch wavefront fill.
 *
 * This is the drop-in boundary for ONE hot path of EricBAndrews/Fast-Needleman-Wunsch: the fill of the
 * scoring table behind
 *
 *     void needlemanWunsch(dnaArray s1, dnaArray s2, int* t);        // reference: src/serial/serial.cpp:4
 *
 * which the reference's src/common/driver.cpp:28 calls exactly once between its two clock reads.
 * Scoring is the reference's compile-time constants (src/common/needleman-wunsch.hpp:11-13):
 * MATCH +1, MISMATCH 0, GAP -1, int32, linear gap.  s1 runs across the top (columns), s2 down the side (rows);
 * the table is row-major int32 with nCols = n1+1 (src/serial/serial.cpp:6-7,31); the score is the last cell
 * (src/common/driver.cpp:19,35).  Every result is bit-exact against src/serial/serial.cpp.
 *
 * Plain pointers and sizes only.  All functions return NW_OK (0) or a negative NW_ERR_* code; the text of the
 * last error of the calling thread is available from nw_cuda_last_error().  There is NO CPU fallback: without a
 * usable CUDA device every compute entry point fails with NW_ERR_CUDA.
 *
 * The reference-side binding (the 10-line cuda.cpp a maintainer adds next to serial.cpp) is shown in
 * INTEGRATION.md and shipped as fast-needleman-wunsch_b200/csrc/cuda.cpp.
 */
#ifndef NW_CUDA_H
#define NW_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NW_OK 0
#define NW_ERR_CUDA (-1)      /* a CUDA runtime call failed, or no device */
#define NW_ERR_ARG (-2)       /* bad argument (negative size, NULL pointer, unknown mode, ...) */
#define NW_ERR_UNSUPPORTED (-3)
#define NW_ERR_STATE (-4)     /* call sequence error on a plan */

/* memory modes (BASELINE.json configs[2]) */
#define NW_MODE_BOUNDARY 0    /* keep strip boundary rows + right column only; host gets the score        */
#define NW_MODE_FULL 1        /* materialise every cell, as the reference does (src/serial/serial.cpp:31)    */
#define NW_MODE_SCORE 2       /* score only, meeting in the middle: one half of the table is filled forwards and,
                                 concurrently, the other backwards (= forwards on the reversed sequences); the score
                                 is the maximum of F + B over a cut that every path crosses -- the row n2/2, or, when
                                 every strip of both halves can have a scheduler to itself (small tables, or two GPUs),
                                 a staircase from the top right to the bottom left corner, which also takes the strips'
                                 start-up lag off the critical path.  Same number of cell updates, bit-exact score;
                                 only nw_plan_score / nw_plan_best are available on such a plan                   */

/* ------------------------------------------------------------------------------------------------------------
 * Library / device
 * ---------------------------------------------------------------------------------------------------------- */
const char* nw_cuda_version(void);
const char* nw_cuda_last_error(void);
int nw_cuda_device_count(void);                 /* >=0, or NW_ERR_CUDA */
/* Create the context on `device` and warm the kernels up so that the first timed call does not pay for it
 * (driver.cpp:26-30 times the whole call).  Idempotent. */
int nw_cuda_init(int device);
/* Name, SM count and SM clock (MHz) of `device`; any out pointer may be NULL. */
int nw_cuda_device_info(int device, char* name, int name_len, int* sm_count, int* sm_clock_mhz);

/* ------------------------------------------------------------------------------------------------------------
 * One-shot entry points with HOST buffers -- what the reference's callers bind.
 * ---------------------------------------------------------------------------------------------------------- */

/* Replaces needlemanWunsch(dnaArray,dnaArray,int*) (src/serial/serial.cpp:4-36).
 * table: caller-owned HOST memory, (n2+1)*(n1+1) int32 (src/common/driver.cpp:19-23).
 * Mode and GPU count come from the environment (the driver's argv is fixed, src/common/driver.cpp:2):
 *   NW_CUDA_MODE = full (default: every cell written, like the reference) | boundary (only table[size-1], computed
 *                  in NW_MODE_SCORE fashion)
 *   NW_CUDA_GPUS = 1 (default) | 2 | 4 | 8   devices of this process: column strips (full mode; boundary mode beyond two),
 *                  one score-mode half per device (boundary mode on two) */
int nw_cuda_fill(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* table);
int nw_cuda_fill_ex(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* table,
                    int mode, int ngpus);

/* Boundary-only result without a table: *score = H[n2][n1] (what driver.cpp:35 prints). */
int nw_cuda_score(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* score);

/* Boundary-only with the table's last row H[n2][0..n1] and/or last column H[0..n2][n1] (either may be NULL). */
int nw_cuda_boundaries(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2,
                       int32_t* last_row, int32_t* last_col, int32_t* score);

/* Batch of independent pairs (BASELINE.json configs[4]): S1 is npairs x len1 bytes, S2 npairs x len2 bytes,
 * both row-major HOST arrays; scores[p] = NW score of pair p.  Uses device `device`. */
int nw_cuda_batch_scores(const int8_t* S1, const int8_t* S2, int64_t npairs, int32_t len1, int32_t len2,
                         int32_t* scores, int device);

/* ------------------------------------------------------------------------------------------------------------
 * Scoring parameters and local alignment.
 *
 * The reference's only scoring configuration is the three compile-time macros MATCH / MISMATCH / GAP of
 * src/common/needleman-wunsch.hpp:11-13 (1, 0, -1); its README.md:2 names Smith-Waterman as a goal.  A NULL
 * nw_scoring* means exactly those macros.  Any integers are accepted as long as every table value fits int32
 * (checked at plan creation); local alignment additionally needs gap <= 0.
 *   local = 0: global alignment (Needleman-Wunsch), H[0][j] = gap*j, H[i][0] = gap*i, bit-exact against serial.cpp built
 *              with the same three macros.
 *   local = 1: Smith-Waterman, H = max(0, diag + s, up + gap, left + gap) with a zero first row and column; the
 *              result is the best cell and its position (smallest column first, then smallest row; (0,0) if the
 *              best score is 0).  Single device; NW_MODE_BOUNDARY (score + end position + checkpoint rows) or
 *              NW_MODE_FULL (the whole H table as well).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct nw_scoring {
    int32_t match, mismatch, gap;
    int32_t local;
    int32_t reserved[4];     /* must be 0 */
} nw_scoring;

/* nw_cuda_fill_ex / nw_cuda_score with explicit scoring (what csrc/cuda.cpp passes from the reference's macros).
 * end_i / end_j (may be NULL): position of the reported cell -- (n2, n1) for global alignment. */
int nw_cuda_fill_scored(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, int32_t* table,
                        int mode, int ngpus, const nw_scoring* scoring);
int nw_cuda_score_scored(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, const nw_scoring* scoring,
                         int32_t* score, int32_t* end_i, int32_t* end_j);
int nw_cuda_batch_scores_scored(const int8_t* S1, const int8_t* S2, int64_t npairs, int32_t len1, int32_t len2,
                                const nw_scoring* scoring /* global only */, int32_t* scores, int device);

/* ------------------------------------------------------------------------------------------------------------
 * Plans: device-resident state for repeated / timed / multi-GPU runs.
 *
 * A plan owns, on one device, the encoded sequences, the strip boundary rows, the progress flags and (in full
 * mode) the table.  In a column-strip pipeline (reference: src/mpi/mpi-vert.cpp:17, mpi-vert-driver.cpp:35-36)
 * the plan of part `part` of `nparts` owns global table columns [start, start+ncols) with
 *   q = (n1+1)/nparts, start = q*part - (part>0), ncols = q + (part>0) + (part==nparts-1 ? (n1+1)%nparts : 0);
 * column `start` of parts > 0 is the halo received from the left neighbour.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct nw_plan nw_plan;

typedef struct nw_tuning {
    int rows_per_lane;   /* table rows per lane (a strip is 32 x this many rows): 1, 2, 4, 8 for the 32-bit kernels,
                            2, 4, 8, 16 for the packed s16x2 kernels (two rows per register); 0 = automatic */
    int warps_per_cta;   /* 0 = automatic */
    int ctas;            /* persistent grid size; 0 = automatic (<= resident capacity of the device) */
    int reserved[5];
} nw_tuning;

/* mode: NW_MODE_BOUNDARY, NW_MODE_FULL or NW_MODE_SCORE (part 0 of 1: both halves on `device`; part 0 of 2: one half on
 * `device`, the other on `device + 1`, no traffic between them until the final combine; offers nw_plan_upload / run / sync / time /
 * score).  Kernel choice is automatic: packed s16x2 kernels when at most four distinct byte values occur in the two
 * sequences (always true for bdna), 32-bit kernels otherwise. */
int nw_plan_create(nw_plan** out, int device, int32_t n1, int32_t n2, int mode,
                   int part, int nparts, const nw_tuning* tuning /* may be NULL */);
/* The same with scoring parameters (NULL = the reference's macros). */
int nw_plan_create_scored(nw_plan** out, int device, int32_t n1, int32_t n2, int mode,
                          int part, int nparts, const nw_tuning* tuning, const nw_scoring* scoring);
int nw_plan_destroy(nw_plan* p);

/* Sequences from HOST memory (H2D inside).  Every part receives the FULL s1 and s2, like every MPI rank loads
 * both files (src/mpi/mpi-vert-driver.cpp:24-26); the plan keeps only its own slice of s1. */
int nw_plan_upload(nw_plan* p, const int8_t* s1, const int8_t* s2);
/* Sequences already on the plan's device (d_s1: n1 bytes, d_s2: n2 bytes). */
int nw_plan_upload_device(nw_plan* p, const int8_t* d_s1, const int8_t* d_s2);

/* Column-strip chaining inside ONE process: `left`'s right boundary column becomes `right`'s halo; the producing
 * kernel stores it straight into the consumer's device memory over NVLink (peer access is enabled here). */
int nw_plan_connect(nw_plan* left, nw_plan* right);
/* Column-strip chaining ACROSS processes (one process per GPU): the consumer exports a 64-byte CUDA IPC handle of
 * its halo mailbox, the producer imports it. */
int nw_plan_export_mailbox(nw_plan* p, void* handle64);
int nw_plan_import_mailbox(nw_plan* p, const void* handle64, int consumer_device);

/* Enqueue one complete fill on the plan's stream (asynchronous).  nw_plan_sync waits for it. */
int nw_plan_run(nw_plan* p);
int nw_plan_sync(nw_plan* p);
/* Run the fill `iters` times back to back and return the mean device time of one fill in milliseconds,
 * measured with CUDA events on the plan's own stream (sequences already resident). */
int nw_plan_time(nw_plan* p, int iters, float* ms_per_fill);
/* CUDA events on the plan's own stream around whatever is enqueued between the two calls (used to time the parts of a
 * multi-GPU pipeline: each rank brackets its K fills, the job time is the maximum over ranks).  _stop synchronises. */
int nw_plan_timer_start(nw_plan* p);
int nw_plan_timer_stop(nw_plan* p, float* ms);
/* Device time of the most recent nw_plan_run, CUDA events around the kernels only. */
int nw_plan_last_ms(nw_plan* p, float* ms);
/* Number of kernels one nw_plan_run launches. */
int nw_plan_launches_per_run(nw_plan* p, int* n);

/* Results (D2H inside; synchronises the plan's stream).  Valid on the LAST part of a pipeline. */
int nw_plan_score(nw_plan* p, int32_t* score);
/* Reported cell: global alignment (n2, n1) and H[n2][n1]; local alignment the best cell (see nw_scoring). */
int nw_plan_best(nw_plan* p, int32_t* score, int32_t* end_i, int32_t* end_j);
int nw_plan_last_row(nw_plan* p, int32_t* last_row /* ncols of this part, H values */);
int nw_plan_last_col(nw_plan* p, int32_t* last_col /* n2+1 H values of this part's right-most column */);
/* Full mode only: copy this part's columns into a HOST table with row pitch (n1+1). */
int nw_plan_table_to_host(nw_plan* p, int32_t* table);
/* Full mode only: device pointer and row pitch (elements) of this part's table (rows 0..n2, local columns). */
int nw_plan_table_device(nw_plan* p, int32_t** d_table, int64_t* pitch);
/* Alignment (the reference defines the gap code and a printer -- README.md:8, src/common/helper.cpp:27-34 -- but never
 * produces an alignment).  Walks back from H[n2][n1] to H[0][0]: diagonal if H[i][j] == H[i-1][j-1] + (s1[j-1]==s2[i-1] ?
 * match : mismatch), else up if H[i][j] == H[i-1][j] + gap, else left.  a1 / a2 (HOST, capacity n1+n2 each) receive the
 * gapped s1 / s2 left to right, gap = 0; *len their common length.
 *   full-table plan (single part, table resident): walks the materialised table; on a local-alignment plan from the best
 *   cell back to the first cell with H = 0 (the aligned segments end at nw_plan_best's cell and begin that many letters
 *   earlier);
 *   boundary-mode plan: NO table -- the path is recovered tile by tile from the strip boundary rows the fill left in HBM
 *   (every tile between two checkpoint rows is recomputed from its exact top row and left column, then walked).
 * nw_plans_traceback does the same for the connected parts 0..nparts-1 of a column-strip pipeline on ONE device that
 * have all run the same fill: their halo columns are checkpoint columns, so a tile is strip_rows x (part width). */
int nw_plan_traceback(nw_plan* p, int8_t* a1, int8_t* a2, int32_t* len);
int nw_plans_traceback(nw_plan* const* parts, int nparts, int8_t* a1, int8_t* a2, int32_t* len);
/* One-shot, HOST buffers, no table anywhere: column parts of ~NW_CUDA_ALIGN_TILE (default 4096) columns on device 0,
 * filled one after the other, then nw_plans_traceback.  scoring may be NULL (global alignment only); score may be NULL. */
int nw_cuda_align(const int8_t* s1, int32_t n1, const int8_t* s2, int32_t n2, const nw_scoring* scoring,
                  int8_t* a1, int8_t* a2, int32_t* len, int32_t* score);
/* Strip boundary rows kept in HBM (checkpoint rows): number of strips, rows per strip and, per strip k, the
 * H values of table row  n2 - (nstrips-1-k)*strip_rows  copied to HOST (ncols of this part). */
int nw_plan_strip_info(nw_plan* p, int* nstrips, int* strip_rows, int* rows_per_lane, int* warps, int* ctas);
int nw_plan_strip_row(nw_plan* p, int strip, int32_t* row);
/* Trace of the most recent fill: per strip, the device %globaltimer (ns) at which its warp had the first block of its
 * top boundary row (i.e. when it really started) and at which it finished, and (sm_cycles, may be NULL) the SM clock
 * cycles between the two -- cycles / ns is the SM clock the strip really ran at.  Consecutive start times give the strip
 * start-up lag that bounds a single-pair fill; packed kernels only (zeros otherwise). */
int nw_plan_strip_times(nw_plan* p, int64_t* start_ns, int64_t* end_ns, int64_t* sm_cycles);

/* ------------------------------------------------------------------------------------------------------------
 * Batch plans (independent pairs; shards trivially, one pair-set per GPU).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct nw_batch nw_batch;
int nw_batch_create(nw_batch** out, int device, int64_t npairs, int32_t len1, int32_t len2);
int nw_batch_set_scoring(nw_batch* b, const nw_scoring* scoring);   /* before nw_batch_upload; global alignment only */
int nw_batch_destroy(nw_batch* b);
int nw_batch_upload(nw_batch* b, const int8_t* S1, const int8_t* S2);            /* HOST -> device */
int nw_batch_upload_device(nw_batch* b, const int8_t* d_S1, const int8_t* d_S2);
int nw_batch_run(nw_batch* b);
int nw_batch_sync(nw_batch* b);
int nw_batch_time(nw_batch* b, int iters, float* ms_per_run);
int nw_batch_scores(nw_batch* b, int32_t* scores /* HOST, npairs */);
/* HOST arrays in, HOST scores out, in `nchunks` chunks (0 = NW_CUDA_BATCH_CHUNKS, default 8): the H2D copy of chunk c+1
 * overlaps the kernel of chunk c and every chunk's scores leave as soon as they exist.  Pinned host memory makes the
 * copies asynchronous; pageable memory works too (the copy of the next chunk then blocks the host while the GPU computes). */
int nw_batch_run_host(nw_batch* b, const int8_t* S1, const int8_t* S2, int32_t* scores, int nchunks);

/* ------------------------------------------------------------------------------------------------------------
 * Roofline support: measured integer/DPX pipe rate of `device`.
 * Runs independent VIADDMNMX/VIMNMX chains on every SM and returns lane-instructions per second (G/s) and the
 * SM clock seen during the run (from the kernel's own clock64/globaltimer readings).
 * ---------------------------------------------------------------------------------------------------------- */
int nw_cuda_dpx_peak(int device, double* giga_lane_ops_per_s, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* NW_CUDA_H */
