/*
 * wfagpu_b200.h -- C-ABI between the C host code and the hand-written sm_100a
 * CUDA layer (wfa-gpu_b200/csrc), plus the B200-specific extensions of the
 * public API.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * This is the boundary a reference maintainer would bind instead of
 * lib/sequence_packing.cuh:27-40 (pack_sequences_gpu_async) and
 * lib/sequence_alignment.cuh:29-114 (the allocate_xxx / reset_xxx helpers,
 * launch_alignments_async, launch_alignments_distance_async, copyInResults),
 * see INTEGRATION.md.
 */
#ifndef WFAGPU_B200_H
#define WFAGPU_B200_H

#include <stddef.h>
#include <stdint.h>
#include "wfa_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Longest sequence wfagpu_add_sequences accepts (the reference: MAX_SEQ_LEN - 1 = 32767). */
/* The large tier stages both packed sequences of a pair in one CTA's shared memory (227 KB): 2 x len / 2 bytes. */
#define WFAGPU_MAX_SEQ_LEN ((size_t)220000)

/* ---------------------------------------------------------- step table --- */
/* Per-score control record.  It depends on the penalties and the step budget
 * only (the existence logic of lib/kernels/sequence_alignment_kernel.cu:584-631
 * never looks at the sequences), so the host builds it once per launch. */
#define WFAGPU_STEP_NULL 0u
#define WFAGPU_STEP_M    1u
#define WFAGPU_STEP_MDI  2u
typedef struct {
    uint32_t row_off; /* offset of this score's decision row, in 16-byte units */
    uint16_t n;       /* half width of the diagonal range [-n, n] after this step */
    uint16_t kind;    /* WFAGPU_STEP_*                                          */
} wfagpu_step_t;

/* Builds the table for scores 0 .. d_end-1 (returns d_end; tab may be NULL to
 * size it).  `arena_units` receives the number of 16-byte decision units one
 * alignment can write.  Budget rule = reference: MDI steps stop once
 * steps >= max_steps-1 (sequence_alignment_kernel.cu:584, steps starts at 1).
 * banded_win == 0: exact kernels (n = computed half width, row_off = decision-byte row of the
 * warp-per-pair and large-tier kernels; the CTA kernels keep ring snapshots instead, laid out by the device layer);
 * banded_win  > 0: banded kernels (n = number of M/I/D steps so far, rows of
 * ceil(win/16) units).  A decision row holds one byte per diagonal. */
int wfagpu_build_step_table(int x, int o, int e, int max_steps, int max_dist, int banded_win,
                            wfagpu_step_t *tab, uint64_t *arena_units);

/* ------------------------------------------------------- device batches --- */
/* One pair as the device sees it (built by the host from sequence_pair_t). */
typedef struct {
    uint32_t p_ascii; /* byte offset of the pattern in the batch ASCII buffer */
    uint32_t t_ascii;
    uint32_t p_word;  /* u32 offset of the packed pattern in the packed buffer (multiple of 4) */
    uint32_t t_word;
    uint32_t plen;
    uint32_t tlen;
    uint32_t flags;   /* WFAGPU_PAIR_* (has_N written by the pack kernel)      */
    uint32_t reserved;
} wfagpu_pair_t;
#define WFAGPU_PAIR_HAS_N 1u

/* Per-pair result record written by the alignment kernels. */
typedef struct {
    int32_t distance;  /* score if finished                                     */
    uint32_t status;   /* WFAGPU_ST_*                                           */
    uint32_t ops_off;  /* u32 offset of the packed op stream in the ops pool    */
    uint32_t n_ops;    /* number of 2-bit backtrace ops, stored newest first    */
} wfagpu_pair_out_t;
#define WFAGPU_ST_FINISHED   1u
#define WFAGPU_ST_OVERBUDGET 2u /* needs a larger wavefront budget (re-dispatch) */
#define WFAGPU_ST_NEEDS_ASCII 4u /* flagged by the packer: byte-compare kernel    */
#define WFAGPU_ST_FAILED 8u      /* the GPU cannot finish this pair (> 60000 wavefront steps or wavefronts wider than
                                  * one CTA holds); set by wfagpu_device_download, the other pairs keep their results */

/* Where a pair's CIGAR text sits in the text pool returned by wfagpu_device_download_text. */
typedef struct {
    uint32_t off; /* byte offset in the text pool */
    uint32_t len; /* characters (no terminator); 0 = none */
} wfagpu_cigar_ref_t;

typedef struct wfagpu_device wfagpu_device_t; /* opaque: streams, buffers, arenas of one GPU */

typedef struct {
    int x, o, e;
    int max_steps;   /* budget of the first pass (reference `max_error`)        */
    int band;        /* <=0 exact                                               */
    int band_width;  /* banded window width (reference threads_per_block)       */
    int with_cigar;
    int threads_hint;
    int workers_hint;
} wfagpu_plan_t;

/* Statistics of the last batch (kernel times from CUDA events on the launch
 * stream; counts of kernels launched). */
typedef struct {
    float ms_h2d, ms_pack, ms_align, ms_d2h, ms_total;
    uint32_t launches;        /* kernels launched for this batch                */
    uint32_t redispatched;    /* pairs that needed a larger budget              */
    uint32_t ascii_pairs;     /* pairs routed to the byte-compare kernel        */
    uint64_t cells;           /* wavefront cells computed (0 unless profiling)  */
    uint64_t h2d_bytes, d2h_bytes;
    float ms_wavefront;       /* the first pass's wavefront kernel alone (0 if it ran in several sub-launches) */
    uint32_t failed_pairs;    /* pairs reported as WFAGPU_ST_FAILED                                            */
    uint32_t pending_pairs;   /* after wfagpu_device_wait: pairs the first pass left to the re-dispatch tier   */
    uint32_t n_cap, cta_threads, ctas, d_end; /* launch shape of the first pass: ring half width, CTA size, grid, score limit */
    uint32_t reserved1;
} wfagpu_batch_stats_t;

/* Leases a context of CUDA device `dev`: an idle one from the pool (with its grown buffers) or a new one; the
 * caller owns it until wfagpu_device_release.  Two leases never share streams, slots or staging buffers, so
 * threads (or two workers on one GPU) do not interfere.  NULL on failure (message on stderr); never falls
 * back to the CPU.  wfagpu_device_close_all frees every context (none may be in use). */
wfagpu_device_t *wfagpu_device_open(int dev);
void wfagpu_device_release(wfagpu_device_t *d);
void wfagpu_device_close_all(void);

/*
 * One batch on one device, in three phases on the slot's own stream (two slots
 * per device let the host overlap batch c+1's copies with batch c's kernels):
 *   upload    H2D of the batch's ASCII (pairs[i].*_ascii are offsets into
 *             `ascii`), pair descriptors and the longest-first schedule;
 *   align     pack kernel, score-bound kernel, alignment kernel (+ traceback, CIGAR text); may be
 *             repeated on a resident batch.  Returns with the alignment kernels queued; for long
 *             reads it first waits for the bound kernel (the rings of the pass are sized from the
 *             batch's own bounds: 4 bytes per pair come back to the host), the rest is asynchronous;
 *   download  waits, finishes over-budget / non-ACGT pairs on the GPU
 *             (re-dispatch with a doubled budget, byte-compare kernel) and copies
 *             out[i] and the packed op streams back (`*ops` points into pinned
 *             memory owned by the slot, valid until the slot's next download).
 * All return 0 on success; failures print to stderr -- there is no CPU path.
 */
int wfagpu_device_upload(wfagpu_device_t *d, int slot, const char *ascii, size_t ascii_bytes,
                         const wfagpu_pair_t *pairs, size_t n);
int wfagpu_device_align(wfagpu_device_t *d, int slot, size_t n, const wfagpu_plan_t *plan,
                        int resident);
int wfagpu_device_download(wfagpu_device_t *d, int slot, size_t n, wfagpu_pair_out_t *out,
                           uint32_t **ops, size_t *ops_used, uint32_t *pair_flags);
/* Waits for the slot's stream; event-timed durations (milliseconds) of the pack kernel and of the
 * first pass of the alignment: score bounds + wavefronts + traceback + CIGAR text. */
int wfagpu_device_wait(wfagpu_device_t *d, int slot, float *ms_pack, float *ms_align);
/* After wfagpu_device_download: CIGAR text emitted on the device (one warp per pair, the
 * reference's exact format) and compacted; `*text` / `*refs` point into pinned memory owned by
 * the slot.  Replaces the host loop over recover_cigar_affine (utils/wfa_cpu.c:88-107). */
int wfagpu_device_download_text(wfagpu_device_t *d, int slot, size_t n, const char **text, size_t *text_bytes,
                                const wfagpu_cigar_ref_t **refs);
void wfagpu_device_last_stats(wfagpu_device_t *d, int slot, wfagpu_batch_stats_t *st);
/* Page-locks a caller buffer (e.g. the aligner's sequence buffer) so that the
 * H2D copies run as asynchronous DMA. */
int wfagpu_host_register(void *ptr, size_t bytes);
int wfagpu_host_unregister(void *ptr);
/* 1 if `ptr` is page-locked memory. */
int wfagpu_host_is_pinned(const void *ptr);
/* Page-locked staging area of a slot for callers with a pageable sequence buffer (see driver.c). */
char *wfagpu_device_staging(wfagpu_device_t *d, int slot, size_t bytes);
/* Page-locked zeroed host memory (NULL without a usable CUDA device: callers fall back to calloc). */
void *wfagpu_host_alloc(size_t bytes);
void wfagpu_host_free(void *p);
int wfagpu_device_sm_count(wfagpu_device_t *d);

/* Builds the device-side descriptors of pairs [from, from+n) of a host buffer
 * (also rewrites their *_offset_packed like lib/align.cu:103-115). `base_out` /
 * `bytes_out`: the slice of the buffer to hand to wfagpu_device_upload. */
int wfagpu_pairs_from_metadata(sequence_pair_t *meta, size_t from, size_t n, size_t buf_size,
                               wfagpu_pair_t *pairs, size_t *base_out, size_t *bytes_out);

/* Pack kernel alone (tests, replaces prepare_pack_sequences_gpu +
 * pack_sequences_gpu_async, lib/sequence_packing.cu:27-116): uploads, packs and
 * returns the packed words (layout: word j = bases [8j, 8j+16), first base in
 * bits 31:30, code (c&6)>>1) and the has_N flags. */
int wfagpu_device_pack_only(wfagpu_device_t *d, const char *ascii, size_t ascii_bytes,
                            wfagpu_pair_t *pairs, size_t n, uint32_t *packed_out,
                            size_t packed_words);

/* ------------------------------------------------------------ host side --- */
/* CIGAR text from a packed op stream (newest op first, 16 ops per u32, op j
 * of a word in bits 2j+1:2j).  Same text as utils/cigar.c:96-272 produces
 * from the reference's backtrace chain.  Appends to `cigar`. */
bool wfagpu_ops_to_cigar(const char *pattern, size_t plen, const char *text, size_t tlen,
                         int distance, const uint32_t *ops, uint32_t n_ops, wfa_cigar_t *cigar);

/* Appends formatted CIGAR text to a result buffer (growing it like insert_ops does). */
bool wfagpu_cigar_append(wfa_cigar_t *cigar, const char *text, size_t len);

/* Device selection for launch_alignments*: "0", "0,1,2", "all", or a count
 * ("n:4").  Default (NULL/unset): environment WFAGPU_DEVICES, else device 0. */
void wfagpu_set_devices(const char *spec);

/* Chunking of a job over `n_devices` GPUs (pure function, used by launch_alignments*). */
void wfagpu_plan_chunks(size_t n, size_t batch_size, int n_devices, size_t ascii_span,
                        size_t *chunk_out, size_t *n_chunks_out);

/* How the chunks of a job are dealt to `nworkers` host threads (one per device-list entry): every worker owns a contiguous
 * share [share_next[i], share_end[i]) and takes `want` pairs from its front; a worker whose share is used up takes `steal`
 * pairs from the back of the fullest share.  Pure functions on the share table (the caller serialises); false = nothing left. */
void wfagpu_plan_shares(size_t n, int nworkers, size_t *share_next, size_t *share_end);
bool wfagpu_share_take(size_t *share_next, size_t *share_end, int nworkers, int index, size_t want, size_t steal,
                       size_t *from, size_t *n);

/* Host threads this library may use per call (result loop, staging copies, generators); 0 = OpenMP default.
 * torchrun pins OMP_NUM_THREADS=1 per rank: a rank that works alone can lift that. */
void wfagpu_set_host_threads(int n);

/* Stats of the last launch_alignments* call (summed over batches/devices). */
typedef struct {
    double wall_s;
    double gpu_align_ms;  /* sum of alignment-kernel time over batches (max over devices per batch not tracked) */
    double gpu_pack_ms;
    uint64_t launches;
    uint64_t redispatched;
    uint64_t ascii_pairs;
    uint64_t h2d_bytes, d2h_bytes;
    int devices;          /* workers (one leased context each) */
    int staged;           /* 1: the caller's buffer was pageable, chunks went through page-locked staging */
    uint64_t failed_pairs; /* pairs the GPU could not finish (results[i].error == UINT_MAX)               */
    uint64_t checked, incorrect; /* check_correctness: pairs validated / found wrong                      */
} wfagpu_run_stats_t;
/* Statistics / outcome of the calling thread's last launch_alignments* call (thread-local). */
void wfagpu_last_run_stats(wfagpu_run_stats_t *st);
/* false when that call failed on the GPU or left pairs unaligned (nothing is computed on the CPU) */
bool wfagpu_last_launch_ok(void);
/* Device list syntax of wfagpu_set_devices / WFAGPU_DEVICES (pure function; -1 on a bad list). */
int wfagpu_parse_devices(const char *spec, int visible, int *devs, int max_devs);
/* Independent re-computation of the scores of the batch resident in `slot` (check_correctness): score only,
 * one-diagonal-per-thread kernels, launch bound only -- no per-pair bounds, packed-SIMD kernel, snapshots or
 * provisioning hints.  scores[i] = -1 for a pair it could not finish. */
int wfagpu_device_rescore(wfagpu_device_t *d, int slot, size_t n, const wfagpu_plan_t *plan, int32_t *scores);

/* Makes room for `bytes` more sequence bytes and `pairs` more pairs in one step (readers, generators). */
bool wfagpu_reserve(wfagpu_aligner_t *aligner, size_t bytes, size_t pairs);

/* Clears errors and CIGAR text of a previous wfagpu_align (the reference, like
 * this library, appends to results[i].cigar) so an aligner can be re-aligned. */
void wfagpu_reset_results(wfagpu_aligner_t *aligner);

/* CLI input formats (replace utils/sequence_reader.c:137-392): append the pairs of a .seq file
 * (">pattern" / "<text" lines) or of two FASTA files (n-th query record with n-th target record)
 * to the aligner. max_pairs == 0 reads everything. Return the number of pairs, -1 on error. */
long wfagpu_read_seq_file(wfagpu_aligner_t *aligner, const char *path, size_t max_pairs);
long wfagpu_read_fasta_files(wfagpu_aligner_t *aligner, const char *query_path, const char *target_path,
                             size_t max_pairs);

/* Incremental readers (streaming): every call appends the next `max_pairs` pairs (0 = all that are left) to the
 * aligner's page-locked buffer; returns how many (0 at the end of the input, -1 on a format error). */
typedef struct wfagpu_reader wfagpu_reader_t;
wfagpu_reader_t *wfagpu_reader_open_seq(const char *path);
wfagpu_reader_t *wfagpu_reader_open_fasta(const char *query_path, const char *target_path);
long wfagpu_reader_next(wfagpu_reader_t *r, wfagpu_aligner_t *aligner, size_t max_pairs);
long wfagpu_reader_total_bytes(const wfagpu_reader_t *r);
void wfagpu_reader_close(wfagpu_reader_t *r);
/* Forgets the pairs of an aligner but keeps its (page-locked) buffers, so that the next window of a stream reuses them. */
void wfagpu_clear_sequences(wfagpu_aligner_t *aligner);

/* `-c`: true iff `cigar` is a valid global alignment of (pattern, text) whose gap-affine cost is
 * `error` (replaces check_cigar_edit + check_affine_distance, utils/verification.c:27-146). */
bool wfagpu_check_result(const char *pattern, size_t plen, const char *text, size_t tlen,
                         affine_penalties_t pen, unsigned int error, const char *cigar);

/* The reference library's generic validators, same names, argument order (text first) and input format
 * (utils/verification.h:37-58): the op string may be UNROLLED ("MMXMMI", what the reference's recover_cigar
 * returns and lib/align.cu:284-293 passes) or the run-length text of results[i].cigar.buffer ("2M1X2M1I").
 * check_cigar_edit: the ops are a global alignment of (pattern, text); check_affine_distance: their gap-affine
 * cost equals `distance`. */
bool check_cigar_edit(const char *text, const char *pattern, const int tlen, const int plen, const char *curr_cigar);
bool check_affine_distance(const char *text, const char *pattern, const int tlen, const int plen, const int distance,
                           const affine_penalties_t penalties, const char *cigar);
/* recover_cigar (utils/verification.h:52-58): unrolled op string from a reference-format backtrace chain
 * (calloc'ed, caller frees). */
char *recover_cigar(const char *text, const char *pattern, const size_t tlen, const size_t plen,
                    wfa_backtrace_t final_backtrace, wfa_backtrace_t *offloaded_backtraces_array,
                    alignment_result_t result);
/* Unrolled op string of run-length CIGAR text (malloc'ed, caller frees; NULL on malformed text). */
char *wfagpu_unroll_cigar(const char *rle);

/* Deterministic synthetic pairs (text uniform over ACGT, pattern =
 * text with ceil(L*err) edits, each uniformly mismatch / 1-base deletion /
 * 1-base insertion; splitmix64 seeded).  Appends n pairs to the aligner. */
bool wfagpu_synth_add_pairs(wfagpu_aligner_t *aligner, uint64_t seed, size_t n, int length,
                            double err_lo, double err_hi);

#ifdef __cplusplus
}
#endif
#endif
