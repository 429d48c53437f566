/*
 * wfa_gpu.h -- public C API of the B200-native gap-affine wavefront aligner.
 *
 * Drop-in for WFA-GPU's `#include "include/wfa_gpu.h"`: same type names, same
 * field order/sizes (x86-64), same function names and argument meaning.  Each
 * declaration cites the reference interface it replaces (paths relative to
 * the reference tree).  Users compile with `-I <repo>` and link `-lwfagpu`.
 *
 * Behavioural differences: there is no CPU fallback -- pairs
 * that exceed `max_error` or contain non-ACGT bases are re-dispatched on the
 * GPU; `threads_per_block` / `num_workers` are hints (except that
 * `threads_per_block` is the band width in banded mode, as in the reference).
 */
#ifndef WFA_GPU_H
#define WFA_GPU_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ types */

/* replaces lib/affine_penalties.h:24-30 (match is always 0) */
typedef struct {
    int x; /* mismatch   */
    int o; /* gap open   */
    int e; /* gap extend */
} affine_penalties_t;

/* replaces utils/sequences.h:28-36; 48 bytes, offsets are absolute positions
 * in the host sequence buffer, every sequence starts 4-byte aligned and is
 * followed by at least one NUL */
typedef struct {
    size_t text_offset;
    size_t pattern_offset;
    size_t text_offset_packed;
    size_t pattern_offset_packed;
    unsigned int text_len;
    unsigned int pattern_len;
    bool has_N;
} sequence_pair_t;

/* replaces lib/wfa_types.h:28-56 */
#define MAX_SEQ_LEN (1UL << 15)
typedef int16_t wfa_offset_t;
typedef uint32_t bt_vector_t;
typedef uint32_t bt_prev_t;
typedef struct {
    bt_vector_t backtrace;
    bt_prev_t prev;
} wfa_backtrace_t;
typedef enum { OP_NOOP = 0, OP_INS = 1, OP_SUB = 2, OP_DEL = 3 } affine_op_t;

/* replaces lib/alignment_results.h:30-48 */
typedef struct {
    char *buffer;              /* NUL-terminated RLE CIGAR, e.g. "12M1X3I40M" */
    size_t buffer_size;
    size_t last_free_position;
} wfa_cigar_t;

typedef struct {
    bool finished;
    int distance;
    wfa_backtrace_t backtrace;
    int num_bt_blocks;
} alignment_result_t;

typedef struct {
    unsigned int error; /* gap-affine score (positive) */
    wfa_cigar_t cigar;
} wfa_alignment_result_t;

/* replaces lib/alignment_parameters.h:29-58 */
#define BAND_NONE (-1)
typedef struct {
    int max_error;         /* wavefront step budget of the first GPU pass        */
    int threads_per_block; /* hint; in banded mode: the band width (diagonals)   */
    int num_workers;       /* hint                                               */
    int band;              /* <=0: exact; >0: re-centre the band every `band` scores */
    size_t batch_size;     /* pairs per device batch; 0 = one batch              */
    size_t num_alignments;
    affine_penalties_t penalties;
    bool compute_cigar;
} wfa_alignment_options_t;

/* replaces lib/aligner.h:30-43 */
#define WFA_ALIGN_32_BITS(x) ((x) + (4 - ((x) % 4)))
typedef char wfagpu_seqbuf_t;
typedef struct {
    wfagpu_seqbuf_t *sequences_buffer;
    size_t sequences_buffer_len;
    sequence_pair_t *sequences_metadata;
    size_t sequences_metadata_len;
    size_t num_sequence_pairs;
    wfa_alignment_result_t *results;
    int64_t last_sequence_pair_idx;
    wfa_alignment_options_t alignment_options;
} wfagpu_aligner_t;

/* -------------------------------------------------------------- functions */

/* lib/aligner.h:49-62, lib/aligner.c:114-263 */
bool wfagpu_initialize_aligner(wfagpu_aligner_t *aligner);
bool wfagpu_add_sequences(wfagpu_aligner_t *aligner, const char *query, const char *target);
bool wfagpu_initialize_parameters(wfagpu_aligner_t *aligner, affine_penalties_t penalties);
bool wfagpu_set_batch_size(wfagpu_aligner_t *aligner, size_t batch_size);
bool wfagpu_align(wfagpu_aligner_t *aligner);
void wfagpu_destroy_aligner(wfagpu_aligner_t *aligner);

/* lib/align.cuh:35-47 -- batch drivers the reference CLI calls directly.
 * They overwrite sequences_metadata[].{pattern,text}_offset_packed, write
 * alignment_results[i].error and (launch_alignments) append the CIGAR text. */
void launch_alignments(char *sequences_buffer, const size_t sequences_buffer_size,
                       sequence_pair_t *const sequences_metadata,
                       wfa_alignment_result_t *const alignment_results,
                       wfa_alignment_options_t options, bool check_correctness);
void launch_alignments_distance(char *sequences_buffer, const size_t sequences_buffer_size,
                                sequence_pair_t *const sequences_metadata,
                                wfa_alignment_result_t *const alignment_results,
                                wfa_alignment_options_t options, bool check_correctness);

/* lib/alignment_results.h:54-59 */
bool initialize_wfa_results(wfa_alignment_result_t **results, const size_t num_alignments,
                            const size_t cigar_length);
bool destroy_wfa_results(wfa_alignment_result_t *results, const size_t num_alignments);

/* utils/device_query.cuh:29-33 */
void get_num_cuda_devices(int *n);
char *get_cuda_dev_name(int dev); /* caller frees */
int get_cuda_SM_count(int dev);
void get_cuda_capability(int dev, int *major, int *minor);

/* utils/cigar.h:39-41 */
bool insert_ops(wfa_cigar_t *const cigar, const char op, const unsigned int rep);

/* ------------------------------------- header helpers the examples rely on */

/* lib/alignment_parameters.h:60-71 */
static inline int wfa_get_threads_per_alignment(const size_t max_error)
{
    const size_t width = 2 * max_error + 1;
    int t = 64;
    while (t < 1024 && (size_t)(2 * t) < width) t *= 2;
    return t;
}

/* lib/alignment_parameters.h:73-81 */
static inline int get_num_workers(const int num_threads)
{
    const int warps = num_threads / 32;
    return get_cuda_SM_count(0) * (32 / (warps > 0 ? warps : 1));
}

/* lib/alignment_parameters.h:83-106 */
static inline void wfagpu_set_default_options(wfa_alignment_options_t *opt,
                                              sequence_pair_t *sequences_metadata,
                                              affine_penalties_t penalties, size_t num_alignments)
{
    int slen = (int)(sequences_metadata[0].pattern_len > sequences_metadata[0].text_len
                         ? sequences_metadata[0].pattern_len : sequences_metadata[0].text_len);
    slen = (int)(slen * 0.1);
    int pmax = penalties.x > penalties.o ? penalties.x : penalties.o;
    if (penalties.e > pmax) pmax = penalties.e;
    int max_error = slen * pmax;
    if (max_error < 50) max_error = 50;
    opt->max_error = max_error;
    opt->threads_per_block = wfa_get_threads_per_alignment((size_t)max_error);
    opt->num_workers = get_num_workers(opt->threads_per_block);
    opt->band = BAND_NONE;
    opt->num_alignments = num_alignments;
    opt->batch_size = num_alignments > 10 ? num_alignments / 10 : num_alignments;
    opt->penalties = penalties;
    opt->compute_cigar = false;
}

#ifdef __cplusplus
}
#endif

#endif /* WFA_GPU_H */
