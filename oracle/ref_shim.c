/*
 * ref_shim.c -- thin C entry points around the UNMODIFIED reference sources,
 * compiled where they lie under /root/reference into oracle/_ref/ (see
 * oracle/Makefile, target `ref`).  TEST INFRASTRUCTURE ONLY.
 *
 * Exposes:
 *   - the reference's CPU path for this hot path: compute_alignments_cpu_threaded
 *     / compute_distance_cpu_threaded (utils/wfa_cpu.c:30-164), i.e. WFA2-lib
 *     v2.3 configured as utils/wfa_cpu.c:40-48 and looped with OpenMP, run on
 *     every pair by presenting all pairs as `!finished`;
 *   - the reference's own host CIGAR decoder recover_cigar_affine
 *     (utils/cigar.c:96-272) on a caller-provided backtrace chain, used to pin
 *     the oracle's decoder restatement.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdbool.h>
#include <omp.h>

#include "utils/sequences.h"
#include "lib/wfa_types.h"
#include "lib/alignment_results.h"
#include "utils/wfa_cpu.h"
#include "utils/cigar.h"

/* Lay the pairs out exactly like wfagpu_add_sequences (lib/aligner.c:127-166). */
static size_t align4(size_t x) { return x + (4 - (x % 4)); }

typedef struct {
    char *buf;
    size_t buf_len;
    sequence_pair_t *meta;
} ref_batch_t;

static int build_batch(ref_batch_t *b, int n, const char **patterns, const int *plens,
                       const char **texts, const int *tlens)
{
    size_t total = 0;
    for (int i = 0; i < n; i++) {
        total = align4(total + (size_t)plens[i] + 1);
        total = align4(total + (size_t)tlens[i] + 1);
    }
    b->buf = (char *)calloc(total + 64, 1);
    b->meta = (sequence_pair_t *)calloc((size_t)n, sizeof(sequence_pair_t));
    if (!b->buf || !b->meta) return -1;
    b->buf_len = total + 64;
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        b->meta[i].pattern_offset = off;
        b->meta[i].pattern_len = (unsigned)plens[i];
        memcpy(b->buf + off, patterns[i], (size_t)plens[i]);
        off = align4(off + (size_t)plens[i] + 1);
        b->meta[i].text_offset = off;
        b->meta[i].text_len = (unsigned)tlens[i];
        memcpy(b->buf + off, texts[i], (size_t)tlens[i]);
        off = align4(off + (size_t)tlens[i] + 1);
    }
    return 0;
}

/*
 * Run the reference CPU path on n pairs. errors[i] receives
 * wfa_alignment_result_t.error; when with_cigar, cigars[i] receives a
 * malloc'ed copy of the CIGAR text (caller frees with ref_free).
 * threads <= 0 keeps the OpenMP default. Returns the number of pairs aligned.
 */
int ref_cpu_align_batch(int n, const char **patterns, const int *plens,
                        const char **texts, const int *tlens,
                        int x, int o, int e, int with_cigar, int adaptive, int threads,
                        int *errors, char **cigars)
{
    ref_batch_t b;
    if (build_batch(&b, n, patterns, plens, texts, tlens)) return -1;
    alignment_result_t *res = (alignment_result_t *)calloc((size_t)n, sizeof(alignment_result_t));
    wfa_alignment_result_t *out = NULL;
    initialize_wfa_results(&out, (size_t)n, 64);
    if (threads > 0) omp_set_num_threads(threads);
    int done;
    if (with_cigar) {
        done = compute_alignments_cpu_threaded(n, 0, res, out, b.meta, b.buf, NULL, 0,
                                               x, o, e, adaptive != 0);
    } else {
        done = compute_distance_cpu_threaded(n, 0, res, out, b.meta, b.buf,
                                             x, o, e, adaptive != 0);
    }
    for (int i = 0; i < n; i++) {
        errors[i] = (int)out[i].error;
        if (with_cigar && cigars) cigars[i] = strdup(out[i].cigar.buffer);
    }
    destroy_wfa_results(out, (size_t)n);
    free(res);
    free(b.buf);
    free(b.meta);
    return done;
}

/*
 * The reference decoder on a given chain. `words`/`prevs` hold the offloaded
 * blocks newest-first (as alignment_kernel stores them), `final_word` is
 * alignment_result_t.backtrace.backtrace. Returns a malloc'ed string.
 */
char *ref_recover_cigar(const char *pattern, int plen, const char *text, int tlen,
                        int distance, uint32_t final_word, int num_blocks,
                        const uint32_t *words)
{
    /* the decoder writes a sentinel at pattern[plen]: work on padded copies */
    char *p = (char *)calloc((size_t)plen + 16, 1);
    char *t = (char *)calloc((size_t)tlen + 16, 1);
    memcpy(p, pattern, (size_t)plen);
    memcpy(t, text, (size_t)tlen);
    wfa_backtrace_t *arr = (wfa_backtrace_t *)calloc((size_t)num_blocks + 1, sizeof(wfa_backtrace_t));
    for (int i = 0; i < num_blocks; i++) { arr[i].backtrace = words[i]; arr[i].prev = 0; }
    alignment_result_t r;
    memset(&r, 0, sizeof(r));
    r.finished = true;
    r.distance = distance;
    r.backtrace.backtrace = final_word;
    r.num_bt_blocks = num_blocks;
    wfa_cigar_t c;
    c.buffer = (char *)calloc(64, 1);
    c.buffer_size = 64;
    c.last_free_position = 0;
    recover_cigar_affine(t, p, (size_t)tlen, (size_t)plen, r.backtrace, arr, r, &c);
    free(p); free(t); free(arr);
    return c.buffer;
}

void ref_free(void *p) { free(p); }
int ref_max_threads(void) { return omp_get_max_threads(); }
