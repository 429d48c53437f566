/*
 * results.c -- result arrays and CIGAR text buffers of the public API
 * (replaces lib/alignment_results.c:24-54 and insert_ops, utils/cigar.c:31-61).
 */
#include <string.h>
#include "wfagpu_b200.h"

bool initialize_wfa_results(wfa_alignment_result_t **results, const size_t num_alignments,
                            const size_t cigar_length)
{
    if (!results) return false;
    wfa_alignment_result_t *r = (wfa_alignment_result_t *)calloc(num_alignments ? num_alignments : 1, sizeof(*r));
    if (!r) return false;
    *results = r;
    const size_t len = cigar_length ? cigar_length : 1;
    for (size_t i = 0; i < num_alignments; ++i) {
        r[i].cigar.buffer = (char *)calloc(len, 1);
        if (!r[i].cigar.buffer) return false;
        r[i].cigar.buffer_size = len;
        r[i].cigar.last_free_position = 0;
    }
    return true;
}

bool destroy_wfa_results(wfa_alignment_result_t *results, const size_t num_alignments)
{
    if (!results) return false;
    for (size_t i = 0; i < num_alignments; ++i) free(results[i].cigar.buffer);
    free(results);
    return true;
}

/* Make room for `extra` more characters plus the terminator (grows by 1.5x like
 * the reference, but keeps doing so until the text fits). */
static bool cigar_reserve(wfa_cigar_t *c, size_t extra)
{
    if (c->buffer && c->buffer_size - c->last_free_position > extra) return true;
    size_t nsz = c->buffer_size ? c->buffer_size : 16;
    while (nsz - c->last_free_position <= extra) nsz = nsz + nsz / 2 + 8;
    char *nb = (char *)realloc(c->buffer, nsz);
    if (!nb) return false;
    memset(nb + c->buffer_size, 0, nsz - c->buffer_size);
    c->buffer = nb;
    c->buffer_size = nsz;
    return true;
}

bool insert_ops(wfa_cigar_t *const cigar, const char op, const unsigned int rep)
{
    if (rep == 0) return true;
    char digits[12];
    int nd = 0;
    unsigned int v = rep;
    while (v) { digits[nd++] = (char)('0' + v % 10); v /= 10; }
    if (!cigar_reserve(cigar, (size_t)nd + 1)) return false;
    char *w = cigar->buffer + cigar->last_free_position;
    for (int i = nd - 1; i >= 0; --i) *w++ = digits[i];
    *w++ = op;
    *w = 0;
    cigar->last_free_position += (size_t)nd + 1;
    return true;
}

/* Appends `len` characters of already formatted CIGAR text (printed on the GPU). */
bool wfagpu_cigar_append(wfa_cigar_t *cigar, const char *text, size_t len)
{
    if (len == 0) return true;
    if (!cigar_reserve(cigar, len)) return false;
    memcpy(cigar->buffer + cigar->last_free_position, text, len);
    cigar->last_free_position += len;
    cigar->buffer[cigar->last_free_position] = 0;
    return true;
}
