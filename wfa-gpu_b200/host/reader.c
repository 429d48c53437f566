/*
 * reader.c -- input formats of the CLI (replaces utils/sequence_reader.c:137-392).
 *
 *   .seq   : one pair per two lines, ">PATTERN" then "<TEXT"; blank lines are skipped.
 *   FASTA  : two files, the n-th record of the query file pairs with the n-th record of
 *            the target file; sequences may span several lines; header lines start with
 *            '>' (leading blanks allowed).
 *
 * Both fill a wfagpu_aligner_t through wfagpu_add_sequences, i.e. the same host layout
 * (4-byte aligned, NUL padded) the reference readers produce.  Unlike the reference, a
 * final line without a newline keeps its last base and '\r' is dropped.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <string.h>
#include "wfagpu_b200.h"

static size_t chomp(char *line, ssize_t n)
{
    while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
    return (size_t)n;
}

long wfagpu_read_seq_file(wfagpu_aligner_t *aligner, const char *path, size_t max_pairs)
{
    FILE *fp = fopen(path, "r");
    if (!fp) { fprintf(stderr, "[!] ERROR: Could not open %s\n", path); return -1; }
    char *line = NULL, *pattern = NULL;
    size_t cap = 0, lineno = 0;
    ssize_t n;
    long pairs = 0;
    bool ok = true;
    while (ok && (max_pairs == 0 || (size_t)pairs < max_pairs) && (n = getline(&line, &cap, fp)) != -1) {
        ++lineno;
        const size_t len = chomp(line, n);
        if (len == 0) continue;
        if (!pattern) {
            if (line[0] != '>') { fprintf(stderr, "[!] ERROR: Invalid file format. Could not read pattern in line %zu\n", lineno); ok = false; break; }
            pattern = strdup(line + 1);
        } else {
            if (line[0] != '<') { fprintf(stderr, "[!] ERROR: Invalid file format. Could not read text in line %zu\n", lineno); ok = false; break; }
            ok = wfagpu_add_sequences(aligner, pattern, line + 1);
            free(pattern);
            pattern = NULL;
            ++pairs;
        }
    }
    free(pattern);
    free(line);
    fclose(fp);
    return ok ? pairs : -1;
}

typedef struct { FILE *fp; char *line; size_t cap; bool eof; bool pending_header; } fasta_t;

static bool is_header(const char *s) { while (*s == ' ') ++s; return *s == '>'; }

/* Reads the next record's sequence into *seq (realloc'ed). Returns false when no record is left. */
static bool fasta_next(fasta_t *f, char **seq, size_t *seq_cap)
{
    if (f->eof) return false;
    size_t len = 0;
    bool any = false;
    ssize_t n;
    if (*seq_cap == 0) { *seq_cap = 1 << 16; *seq = (char *)malloc(*seq_cap); }
    (*seq)[0] = 0;
    while ((n = getline(&f->line, &f->cap, f->fp)) != -1) {
        const size_t l = chomp(f->line, n);
        if (l == 0) continue;
        if (is_header(f->line)) {
            if (!f->pending_header) { f->pending_header = true; continue; }  /* very first header */
            return true;                                                       /* next record starts */
        }
        f->pending_header = true;
        if (len + l + 1 > *seq_cap) { while (len + l + 1 > *seq_cap) *seq_cap *= 2; *seq = (char *)realloc(*seq, *seq_cap); }
        memcpy(*seq + len, f->line, l + 1);
        len += l;
        any = true;
    }
    f->eof = true;
    return any;
}

long wfagpu_read_fasta_files(wfagpu_aligner_t *aligner, const char *query_path, const char *target_path, size_t max_pairs)
{
    fasta_t q = {0}, t = {0};
    q.fp = fopen(query_path, "r");
    t.fp = fopen(target_path, "r");
    if (!q.fp || !t.fp) {
        fprintf(stderr, "[!] ERROR: Could not open %s\n", !q.fp ? query_path : target_path);
        if (q.fp) fclose(q.fp);
        if (t.fp) fclose(t.fp);
        return -1;
    }
    char *qs = NULL, *ts = NULL;
    size_t qc = 0, tc = 0;
    long pairs = 0;
    bool ok = true;
    while (ok && (max_pairs == 0 || (size_t)pairs < max_pairs)) {
        const bool hq = fasta_next(&q, &qs, &qc);
        const bool ht = fasta_next(&t, &ts, &tc);
        if (!hq || !ht) break;
        ok = wfagpu_add_sequences(aligner, qs, ts);
        ++pairs;
    }
    free(qs); free(ts); free(q.line); free(t.line);
    fclose(q.fp); fclose(t.fp);
    if (pairs == 0 && ok) { fprintf(stderr, "[!] ERROR: Empty FASTA file.\n"); return -1; }
    return ok ? pairs : -1;
}

/* -c: validates a result without any CPU aligner (replaces check_cigar_edit +
 * check_affine_distance, utils/verification.c:27-146): the CIGAR must transform the pattern
 * into the text (I consumes text, D consumes pattern) and its gap-affine cost must be `error`. */
bool wfagpu_check_result(const char *pattern, size_t plen, const char *text, size_t tlen,
                         affine_penalties_t pen, unsigned int error, const char *cigar)
{
    size_t v = 0, h = 0;
    unsigned long score = 0;
    const char *c = cigar;
    if (!c) return false;
    while (*c) {
        unsigned long rep = 0;
        if (*c < '0' || *c > '9') return false;
        while (*c >= '0' && *c <= '9') rep = rep * 10 + (unsigned long)(*c++ - '0');
        const char op = *c++;
        if (op == 'M' || op == 'X') {
            if (v + rep > plen || h + rep > tlen) return false;
            for (unsigned long i = 0; i < rep; ++i)
                if ((pattern[v + i] == text[h + i]) != (op == 'M')) return false;
            v += rep; h += rep;
            if (op == 'X') score += rep * (unsigned long)pen.x;
        } else if (op == 'I') { h += rep; score += (unsigned long)pen.o + rep * (unsigned long)pen.e; }
        else if (op == 'D') { v += rep; score += (unsigned long)pen.o + rep * (unsigned long)pen.e; }
        else return false;
        if (v > plen || h > tlen) return false;
    }
    return v == plen && h == tlen && score == error;
}


/* The reference's two generic validators under their own names and argument order (text first):
 * utils/verification.h:37-49.  Exported by the reference library and used by its `-c` path. */
static bool walk_cigar(const char *text, const char *pattern, size_t tlen, size_t plen, const char *cigar,
                       const affine_penalties_t *pen, unsigned long *score_out)
{
    size_t v = 0, h = 0;
    unsigned long score = 0;
    const char *c = cigar;
    if (!c || !text || !pattern) return false;
    while (*c) {
        unsigned long rep = 0;
        if (*c < '0' || *c > '9') return false;
        while (*c >= '0' && *c <= '9') rep = rep * 10 + (unsigned long)(*c++ - '0');
        const char op = *c++;
        if (op == 'M' || op == 'X') {
            if (v + rep > plen || h + rep > tlen) return false;
            for (unsigned long i = 0; i < rep; ++i)
                if ((pattern[v + i] == text[h + i]) != (op == 'M')) return false;
            v += rep; h += rep;
            if (op == 'X' && pen) score += rep * (unsigned long)pen->x;
        } else if (op == 'I') { h += rep; if (pen) score += (unsigned long)pen->o + rep * (unsigned long)pen->e; }
        else if (op == 'D') { v += rep; if (pen) score += (unsigned long)pen->o + rep * (unsigned long)pen->e; }
        else return false;
        if (v > plen || h > tlen) return false;
    }
    if (score_out) *score_out = score;
    return v == plen && h == tlen;
}

bool check_cigar_edit(const char *text, const char *pattern, const int tlen, const int plen, const char *curr_cigar)
{
    if (tlen < 0 || plen < 0) return false;
    return walk_cigar(text, pattern, (size_t)tlen, (size_t)plen, curr_cigar, NULL, NULL);
}

bool check_affine_distance(const char *text, const char *pattern, const int tlen, const int plen, const int distance,
                           const affine_penalties_t penalties, const char *cigar)
{
    unsigned long score = 0;
    if (tlen < 0 || plen < 0 || distance < 0) return false;
    if (!walk_cigar(text, pattern, (size_t)tlen, (size_t)plen, cigar, &penalties, &score)) return false;
    return score == (unsigned long)distance;
}
