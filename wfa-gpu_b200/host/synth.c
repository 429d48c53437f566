/*
 * synth.c -- deterministic synthetic read pairs for benchmarks and tests.
 *
 * Error model of the reference's dataset tool
 * (external/WFA/tools/generate_dataset/generate_dataset.c:144-214): text is
 * uniform over ACGT; the pattern is a copy with ceil(L*err) edits, each
 * uniformly a mismatch (to a different base), a 1-base deletion or a 1-base
 * insertion at a uniform position.  Unlike that tool (seeded with time(0)) the
 * stream is a seeded splitmix64, so every run and every rank sees the same data.
 */
#include <math.h>
#include <string.h>
#include "wfagpu_b200.h"

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static inline uint32_t rnd_below(uint64_t *s, uint32_t n)
{
    return (uint32_t)(((splitmix64(s) >> 32) * (uint64_t)n) >> 32);
}

/* pair i of stream `seed` into text / pattern (each with room for cap bytes); returns the pattern length */
static int synth_pair(uint64_t seed, size_t i, int length, double err_lo, double err_hi, char *text, char *pattern, size_t cap)
{
    static const char alphabet[4] = {'A', 'C', 'G', 'T'};
    /* one independent stream per pair: pair i is the same whatever n is */
    uint64_t s = seed ^ (0xd1342543de82ef95ull * (uint64_t)(i + 1));
    uint64_t r = splitmix64(&s);
    for (int j = 0; j < length; ++j) {
        if ((j & 31) == 0) r = splitmix64(&s);
        text[j] = alphabet[r & 3];
        r >>= 2;
    }
    text[length] = 0;
    memcpy(pattern, text, (size_t)length + 1);
    int plen = length;
    const double u = (double)(splitmix64(&s) >> 11) * (1.0 / 9007199254740992.0);
    const double err = err_lo + (err_hi - err_lo) * u;
    const int nerr = (int)ceil((double)length * err - 1e-9);
    for (int k = 0; k < nerr; ++k) {
        const uint32_t type = rnd_below(&s, 3);
        if (type == 0 && plen > 0) {
            const uint32_t pos = rnd_below(&s, (uint32_t)plen);
            char c;
            do { c = alphabet[rnd_below(&s, 4)]; } while (c == pattern[pos]);
            pattern[pos] = c;
        } else if (type == 1 && plen > 0) {
            const uint32_t pos = rnd_below(&s, (uint32_t)plen);
            memmove(pattern + pos, pattern + pos + 1, (size_t)plen - pos);
            --plen;
        } else if ((size_t)plen + 2 < cap) {
            const uint32_t pos = plen > 0 ? rnd_below(&s, (uint32_t)plen) : 0;
            memmove(pattern + pos + 1, pattern + pos, (size_t)plen - pos + 1);
            pattern[pos] = alphabet[rnd_below(&s, 4)];
            ++plen;
        }
    }
    pattern[plen] = 0;
    return plen;
}

bool wfagpu_synth_add_pairs(wfagpu_aligner_t *aligner, uint64_t seed, size_t n, int length,
                            double err_lo, double err_hi)
{
    if (!aligner || length < 0 || (size_t)length * 2 >= WFAGPU_MAX_SEQ_LEN) return false;
    const size_t cap = (size_t)length * 2 + 64;
    /* generated block by block with all host threads (a 100 000 x 10 kbp workload is 2 GB of bases and
     * 5 * 10^7 edits), appended in order: the data depends on (seed, i) only */
    enum { BLOCK = 256 };
    char *text = (char *)malloc(cap * BLOCK);
    char *pattern = (char *)malloc(cap * BLOCK);
    if (!text || !pattern) { free(text); free(pattern); return false; }
    bool ok = wfagpu_reserve(aligner, n * (2 * (size_t)length + (size_t)((double)length * err_hi) + 24), n);
    for (size_t i0 = 0; i0 < n && ok; i0 += BLOCK) {
        const long cnt = (long)(n - i0 < BLOCK ? n - i0 : BLOCK);
        #pragma omp parallel for schedule(dynamic, 4) if (cnt > 8 && length >= 512)
        for (long b = 0; b < cnt; ++b)
            synth_pair(seed, i0 + (size_t)b, length, err_lo, err_hi, text + (size_t)b * cap, pattern + (size_t)b * cap, cap);
        for (long b = 0; b < cnt && ok; ++b)
            ok = wfagpu_add_sequences(aligner, pattern + (size_t)b * cap, text + (size_t)b * cap);
    }
    free(text);
    free(pattern);
    return ok;
}
