/*
 * kernel_model.c -- sequential CPU model of the B200 kernel's *algorithm*
 * (wfa-gpu_b200/csrc/wfa_kernels.cu), NOT of the reference.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as wfagpu_oracle.c).  It exists so the
 * re-designed data flow can be proven equal to the faithful restatement on the
 * CPU, where there is no GPU:
 *
 *   - penalty-only step table (kind / half-width / decision-row offset),
 *   - symmetric range [-n, n] with NULL guard cells instead of per-pair
 *     re-initialisation, rings of depth A (M) and e+1 (I, D),
 *   - "clean" source semantics (a source that does not exist reads as NULL;
 *     the reference reads stale ring contents instead, which can never lie on
 *     an optimal path -- see DESIGN.md "Why clean rings are exact"),
 *   - one decision byte per cell (I ext, D ext, M winner) instead of piggy-backed 32-bit
 *     words + prev pointers, followed by a geometric traceback that emits the
 *     same 2-bit op stream the reference's chain decodes to.
 *
 * tests/test_kernel_model.py checks (finished, distance, CIGAR) of this model
 * against wfagpu_oracle.c over random pairs and penalty sets.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KM_NULL (-32000)
#define KM_KIND_NULL 0
#define KM_KIND_M 1
#define KM_KIND_MDI 2

typedef struct {
    int32_t n;        /* half width of the range after this step            */
    uint32_t row_off; /* decision row offset, in bytes, MDI steps only      */
    uint8_t kind;
} km_step_t;

/* Host-side table construction; mirrors wfagpu_build_step_table() in
 * wfa-gpu_b200/host/step_table.c.  Existence logic follows
 * lib/kernels/sequence_alignment_kernel.cu:584-631 (it depends on the
 * penalties only).  Returns d_end: scores 1 .. d_end-1 may be computed. */
int km_build_steps(int x, int o, int e, int max_steps, int max_dist, km_step_t *tab, uint64_t *arena_bytes)
{
    uint8_t *exM = (uint8_t *)calloc((size_t)max_dist + 1, 1);
    uint8_t *exI = (uint8_t *)calloc((size_t)max_dist + 1, 1);
    int steps = 1, n = 0, mdi = 0, d;
    uint64_t off = 0;
    exM[0] = 1;
    tab[0].kind = KM_KIND_M; tab[0].n = 0; tab[0].row_off = 0;
    for (d = 1; d < max_dist; d++) {
        if (!(steps < max_steps - 1)) break;
        int gap = 0, mx = 0;
        if (d - o - e >= 0) gap = exM[d - o - e] || exI[d - e];
        if (gap) mx = 1;
        else if (d - x >= 0) mx = exM[d - x];
        if (!gap && !mx) {
            tab[d].kind = KM_KIND_NULL;
        } else if (!gap) {
            tab[d].kind = KM_KIND_M; exM[d] = 1;
        } else {
            tab[d].kind = KM_KIND_MDI; exM[d] = 1; exI[d] = 1;
            mdi++; steps++;
            /* reachable diagonals: a gap of |k| bases costs at least o + |k| e */
            { int reach = (d - o) / e; n = mdi < reach ? mdi : reach; if (n < 0) n = 0; }
        }
        tab[d].n = n;
        tab[d].row_off = (uint32_t)off;
        if (tab[d].kind == KM_KIND_MDI) off += 16u * (uint64_t)((2 * n + 1 + 15) / 16);
    }
    free(exM); free(exI);
    if (arena_bytes) *arena_bytes = off;
    return d;
}

static inline int km_code(char c) { return (c & 6) >> 1; }

static int km_extend(const char *text, const char *pattern, int tlen, int plen, int k, int off)
{
    int v = off - k, h = off;
    if (v > plen || h > tlen) return KM_NULL;
    while (v < plen && h < tlen && km_code(pattern[v]) == km_code(text[h])) { v++; h++; off++; }
    return off;
}

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/*
 * One pair.  ops_out receives the 2-bit ops NEWEST FIRST (traceback order),
 * one per byte; *n_ops their count.  Returns 0, or -1 on capacity problems.
 */
int km_align_pair(const char *pattern, int plen, const char *text, int tlen,
                  int x, int o, int e, const km_step_t *tab, int d_end, int n_cap,
                  int with_bt, int *finished, int *distance,
                  uint8_t *ops_out, int ops_cap, int *n_ops, long *cells_out)
{
    const int A = imax(o + e, x) + 1;
    const int E1 = e + 1;
    const int G = A;
    const int C = n_cap + 2 * G + 2;         /* centre index */
    const int W = 2 * C + 1;
    int16_t *Mr = (int16_t *)malloc((size_t)A * W * sizeof(int16_t));
    int16_t *Ir = (int16_t *)malloc((size_t)E1 * W * sizeof(int16_t));
    int16_t *Dr = (int16_t *)malloc((size_t)E1 * W * sizeof(int16_t));
    uint64_t arena_bytes = 0;
    for (int d = 0; d < d_end; d++)
        if (tab[d].kind == KM_KIND_MDI) arena_bytes = tab[d].row_off + 16u * (uint64_t)((2 * tab[d].n + 1 + 15) / 16);
    uint8_t *arena = with_bt ? (uint8_t *)calloc(arena_bytes + 16, 1) : NULL;
    /* poison the rings: the kernel never clears them between pairs */
    for (long i = 0; i < (long)A * W; i++) Mr[i] = 12345;
    for (long i = 0; i < (long)E1 * W; i++) { Ir[i] = 12345; Dr[i] = 12345; }
    long cells = 0;

    /* pair prologue: NULL-fill every ring row over [-2G, 2G] */
    for (int r = 0; r < A; r++) for (int k = -2 * G; k <= 2 * G; k++) Mr[r * W + C + k] = KM_NULL;
    for (int r = 0; r < E1; r++) for (int k = -2 * G; k <= 2 * G; k++) { Ir[r * W + C + k] = KM_NULL; Dr[r * W + C + k] = KM_NULL; }
    Mr[0 * W + C + 0] = (int16_t)km_extend(text, pattern, tlen, plen, 0, 0);

    const int kt = tlen - plen;
    int fin = 0, d = 0;
    if (kt == 0 && Mr[C] == tlen) {
        fin = 1;
    } else {
        for (d = 1; d < d_end; d++) {
            const km_step_t st = tab[d];
            const int n = st.n;
            if (n > n_cap) break;                       /* ring capacity exceeded */
            int16_t *Mc = Mr + (d % A) * W + C;
            int16_t *Ic = Ir + (d % E1) * W + C;
            int16_t *Dc = Dr + (d % E1) * W + C;
            if (st.kind == KM_KIND_NULL) {
                for (int k = -n - G; k <= n + G; k++) { Mc[k] = KM_NULL; Ic[k] = KM_NULL; Dc[k] = KM_NULL; }
                continue;
            }
            const int16_t *Mx = (d - x >= 0) ? Mr + ((d - x) % A) * W + C : NULL;
            if (st.kind == KM_KIND_M) {
                for (int k = -n - G; k <= n + G; k++) {
                    Ic[k] = KM_NULL; Dc[k] = KM_NULL;
                    if (k < -n || k > n) { Mc[k] = KM_NULL; continue; }
                    int m = Mx[k] + 1;
                    if (m >= 0) m = km_extend(text, pattern, tlen, plen, k, m);
                    Mc[k] = (int16_t)m;
                    cells++;
                }
            } else {
                /* sources that do not exist read as NULL rows: by construction their
                 * ring rows hold NULL over the whole range that can be touched. */
                const int16_t *Mo = Mr + ((((d - o - e) % A) + A) % A) * W + C;
                const int16_t *Ie = Ir + ((((d - e) % E1) + E1) % E1) * W + C;
                const int16_t *De = Dr + ((((d - e) % E1) + E1) % E1) * W + C;
                const int16_t *Mxx = Mr + ((((d - x) % A) + A) % A) * W + C;
                uint8_t *row = with_bt ? arena + st.row_off : NULL;
                for (int k = -n - G; k < -n; k++) { Mc[k] = KM_NULL; Ic[k] = KM_NULL; Dc[k] = KM_NULL; }
                for (int k = n + 1; k <= n + G; k++) { Mc[k] = KM_NULL; Ic[k] = KM_NULL; Dc[k] = KM_NULL; }
                for (int k = -n; k <= n; k++) {
                    const int io = Mo[k - 1] + 1, ie = Ie[k - 1] + 1;
                    const int pI = imax(io * 2, ie * 2 + 1);
                    const int I = pI >> 1;
                    const int dopen = Mo[k + 1], dext = De[k + 1];
                    const int pD = imax(dopen * 2, dext * 2 + 1);
                    const int D = pD >> 1;
                    const int X = Mxx[k] + 1;
                    const int pM = imax(imax(X * 4 + 2, D * 4 + 3), I * 4 + 1);
                    int M = pM >> 2;
                    const int mop = pM & 3;
                    if (M >= 0) M = km_extend(text, pattern, tlen, plen, k, M);
                    Ic[k] = (int16_t)I; Dc[k] = (int16_t)D; Mc[k] = (int16_t)M;
                    if (with_bt) row[k + n] = (uint8_t)((pI & 1) | ((pD & 1) << 1) | (mop << 2));
                    cells++;
                }
            }
            if (kt >= -n && kt <= n && Mc[kt] == tlen) { fin = 1; break; }
        }
    }
    *finished = fin;
    *distance = fin ? d : 0;
    if (cells_out) *cells_out = cells;
    *n_ops = 0;

    int rc = 0;
    if (fin && with_bt) {
        /* traceback: emits SUB for every M cell, INS/DEL for every I/D cell */
        int cd = d, ck = kt, comp = 0, cnt = 0; /* comp: 0 M, 1 I, 2 D */
        while (!(comp == 0 && cd == 0)) {
            if (cd < 0 || cnt >= ops_cap) { rc = -1; break; }
            const km_step_t st = tab[cd];
            if (comp == 0) {
                ops_out[cnt++] = 2;
                if (st.kind == KM_KIND_M) { cd -= x; continue; }
                const int mop = (arena[st.row_off + ck + st.n] >> 2) & 3;
                if (mop == 2) cd -= x;
                else if (mop == 1) comp = 1;
                else comp = 2;
            } else {
                const uint8_t dec = arena[st.row_off + ck + st.n];
                if (comp == 1) {
                    ops_out[cnt++] = 1;
                    const int ext = dec & 1;
                    ck -= 1;
                    if (ext) cd -= e; else { cd -= o + e; comp = 0; }
                } else {
                    ops_out[cnt++] = 3;
                    const int ext = (dec >> 1) & 1;
                    ck += 1;
                    if (ext) cd -= e; else { cd -= o + e; comp = 0; }
                }
            }
        }
        *n_ops = cnt;
    }
    free(Mr); free(Ir); free(Dr); free(arena);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Checkpointed traceback (model of wfa_exact_kernel's CKPT mode).
 *
 * Instead of one decision byte per cell, the forward pass copies the ring rows (the last A
 * scores of M, the last e+1 of I and D) every P scores.  The traceback then re-derives the
 * offsets it needs: from the top cell (d_e, k_e) of a segment it loads the checkpoint c < d_e,
 * recomputes scores c+1 .. d_e on the shrinking cone |k - k_e| <= d_e - d (every cell of the
 * cone only depends on cells of the cone one level down, e >= 1, x >= 1), and walks the path
 * inside the segment by comparing candidate offsets with the forward tie-breaks.  Work:
 * ~ score * P cells per pair instead of a store per cell in the forward pass.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int A, E1, W, C, P;
    const km_step_t *tab;
    int x, o, e;
    /* checkpoint j (score j*P): rows by age a: M[a] = M of score j*P - a */
    int16_t **ckM, **ckI, **ckD;
    /* segment scratch: rows by level i (score c + i), index k - klo */
    int16_t *sM, *sI, *sD;
    int c, klo, khi, R;
    const char *pattern, *text;
    int plen, tlen;
    int16_t m00; /* M[0][0] */
    int kt, dmax; /* score-bound pruning (dmax < 0: off) */
} km_tb_t;

/* Score-bound pruning: when the pair is known to finish with a score <= dmax, a cell (d, k) can
 * only lie on an alignment of that score if d + e * |k - kt| <= dmax (every diagonal between k and
 * the target diagonal kt costs at least one gap extension).  Cells outside are never computed and
 * read as NULL.  Every complete path through a pruned cell scores more than dmax, so the cells of
 * the optimal path the reference backtraces keep their values and win the same tie-breaks:
 * identical score and CIGAR for every pair that finishes within dmax. */
static void km_prune_range(int d, int n, int kt, int dmax, int e, int *lo, int *hi)
{
    *lo = -n; *hi = n;
    if (dmax < 0) return;
    if (d > dmax) { *lo = 1; *hi = 0; return; }
    const int g = (dmax - d) / e;
    if (kt - g > *lo) *lo = kt - g;
    if (kt + g < *hi) *hi = kt + g;
}

static int km_in_range(const km_tb_t *t, int d, int k)
{
    if (d < 0) return 0;
    int lo, hi;
    km_prune_range(d, t->tab[d].n, t->kt, t->dmax, t->e, &lo, &hi);
    return k >= lo && k <= hi;
}

/* offset of component comp (0 M, 1 I, 2 D) at (d, k) as the forward pass saw it */
static int km_tb_get(const km_tb_t *t, int comp, int d, int k)
{
    if (d > t->c && !km_in_range(t, d, k)) return KM_NULL;
    if (d <= t->c && t->c > 0) {
        /* snapshot rows are masked with ONE window, that of score c widened by a cell (what the
         * GPU traceback does); inside it older rows hold values, guard NULLs -- never stale cells */
        int lo, hi;
        const int nc = t->tab[t->c].n;
        km_prune_range(t->c, nc, t->kt, t->dmax, t->e, &lo, &hi);
        if (t->dmax >= 0) { lo = imax(-nc, lo - 1); hi = imin(nc, hi + 1); }
        if (k < lo || k > hi) return KM_NULL;
    }
    if (d > t->c) {
        const int i = d - t->c;
        if (k < t->klo || k > t->khi) return KM_NULL; /* outside the cone: never needed */
        const int16_t *row = (comp == 0 ? t->sM : comp == 1 ? t->sI : t->sD) + (size_t)i * (t->khi - t->klo + 1);
        return row[k - t->klo];
    }
    if (t->c == 0) {
        /* checkpoint 0 is the initial state */
        if (d == 0 && comp == 0 && k == 0) return t->m00;
        return KM_NULL;
    }
    const int a = t->c - d, j = t->c / t->P;
    if (comp == 0) { if (a >= t->A) return KM_NULL; return t->ckM[j][(size_t)a * t->W + t->C + k]; }
    if (a >= t->E1) return KM_NULL;
    return (comp == 1 ? t->ckI : t->ckD)[j][(size_t)a * t->W + t->C + k];
}

int km_align_pair_ckpt(const char *pattern, int plen, const char *text, int tlen,
                       int x, int o, int e, const km_step_t *tab, int d_end, int n_cap, int period, int dmax,
                       int *finished, int *distance, uint8_t *ops_out, int ops_cap, int *n_ops, long *recomputed)
{
    const int A = imax(o + e, x) + 1, E1 = e + 1, G = A;
    const int C = n_cap + 2 * G + 2, W = 2 * C + 1, P = period;
    int16_t *Mr = (int16_t *)malloc((size_t)A * W * sizeof(int16_t));
    int16_t *Ir = (int16_t *)malloc((size_t)E1 * W * sizeof(int16_t));
    int16_t *Dr = (int16_t *)malloc((size_t)E1 * W * sizeof(int16_t));
    const int n_ck = d_end / P + 2;
    int16_t **ckM = (int16_t **)calloc((size_t)n_ck, sizeof(*ckM));
    int16_t **ckI = (int16_t **)calloc((size_t)n_ck, sizeof(*ckI));
    int16_t **ckD = (int16_t **)calloc((size_t)n_ck, sizeof(*ckD));
    for (long i = 0; i < (long)A * W; i++) Mr[i] = 12345;
    for (long i = 0; i < (long)E1 * W; i++) { Ir[i] = 12345; Dr[i] = 12345; }
    for (int r = 0; r < A; r++) for (int k = -2 * G; k <= 2 * G; k++) Mr[r * W + C + k] = KM_NULL;
    for (int r = 0; r < E1; r++) for (int k = -2 * G; k <= 2 * G; k++) { Ir[r * W + C + k] = KM_NULL; Dr[r * W + C + k] = KM_NULL; }
    Mr[C] = (int16_t)km_extend(text, pattern, tlen, plen, 0, 0);
    const int16_t m00 = Mr[C];
    const int kt = tlen - plen;
    int fin = 0, d = 0;
    if (kt == 0 && Mr[C] == tlen) {
        fin = 1;
    } else {
        for (d = 1; d < d_end; d++) {
            const km_step_t st = tab[d];
            const int n = st.n;
            if (n > n_cap) break;
            int16_t *Mc = Mr + (d % A) * W + C, *Ic = Ir + (d % E1) * W + C, *Dc = Dr + (d % E1) * W + C;
            if (st.kind == KM_KIND_NULL) {
                for (int k = -n - G; k <= n + G; k++) { Mc[k] = KM_NULL; Ic[k] = KM_NULL; Dc[k] = KM_NULL; }
            } else if (st.kind == KM_KIND_M) {
                const int16_t *Mx = Mr + ((((d - x) % A) + A) % A) * W + C;
                int lo, hi;
                km_prune_range(d, n, kt, dmax, e, &lo, &hi);
                if (dmax >= 0) for (int k = -n - G; k <= n + G; k++) { Mc[k] = 12345; Ic[k] = 12345; Dc[k] = 12345; } /* poison */
                for (int k = lo - G; k <= hi + G; k++) {
                    Ic[k] = KM_NULL; Dc[k] = KM_NULL;
                    if (k < lo || k > hi) { Mc[k] = KM_NULL; continue; }
                    int m = Mx[k] + 1;
                    if (m >= 0) m = km_extend(text, pattern, tlen, plen, k, m);
                    Mc[k] = (int16_t)m;
                }
            } else {
                const int16_t *Mo = Mr + ((((d - o - e) % A) + A) % A) * W + C;
                const int16_t *Ie = Ir + ((((d - e) % E1) + E1) % E1) * W + C;
                const int16_t *De = Dr + ((((d - e) % E1) + E1) % E1) * W + C;
                const int16_t *Mxx = Mr + ((((d - x) % A) + A) % A) * W + C;
                int lo, hi;
                km_prune_range(d, n, kt, dmax, e, &lo, &hi);
                if (dmax >= 0) for (int k = -n - G; k <= n + G; k++) { Mc[k] = 12345; Ic[k] = 12345; Dc[k] = 12345; } /* poison */
                for (int k = lo - G; k < lo; k++) { Mc[k] = KM_NULL; Ic[k] = KM_NULL; Dc[k] = KM_NULL; }
                for (int k = hi + 1; k <= hi + G; k++) { Mc[k] = KM_NULL; Ic[k] = KM_NULL; Dc[k] = KM_NULL; }
                for (int k = lo; k <= hi; k++) {
                    const int I = imax(Mo[k - 1] + 1, Ie[k - 1] + 1);
                    const int D = imax(Mo[k + 1], De[k + 1]);
                    int M = imax(imax(Mxx[k] + 1, D), I);
                    if (M >= 0) M = km_extend(text, pattern, tlen, plen, k, M);
                    Ic[k] = (int16_t)I; Dc[k] = (int16_t)D; Mc[k] = (int16_t)M;
                }
            }
            const int done = (st.kind != KM_KIND_NULL) && kt >= -n && kt <= n && Mc[kt] == tlen;
            if (!done && d % P == 0) {
                /* checkpoint: ring rows by age */
                const int j = d / P;
                ckM[j] = (int16_t *)malloc((size_t)A * W * sizeof(int16_t));
                ckI[j] = (int16_t *)malloc((size_t)E1 * W * sizeof(int16_t));
                ckD[j] = (int16_t *)malloc((size_t)E1 * W * sizeof(int16_t));
                for (int a = 0; a < A; a++)
                    memcpy(ckM[j] + (size_t)a * W, Mr + ((((d - a) % A) + A) % A) * W, (size_t)W * sizeof(int16_t));
                for (int a = 0; a < E1; a++) {
                    memcpy(ckI[j] + (size_t)a * W, Ir + ((((d - a) % E1) + E1) % E1) * W, (size_t)W * sizeof(int16_t));
                    memcpy(ckD[j] + (size_t)a * W, Dr + ((((d - a) % E1) + E1) % E1) * W, (size_t)W * sizeof(int16_t));
                }
            }
            if (done) { fin = 1; break; }
        }
    }
    *finished = fin;
    *distance = fin ? d : 0;
    *n_ops = 0;
    long rec = 0;
    int rc = 0;
    if (fin && d > 0) {
        km_tb_t t;
        memset(&t, 0, sizeof(t));
        t.A = A; t.E1 = E1; t.W = W; t.C = C; t.P = P; t.tab = tab; t.x = x; t.o = o; t.e = e;
        t.ckM = ckM; t.ckI = ckI; t.ckD = ckD; t.pattern = pattern; t.text = text; t.plen = plen; t.tlen = tlen;
        t.m00 = m00; t.kt = kt; t.dmax = dmax;
        const int SW = 2 * P + 3;
        t.sM = (int16_t *)malloc((size_t)(P + 2) * SW * sizeof(int16_t));
        t.sI = (int16_t *)malloc((size_t)(P + 2) * SW * sizeof(int16_t));
        t.sD = (int16_t *)malloc((size_t)(P + 2) * SW * sizeof(int16_t));
        int cd = d, ck = kt, comp = 0, cnt = 0;
        while (!(comp == 0 && cd == 0) && rc == 0) {
            if (cd <= 0) { rc = -1; break; }
            /* ---- open the segment that contains (cd, ck) ---- */
            const int c = ((cd - 1) / P) * P;
            const int R = cd - c;
            t.c = 0;  /* while filling, read sources through the same accessor: set bounds first */
            t.klo = ck - R; t.khi = ck + R; t.R = R;
            const int sw = t.khi - t.klo + 1;
            t.c = c;
            for (int i = 1; i <= R; i++) {
                const int dd = c + i;
                const km_step_t st = tab[dd];
                int16_t *rm = t.sM + (size_t)i * sw, *ri = t.sI + (size_t)i * sw, *rd = t.sD + (size_t)i * sw;
                for (int k = t.klo; k <= t.khi; k++) { rm[k - t.klo] = KM_NULL; ri[k - t.klo] = KM_NULL; rd[k - t.klo] = KM_NULL; }
                if (st.kind == KM_KIND_NULL) continue;
                const int half = R - i;
                for (int k = ck - half; k <= ck + half; k++) {
                    if (!km_in_range(&t, dd, k)) continue;
                    rec++;
                    if (st.kind == KM_KIND_M) {
                        int m = km_tb_get(&t, 0, dd - x, k) + 1;
                        if (m >= 0) m = km_extend(text, pattern, tlen, plen, k, m);
                        rm[k - t.klo] = (int16_t)m;
                    } else {
                        const int I = imax(km_tb_get(&t, 0, dd - o - e, k - 1) + 1, km_tb_get(&t, 1, dd - e, k - 1) + 1);
                        const int D = imax(km_tb_get(&t, 0, dd - o - e, k + 1), km_tb_get(&t, 2, dd - e, k + 1));
                        int M = imax(imax(km_tb_get(&t, 0, dd - x, k) + 1, D), I);
                        if (M >= 0) M = km_extend(text, pattern, tlen, plen, k, M);
                        ri[k - t.klo] = (int16_t)I; rd[k - t.klo] = (int16_t)D; rm[k - t.klo] = (int16_t)M;
                    }
                }
            }
            /* ---- walk the path while it stays above the checkpoint ---- */
            while (cd > c && !(comp == 0 && cd == 0)) {
                if (cnt >= ops_cap) { rc = -1; break; }
                const km_step_t st = tab[cd];
                if (comp == 0) {
                    ops_out[cnt++] = 2;
                    if (st.kind == KM_KIND_M) { cd -= x; continue; }
                    const int X = km_tb_get(&t, 0, cd - x, ck) + 1;
                    const int I = km_tb_get(&t, 1, cd, ck), D = km_tb_get(&t, 2, cd, ck);
                    if (D >= X && D >= I) comp = 2;          /* D beats X beats I */
                    else if (X >= I) cd -= x;
                    else comp = 1;
                } else if (comp == 1) {
                    ops_out[cnt++] = 1;
                    const int op = km_tb_get(&t, 0, cd - o - e, ck - 1) + 1, ex = km_tb_get(&t, 1, cd - e, ck - 1) + 1;
                    ck -= 1;
                    if (ex >= op) cd -= e; else { cd -= o + e; comp = 0; }   /* extend beats open */
                } else {
                    ops_out[cnt++] = 3;
                    const int op = km_tb_get(&t, 0, cd - o - e, ck + 1), ex = km_tb_get(&t, 2, cd - e, ck + 1);
                    ck += 1;
                    if (ex >= op) cd -= e; else { cd -= o + e; comp = 0; }
                }
            }
        }
        *n_ops = cnt;
        free(t.sM); free(t.sI); free(t.sD);
    }
    if (recomputed) *recomputed = rec;
    for (int j = 0; j < n_ck; j++) { free(ckM[j]); free(ckI[j]); free(ckD[j]); }
    free(ckM); free(ckI); free(ckD);
    free(Mr); free(Ir); free(Dr);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Adaptive band on packed int16 quads (model of wfa_bandq_kernel + wfa_band_traceback_kernel).
 *
 * The heuristic is the reference's (window clip hi--/lo++, re-centre every `band` scores on the
 * first diagonal with the smallest distance to the target, stale ring slots on null steps); what
 * is modelled here is the kernel's DATA LAYOUT: a ring row stores the cells of its score's window
 * at index k - base, base = lo rounded down to a multiple of four, in whole quads, cells outside
 * the window as NULL; a read outside the stored part of a row (index < 0 or >= cells) yields NULL
 * without touching memory; windows are records {lo, hi, base, cells} per slot (M rows, and one
 * shared by the I and D rows); decisions come from the "which operand won" predicates of the max
 * instructions (extend >= open, D >= X, max(D, X) >= I) and sit at byte k - base of the score's
 * row; the backtrace reads the row base of every score from a table.  The rings are poisoned so
 * that a read of a cell the kernel never wrote would show.
 * ------------------------------------------------------------------------------------------ */
typedef struct { int lo, hi, base, cells; } km_win_t;

static inline int km_bq_get(const int16_t *row, km_win_t w, int k)
{
    const int idx = k - w.base;
    return (idx >= 0 && idx < w.cells) ? row[idx] : KM_NULL;
}

int km_align_pair_bandq(const char *pattern, int plen, const char *text, int tlen,
                        int x, int o, int e, const km_step_t *tab, int d_end, int band, int W,
                        int with_bt, int *finished, int *distance,
                        uint8_t *ops_out, int ops_cap, int *n_ops)
{
    const int A = imax(o + e, x) + 1;
    const int oe = o + e;
    const int RW = ((W + 3) & ~3) + 8;
    int16_t *R[3];
    for (int c = 0; c < 3; c++) {
        R[c] = (int16_t *)malloc((size_t)A * RW * sizeof(int16_t));
        for (long i = 0; i < (long)A * RW; i++) R[c][i] = 12345;           /* poison */
    }
    km_win_t *WM = (km_win_t *)malloc((size_t)A * sizeof(km_win_t));
    km_win_t *WG = (km_win_t *)malloc((size_t)A * sizeof(km_win_t));
    uint8_t *arena = with_bt ? (uint8_t *)malloc((size_t)d_end * RW) : NULL;
    int *base_tab = (int *)malloc((size_t)d_end * sizeof(int));
    if (arena) memset(arena, 0xEE, (size_t)d_end * RW);
    /* every slot starts as the one-diagonal window [0, 0] holding NULL: one quad of NULLs at base 0 */
    for (int s = 0; s < A; s++) {
        WM[s] = (km_win_t){0, 0, 0, 4};
        WG[s] = (km_win_t){0, 0, 0, 4};
        for (int c = 0; c < 3; c++) for (int i = 0; i < 4; i++) R[c][s * RW + i] = KM_NULL;
    }
    R[0][0] = (int16_t)km_extend(text, pattern, tlen, plen, 0, 0);
    const int kt = tlen - plen, kt_abs = kt < 0 ? -kt : kt;
    int fin = 0, dist = 0, rc = 0;
    if (kt == 0 && R[0][0] == tlen) {
        fin = 1;
    } else {
        for (int d = 1; d < d_end; d++) {
            const km_step_t st = tab[d];
            const int sM = d % A;
            if (st.kind == KM_KIND_NULL) continue;
            const int sx = ((sM - x) % A + A) % A;
            const km_win_t wx = WM[sx];
            const int16_t *Mx = R[0] + sx * RW;
            int16_t *Mc = R[0] + sM * RW;
            int lo, hi, base;
            if (st.kind == KM_KIND_M) {
                lo = wx.lo; hi = wx.hi; base = wx.base;
                const int nq = ((hi - base) >> 2) + 1;
                for (int q = 0; q < nq; q++)
                    for (int i = 0; i < 4; i++) {
                        const int k = base + 4 * q + i;
                        int m = (int16_t)(Mx[k - wx.base] + 1);           /* same base: always a stored cell */
                        if (k < lo || k > hi) m = KM_NULL;
                        if (m >= 0) m = km_extend(text, pattern, tlen, plen, k, m);
                        Mc[k - base] = (int16_t)m;
                    }
                WM[sM] = wx;
            } else {
                const int so = ((sM - oe) % A + A) % A, sg = ((sM - e) % A + A) % A;
                const km_win_t wo = WM[so], wg = WG[sg];
                const int16_t *Mo = R[0] + so * RW, *Ie = R[1] + sg * RW, *De = R[2] + sg * RW;
                int16_t *Ic = R[1] + sM * RW, *Dc = R[2] + sM * RW;
                const int hi_ID = imax(wo.hi, wg.hi) + 1, lo_ID = imin(wo.lo, wg.lo) - 1;
                hi = imax(wx.hi, hi_ID);
                lo = imin(wx.lo, lo_ID);
                const int excess = (hi - lo) - (W - 1);
                if (excess > 0) { hi -= (excess + 1) >> 1; lo += excess >> 1; }
                if ((wx.hi - wx.lo) >= W - 1 && (d % band) == 0) {
                    long best_dt = 0x7fffffffL;
                    int c = wx.lo, found = 0;
                    for (int i = wx.lo; i < wx.hi; i++) {
                        const int off = Mx[i - wx.base];
                        if (off < 0) continue;
                        const int left_v = (int16_t)(plen - (off - i)), left_h = (int16_t)(tlen - off);
                        const long dt = imax(left_v, left_h);
                        if (!found || dt < best_dt) { best_dt = dt; c = i; found = 1; }
                    }
                    if (!found || best_dt >= 2L * (tlen + plen)) c = wx.lo;
                    lo = c - (W / 2);
                    hi = lo + W - 1;
                }
                base = lo & ~3;
                const km_win_t wc = {lo, hi, base, ((hi - base) | 3) + 1};
                if (wc.cells > RW) { rc = -1; break; }
                const int nq = ((hi - base) >> 2) + 1;
                uint8_t *row = with_bt ? arena + (size_t)d * RW : NULL;
                /* all reads of a score happen before its writes land in another slot: sM is never its own source */
                for (int q = 0; q < nq; q++)
                    for (int i = 0; i < 4; i++) {
                        const int k = base + 4 * q + i;
                        const int io = km_bq_get(Mo, wo, k - 1), ie = km_bq_get(Ie, wg, k - 1);
                        const int dopen = km_bq_get(Mo, wo, k + 1), dext = km_bq_get(De, wg, k + 1);
                        const int X = (int16_t)(km_bq_get(Mx, wx, k) + 1);
                        const int pI = ie >= io, pD = dext >= dopen;
                        int I = (int16_t)(imax(ie, io) + 1), D = imax(dext, dopen);
                        const int pa = D >= X, T = imax(D, X), pb = T >= I;
                        int M = imax(T, I);
                        if (k < lo || k > hi) { I = KM_NULL; D = KM_NULL; M = KM_NULL; }
                        if (M >= 0) M = km_extend(text, pattern, tlen, plen, k, M);
                        Ic[k - base] = (int16_t)I; Dc[k - base] = (int16_t)D; Mc[k - base] = (int16_t)M;
                        if (with_bt) row[k - base] = (uint8_t)(pI | (pD << 1) | ((pb ? (pa ? 3 : 2) : 1) << 2));
                    }
                WM[sM] = wc;
                WG[sM] = wc;
                base_tab[d] = base;
            }
            if (kt_abs <= d) {
                const int t = (kt >= lo && kt <= hi) ? Mc[kt - base] : KM_NULL;
                if (t == tlen) { fin = 1; dist = d; break; }
                if (t > tlen) break;
            }
        }
    }
    *finished = fin;
    *distance = fin ? dist : 0;
    *n_ops = 0;
    if (fin && with_bt && dist > 0 && rc == 0) {
        int cd = dist, ck = kt, comp = 0, cnt = 0;
#define KM_RES_M(dd) do { while ((dd) > 0 && tab[dd].kind == KM_KIND_NULL) (dd) -= A; } while (0)
#define KM_RES_G(dd) do { while ((dd) > 0 && tab[dd].kind != KM_KIND_MDI) (dd) -= A; } while (0)
        while (!(comp == 0 && cd == 0)) {
            if (cd < 0 || cnt >= ops_cap) { rc = -1; break; }
            const km_step_t st = tab[cd];
            int op;
            if (comp == 0 && st.kind == KM_KIND_M) {
                op = 2; cd -= x; KM_RES_M(cd);
            } else {
                if (st.kind != KM_KIND_MDI) { rc = -1; break; }
                const int ii = ck - base_tab[cd];
                if (ii < 0 || ii >= RW) { rc = -1; break; }
                const uint8_t dec = arena[(size_t)cd * RW + ii];
                if (dec == 0xEE) { rc = -1; break; }                       /* a byte the forward pass never wrote */
                if (comp == 0) {
                    op = 2;
                    const int mop = (dec >> 2) & 3;
                    if (mop == 2) { cd -= x; KM_RES_M(cd); }
                    else if (mop == 1) comp = 1;
                    else comp = 2;
                } else if (comp == 1) {
                    op = 1; ck -= 1;
                    if (dec & 1) { cd -= e; KM_RES_G(cd); } else { cd -= oe; KM_RES_M(cd); comp = 0; }
                } else {
                    op = 3; ck += 1;
                    if (dec & 2) { cd -= e; KM_RES_G(cd); } else { cd -= oe; KM_RES_M(cd); comp = 0; }
                }
            }
            ops_out[cnt++] = (uint8_t)op;
        }
        *n_ops = cnt;
    }
    for (int c = 0; c < 3; c++) free(R[c]);
    free(WM); free(WG); free(arena); free(base_tab);
    return rc;
}
