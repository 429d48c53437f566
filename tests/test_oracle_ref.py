"""Pins the oracle against the UNMODIFIED reference code compiled into oracle/_ref
(CPU only; skipped where the reference tree was never built, e.g. a bare checkout)."""
import random

import pytest


def mutate(rng, t, nerr):
    s = list(t)
    for _ in range(nerr):
        r = rng.random()
        if r < 1 / 3 and s:
            i = rng.randrange(len(s)); s[i] = rng.choice([c for c in "ACGT" if c != s[i]])
        elif r < 2 / 3 and s:
            del s[rng.randrange(len(s))]
        else:
            s.insert(rng.randrange(len(s) + 1), rng.choice("ACGT"))
    return "".join(s)


def make_pairs(seed, shapes):
    rng = random.Random(seed)
    out = []
    for L, err, n in shapes:
        for _ in range(n):
            t = "".join(rng.choice("ACGT") for _ in range(L))
            out.append((mutate(rng, t, int(L * err + 0.999)), t))
    return out


PENS = [(2, 3, 1), (1, 2, 1), (3, 1, 4), (5, 3, 2), (4, 6, 2), (2, 10, 5), (6, 2, 2)]


@pytest.mark.parametrize("pen", PENS)
def test_scores_equal_reference_cpu_wfa(oracle, refcpu, pen):
    pairs = make_pairs(7, [(150, 0.02, 60), (150, 0.05, 60), (1000, 0.1, 12), (20, 0.3, 60), (0, 0, 1)])
    pairs += [("", "ACGT"), ("ACGT", ""), ("ACGT", "ACGT")]
    errs, cigs = refcpu.align_batch([p for p, _ in pairs], [t for _, t in pairs], *pen, cigar=True)
    for (p, t), err, cg in zip(pairs, errs, cigs):
        r = oracle.align(p, t, *pen, 2000)
        assert r["finished"] and r["distance"] == err
        assert oracle.cigar_score(p, t, r["cigar"], *pen) == err
        assert oracle.cigar_score(p, t, cg, *pen) == err


@pytest.mark.parametrize("pen", [(2, 3, 1), (5, 3, 2), (4, 6, 2)])
def test_decoder_equals_reference_recover_cigar_affine(oracle, refcpu, pen):
    # same backtrace chain through utils/cigar.c (reference, compiled in place) and the restatement
    pairs = make_pairs(11, [(150, 0.05, 80), (600, 0.1, 20), (40, 0.3, 80)])
    for p, t in pairs:
        fin, dist, fw, words = oracle.align_chain(p, t, *pen, 1500)
        assert fin
        want = refcpu.recover_cigar(p, t, dist, fw, words)
        got = oracle.align(p, t, *pen, 1500)["cigar"]
        assert got == want
