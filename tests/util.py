"""Shared helpers for the parity tests."""
import wfagpu


def synth_aligner(specs, seed=0xB2000000):
    """specs: list of (n, length, err_lo, err_hi). Returns a filled Aligner."""
    a = wfagpu.Aligner()
    for i, (n, length, lo, hi) in enumerate(specs):
        a.add_synthetic(seed + i, n, length, lo, hi)
    return a


def pairs_of(a, idx=None):
    idx = range(a.num_pairs) if idx is None else idx
    return [a.pair(i) for i in idx]


def check_against_oracle(oracle, a, x, o, e, max_steps, cigar, sample=None, big_budget=None):
    """Compare every (or sampled) pair of a finished Aligner run with the oracle.

    The GPU library re-dispatches over-budget pairs, so its final answer must equal
    the oracle run with a budget large enough to finish (big_budget)."""
    idx = list(range(a.num_pairs)) if sample is None else sample
    bad = []
    for i in idx:
        p, t = a.pair(i)
        r = oracle.align(p, t, x, o, e, max_steps, cigar=cigar)
        if not r["finished"]:
            r = oracle.align(p, t, x, o, e, big_budget or 8 * max_steps + 64, cigar=cigar)
            assert r["finished"]
        if a.error(i) != r["distance"]:
            bad.append((i, "score", a.error(i), r["distance"]))
        elif cigar and a.cigar(i) != r["cigar"]:
            bad.append((i, "cigar", a.cigar(i)[:80], r["cigar"][:80]))
    return bad
