#!/usr/bin/env python3
"""Randomised parity sweep on a GPU box: random penalties, lengths, error rates, length offsets and
batch sizes through the public API, every pair checked against the oracle (score and CIGAR text).
usage: stress_parity.py <seconds> [seed]     (WFAGPU_* variables select kernel variants as usual)"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("oracle", os.path.join("wfa-gpu_b200", "python"), "tests"):
    sys.path.insert(0, os.path.join(ROOT, d))
from oracle import Oracle
import wfagpu
from util import check_against_oracle

budget_s = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
O = Oracle()
PENS = [(2, 3, 1), (1, 2, 1), (3, 1, 4), (5, 3, 2), (4, 6, 2), (2, 10, 5), (6, 2, 2), (1, 0, 1), (3, 5, 2), (7, 11, 3),
        (2, 24, 9), (1, 1, 1), (9, 1, 1), (2, 0, 3)]
t0 = time.time()
rounds = pairs_total = bad_total = 0
while time.time() - t0 < budget_s:
    pen = rng.choice(PENS)
    cigar = rng.random() < 0.8
    a = wfagpu.Aligner()
    n_groups = rng.randint(1, 4)
    for g in range(n_groups):
        L = rng.choice([rng.randint(1, 40), rng.randint(40, 400), rng.randint(400, 1500), rng.randint(1500, 3000)])
        cnt = max(1, min(400, 60000 // max(L, 30)))
        e_lo = rng.choice([0.0, 0.01, 0.05, 0.15])
        a.add_synthetic(rng.getrandbits(32), cnt, L, e_lo, e_lo + rng.choice([0.0, 0.05, 0.15]))
    # a few structured pairs: long gaps on either side, identical, empty, homopolymers
    base = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 600)))
    gap = rng.randint(1, 200)
    for p, t in ((base, base + "T" * gap), (base + "G" * gap, base), (base, base), ("", base[:50]), (base[:50], ""),
                 ("A" * rng.randint(1, 300), "A" * rng.randint(1, 300)), (base, base[::-1])):
        a.add_sequences(p, t)
    assert a.initialize_parameters(*pen)
    a.options.compute_cigar = cigar
    me = rng.choice([20, 100, 400, 2000])
    a.options.max_error = me
    a.set_batch_size(rng.choice([a.num_pairs, max(1, a.num_pairs // 3), 97]))
    a.align()
    bad = check_against_oracle(O, a, *pen, me, cigar, big_budget=60000)
    rounds += 1
    pairs_total += a.num_pairs
    bad_total += len(bad)
    if bad:
        print("MISMATCH", pen, cigar, me, bad[:3], flush=True)
print(f"rounds={rounds} pairs={pairs_total} mismatches={bad_total} env={ {k: v for k, v in os.environ.items() if k.startswith('WFAGPU_')} }")
sys.exit(1 if bad_total else 0)
