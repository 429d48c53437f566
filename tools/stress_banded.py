#!/usr/bin/env python3
"""Randomised parity sweep of the adaptive-band mode (-B <band> -t <window>): random penalties, lengths,
error rates, bands and windows through the public API; every pair that finishes within the budget must
equal the oracle's banded result (score and CIGAR text), and every CIGAR must be a real alignment of the
reported score.   usage: stress_banded.py <seconds> [seed]"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("oracle", os.path.join("wfa-gpu_b200", "python"), "tests"):
    sys.path.insert(0, os.path.join(ROOT, d))
from oracle import Oracle
import wfagpu

budget_s = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 4242)
O = Oracle()
PENS = [(2, 3, 1), (1, 2, 1), (5, 3, 2), (4, 6, 2), (3, 5, 2), (2, 10, 5), (6, 2, 2)]
t0 = time.time()
rounds = pairs_total = bad_total = checked = 0
while time.time() - t0 < budget_s:
    pen = rng.choice(PENS)
    cigar = rng.random() < 0.8
    band = rng.choice([5, 10, 25, 50])
    window = rng.choice([32, 64, 96, 128, 256, 512])
    a = wfagpu.Aligner()
    L = rng.choice([rng.randint(100, 400), rng.randint(400, 1500), rng.randint(1500, 4000)])
    e_lo = rng.choice([0.01, 0.05, 0.10])
    a.add_synthetic(rng.getrandbits(32), max(1, min(300, 120000 // L)), L, e_lo, e_lo + rng.choice([0.0, 0.05]))
    assert a.initialize_parameters(*pen)
    a.options.compute_cigar = cigar
    me = 4 * L                      # generous: the comparison needs pairs that finish in the first pass
    a.options.max_error = me
    a.options.band = band
    a.options.threads_per_block = window
    a.align()
    rounds += 1
    for i in range(a.num_pairs):
        p, t = a.pair(i)
        pairs_total += 1
        if cigar and O.cigar_score(p, t, a.cigar(i), *pen) != a.error(i):
            bad_total += 1
            print("INVALID CIGAR", pen, band, window, i, flush=True)
            continue
        r = O.align(p, t, *pen, me, band=band, window=window, cigar=cigar)
        if not r["finished"]:
            continue                # re-dispatched on the GPU with a larger budget: no reference counterpart
        checked += 1
        if a.error(i) != r["distance"] or (cigar and a.cigar(i) != r["cigar"]):
            bad_total += 1
            print("MISMATCH", pen, band, window, i, a.error(i), r["distance"], flush=True)
print(f"rounds={rounds} pairs={pairs_total} compared={checked} mismatches={bad_total}")
sys.exit(1 if bad_total else 0)
