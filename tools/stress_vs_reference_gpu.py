#!/usr/bin/env python3
"""Randomised comparison with the UNMODIFIED reference GPU binary (oracle/_ref/gpu/wfa.affine.gpu, built for
sm_100 by `make -C oracle refgpu`): random penalties, lengths, error rates, exact and banded; scores and CIGAR
text must be byte-identical for every pair the reference GPU kernels finish (pairs its CPU fallback finishes
use a different tie-break and are compared by score only in exact mode).
usage: stress_vs_reference_gpu.py <seconds> [seed]"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("oracle", os.path.join("wfa-gpu_b200", "python"), "tests"):
    sys.path.insert(0, os.path.join(ROOT, d))
import refgpu
import wfagpu
from oracle import Oracle

if not refgpu.available():
    print("reference GPU binary not built")
    sys.exit(0)
budget_s = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
O = Oracle()
PENS = [(2, 3, 1), (1, 2, 1), (5, 3, 2), (4, 6, 2), (3, 5, 2), (2, 10, 5), (6, 2, 2), (3, 1, 4)]
t0 = time.time()
rounds = pairs_total = same = score_only = bad = ref_failed = 0
while time.time() - t0 < budget_s:
    pen = rng.choice(PENS)
    banded = rng.random() < 0.35
    L = rng.choice([rng.randint(50, 300), rng.randint(300, 1500), rng.randint(1500, 6000)])
    n = max(4, min(200, 150000 // L))
    e_lo = rng.choice([0.01, 0.04, 0.08])
    a = wfagpu.Aligner()
    a.add_synthetic(rng.getrandbits(32), n, L, e_lo, e_lo + rng.choice([0.0, 0.04]))
    pairs = [a.pair(i) for i in range(a.num_pairs)]
    me = max(64, int(L * 0.6))
    band = rng.choice([10, 25, 50]) if banded else None
    window = rng.choice([64, 128, 256, 512]) if banded else None
    assert a.initialize_parameters(*pen)
    a.options.compute_cigar = True
    a.options.max_error = me
    if banded:
        a.options.band = band
        a.options.threads_per_block = window
    a.align()
    try:
        ref, _, _ = refgpu.run(pairs, pen, me, cigar=True, band=band, threads=window)
    except RuntimeError as ex:           # the reference binary itself gives up on some parameter combinations
        ref_failed += 1
        print("reference failed:", pen, band, window, L, me, str(ex).strip().splitlines()[-1][:120], flush=True)
        continue
    if len(ref) != len(pairs):
        ref_failed += 1
        continue
    rounds += 1
    for i, (p, t) in enumerate(pairs):
        pairs_total += 1
        r = O.align(p, t, *pen, me, band=band if banded else -1, window=window or 0, cigar=False)
        mine = (a.error(i), a.cigar(i))
        if r["finished"]:
            # finished by the reference's GPU kernels: byte-identical
            if mine == ref[i]:
                same += 1
            else:
                bad += 1
                print("MISMATCH", pen, band, window, L, i, mine[0], ref[i][0], flush=True)
        elif not banded:
            # the reference finished it on the CPU (other tie-breaks): the optimal score must agree
            score_only += 1
            if mine[0] != ref[i][0]:
                bad += 1
                print("SCORE MISMATCH", pen, L, i, mine[0], ref[i][0], flush=True)
print(f"rounds={rounds} pairs={pairs_total} identical={same} score_only={score_only} mismatches={bad} reference_failed={ref_failed}")
sys.exit(1 if bad else 0)
