"""Run in a subprocess with WFAGPU_* variables that select a kernel variant (read once when the
device opens): WFAGPU_FORCE_BOUND=1 (per-pair score bounds always), WFAGPU_NO_BOUND=1 (launch
bound only), WFAGPU_NO_CKPT=1 (decision bytes instead of ring snapshots), WFAGPU_CK_PERIOD=7|15|31,
WFAGPU_ARENA_MB=n (snapshot arenas capped: the pass runs in several sub-launches).
Whatever the variant, scores and CIGARs must be bit-exact vs the oracle, including pairs that
outgrow the first pass and are re-dispatched."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import Oracle
from util import synth_aligner, check_against_oracle
O = Oracle()
bad_total = 0
for pen, cigar in (((2, 3, 1), True), ((2, 3, 1), False), ((5, 3, 2), True), ((4, 6, 2), True), ((1, 0, 1), True), ((2, 24, 9), True)):
    # two batches through the same device: the second one runs with the provisioning hint
    for rep in range(2):
        a = synth_aligner([(300, 150, 0.05, 0.05), (80, 1100, 0.02, 0.12), (10, 6000, 0.01, 0.06), (6, 3000, 0.20, 0.20)],
                          0xB2005000 + rep)
        a.add_sequences("ACGT" * 300, "ACGT" * 300 + "T" * 150)      # long gap: target diagonal far from 0
        a.add_sequences("GATTACA" * 200 + "C" * 211, "GATTACA" * 200)
        for pp, tt in (("", "ACGT"), ("ACGT", ""), ("ACGT", "ACGT"), ("A", "C"), ("ACGTACGTAC", "TTTTTTTTTT"), ("A", "A" * 40)):
            a.add_sequences(pp, tt)
        assert a.initialize_parameters(*pen)
        a.options.compute_cigar = cigar
        a.options.max_error = 400          # the 3 kbp / 20 % pairs exceed it: re-dispatched
        a.align()
        bad = check_against_oracle(O, a, *pen, 400, cigar, big_budget=12000)
        print(pen, cigar, rep, "mismatches", len(bad), bad[:3], a.run_stats())
        bad_total += len(bad)
# adaptive band (packed pairs: wfa_bandq_kernel + wfa_band_traceback_kernel; WFAGPU_NO_BAND_TB=1: backtrace inside the kernel;
# WFAGPU_NO_QUAD=1: wfa_banded_kernel, one diagonal per thread): every pair the band finishes equals the oracle's banded result
for pen, cigar, band, window, specs in (((2, 3, 1), True, 10, 64, [(120, 1000, 0.03, 0.10)]), ((2, 3, 1), True, 25, 512, [(12, 6000, 0.04, 0.06)]),
                                        ((4, 6, 2), True, 25, 128, [(60, 1500, 0.05, 0.08)]), ((2, 3, 1), False, 5, 96, [(100, 800, 0.05, 0.05)])):
    a = synth_aligner(specs, 0xB2005100 + band)
    a.add_sequences("ACGT" * 200, "ACGT" * 200 + "T" * 90)
    assert a.initialize_parameters(*pen)
    a.options.compute_cigar = cigar
    a.options.max_error = 2000
    a.options.band = band
    a.options.threads_per_block = window
    a.align()
    bad = 0
    for i in range(a.num_pairs):
        pp, tt = a.pair(i)
        r = O.align(pp, tt, *pen, 2000, band=band, window=window, cigar=cigar)
        if r["finished"]:
            bad += (a.error(i) != r["distance"]) or (cigar and a.cigar(i) != r["cigar"])
    print("banded", pen, cigar, band, window, "mismatches", bad, a.run_stats())
    bad_total += bad
sys.exit(1 if bad_total else 0)
