"""Run in a subprocess with WFAGPU_FORCE_LARGE=1: every pair goes through the large tier
(rings in global memory, int32 offsets); results must still be bit-exact vs the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import Oracle
from util import synth_aligner, check_against_oracle
O = Oracle()
bad_total = 0
for pen, cigar in (((2, 3, 1), True), ((2, 3, 1), False), ((5, 3, 2), True), ((4, 6, 2), True)):
    a = synth_aligner([(200, 150, 0.05, 0.05), (60, 1000, 0.10, 0.10), (8, 10000, 0.05, 0.05)])
    assert a.initialize_parameters(*pen)
    a.options.compute_cigar = cigar
    a.options.max_error = 300          # 1 kbp and 10 kbp pairs exceed it: re-dispatch inside the large tier too
    a.align()
    bad = check_against_oracle(O, a, *pen, 300, cigar, big_budget=6000)
    print(pen, cigar, "mismatches", len(bad), bad[:3], a.run_stats()["redispatched"])
    bad_total += len(bad)
sys.exit(1 if bad_total else 0)
