"""First-contact GPU script (not a test): runs a few shapes, prints timing and
mismatches against the oracle into gpurun_out/."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wfagpu
from oracle import Oracle
from util import synth_aligner, check_against_oracle
O = Oracle()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", "debug.log"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s); log.write(s + "\n"); log.flush()
shapes = [("150bp2%", [(4000, 150, 0.02, 0.02)], None, 200), ("150bp5%", [(4000, 150, 0.05, 0.05)], None, 200),
          ("1k10%", [(2000, 1000, 0.10, 0.10)], None, 100), ("10k5%", [(600, 10000, 0.05, 0.05)], 3000, 24)]
for name, specs, me, nsample in shapes:
    for cigar in (False, True):
        a = synth_aligner(specs)
        a.initialize_parameters(2, 3, 1)
        a.options.compute_cigar = cigar
        if me: a.options.max_error = me
        a.set_batch_size(a.num_pairs)
        try:
            t0 = time.time(); a.align(); t1 = time.time()
            a2 = synth_aligner(specs); a2.initialize_parameters(2, 3, 1); a2.options.compute_cigar = cigar
            if me: a2.options.max_error = me
            a2.set_batch_size(a2.num_pairs)
            t2 = time.time(); a2.align(); t3 = time.time()
        except Exception as ex:
            P(name, cigar, "FAILED", ex); continue
        st = a2.run_stats()
        step = max(1, a.num_pairs // nsample)
        bad = check_against_oracle(O, a, 2, 3, 1, a.options.max_error, cigar, sample=list(range(0, a.num_pairs, step)))
        P(name, "cigar" if cigar else "score", "pairs", a.num_pairs, "first %.3fs second %.3fs" % (t1 - t0, t3 - t2),
          "aln/s %.0f" % (a.num_pairs / (t3 - t2)), "kernel_ms %.2f" % st["gpu_align_ms"], "pack_ms %.3f" % st["gpu_pack_ms"],
          "redisp", st["redispatched"], "launches", st["launches"], "mismatches", len(bad))
        for b in bad[:5]: P("   ", b)
