#!/usr/bin/env python3
"""cfg 5 shape (50 kbp, 15 %, CIGAR, first budget below the scores): end-to-end time of one wfagpu_align call.
usage: cfg5_probe.py <pairs> [length] [err] [max_error]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))
import wfagpu
n = int(sys.argv[1]); L = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
err = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15; me = int(sys.argv[4]) if len(sys.argv) > 4 else 8000
a = wfagpu.Aligner()
a.add_synthetic(0xB2000005, n, L, err, err)
a.initialize_parameters(2, 3, 1)
a.options.max_error = me
a.options.compute_cigar = True
ts = []
for _ in range(2):
    a.reset_results()
    t0 = time.perf_counter(); a.align(); ts.append(time.perf_counter() - t0)
st = a.run_stats()
print(json.dumps({"pairs": n, "len": L, "err": err, "wall_s": [round(t, 3) for t in ts], "pairs_per_s": round(n / min(ts), 1),
                  "redispatched": st["redispatched"], "launches": st["launches"], "mean_score": sum(a.errors()) / n}))
