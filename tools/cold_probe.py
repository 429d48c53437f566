#!/usr/bin/env python3
"""Cold-start timing of wfagpu_align in a fresh process: three calls on the headline batch, wall time of each.
usage: cold_probe.py [pairs] [length] [err] [max_error]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))
t00 = time.perf_counter()
import wfagpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
err = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
me = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
a = wfagpu.Aligner(); a.add_synthetic(0xB2000004, n, L, err, err); a.initialize_parameters(2, 3, 1)
a.options.max_error = me; a.options.compute_cigar = True
t_setup = time.perf_counter() - t00
ts = []
for _ in range(3):
    a.reset_results(); t0 = time.perf_counter(); a.align(); ts.append(time.perf_counter() - t0)
print(json.dumps({"setup_ms": round(t_setup * 1e3, 1), "call_ms": [round(t * 1e3, 1) for t in ts], "stats": a.run_stats()}))
