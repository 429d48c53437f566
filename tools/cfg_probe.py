#!/usr/bin/env python3
"""End-to-end probe of one BASELINE configuration through the public API.
usage: cfg_probe.py <pairs> <length> <err_lo> <err_hi> <max_error> <cigar> <band|0> <window> <batch> [reps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))
import wfagpu
n, L, elo, ehi, me, cigar, band, win, batch = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7]), int(sys.argv[8]), int(sys.argv[9])
reps = int(sys.argv[10]) if len(sys.argv) > 10 else 2
a = wfagpu.Aligner()
a.add_synthetic(0xB2000009, n, L, elo, ehi)
a.initialize_parameters(2, 3, 1)
a.options.max_error = me
a.options.compute_cigar = bool(cigar)
if band > 0:
    a.options.band = band
    a.options.threads_per_block = win
a.set_batch_size(batch)
a.pin_host_buffers()
ts = []
for _ in range(reps + 1):
    a.reset_results()
    t0 = time.perf_counter(); a.align(); ts.append(time.perf_counter() - t0)
st = a.run_stats()
errs = a.errors()
cells = sum(a.s.sequences_metadata[i].pattern_len * a.s.sequences_metadata[i].text_len for i in range(n))
print(json.dumps({"pairs": n, "len": L, "err": [elo, ehi], "max_error": me, "cigar": cigar, "band": band, "window": win,
                  "wall_ms": [round(t * 1e3, 1) for t in ts], "pairs_per_s": round(n / min(ts[1:]), 1),
                  "gcups": round(cells / min(ts[1:]) / 1e9, 1), "mean_score": round(sum(errs) / n, 1), "max_score": max(errs),
                  "redispatched": st["redispatched"], "launches": st["launches"]}))
