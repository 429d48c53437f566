#!/usr/bin/env python3
"""End-to-end timing of the adaptive-band mode (-B band, window W) next to the exact mode on one workload shape.
usage: band_probe.py <pairs> <length> <err_lo> <err_hi> <max_error> [band] [window] [reps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))
import wfagpu
n, L, e0, e1, me = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
band = int(sys.argv[6]) if len(sys.argv) > 6 else 25
win = int(sys.argv[7]) if len(sys.argv) > 7 else 512
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
a = wfagpu.Aligner()
a.add_synthetic(0xB2000004, n, L, e0, e1)
a.initialize_parameters(2, 3, 1)
a.options.max_error = me
a.options.compute_cigar = True
res = {}
for mode in ("exact", "banded"):
    a.options.band = band if mode == "banded" else 0
    a.options.threads_per_block = win
    a.reset_results(); a.align()
    ts = []
    for _ in range(reps):
        a.reset_results()
        t0 = time.perf_counter(); a.align(); ts.append(time.perf_counter() - t0)
    st = a.run_stats()
    res[mode] = {"wall_ms": [round(t * 1e3, 2) for t in ts], "pairs_per_s": round(n / min(ts), 1), "redispatched": st["redispatched"],
                 "errors_sum": sum(a.errors())}
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("WFAGPU_")}, "pairs": n, "len": L, "err": [e0, e1], "band": band, "window": win, **res}))
