#!/usr/bin/env python3
"""End-to-end timing of wfagpu_align for one workload shape (pinned host buffers).
usage: e2e_probe.py <pairs> <length> <err> <max_error> <cigar 0|1> <batch> [reps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))
import wfagpu
n, L, err, me, cigar, batch = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
a = wfagpu.Aligner()
a.add_synthetic(0xB2000004, n, L, err, err)
a.initialize_parameters(2, 3, 1)
a.options.max_error = me
a.options.compute_cigar = bool(cigar)
a.set_batch_size(batch) if batch < n else None
a.pin_host_buffers()
a.align()
ts = []
for _ in range(reps):
    a.reset_results()
    t0 = time.perf_counter(); a.align(); ts.append(time.perf_counter() - t0)
st = a.run_stats()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("WFAGPU_")}, "pairs": n, "len": L, "cigar": cigar, "batch": batch,
                  "wall_ms": [round(t * 1e3, 2) for t in ts], "pairs_per_s": round(n / min(ts), 1), "gpu_align_ms": round(st["gpu_align_ms"], 2),
                  "h2d_MB": round(st["h2d_bytes"] / 1e6, 1), "d2h_MB": round(st["d2h_bytes"] / 1e6, 1), "launches": st["launches"]}))
