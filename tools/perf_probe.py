#!/usr/bin/env python3
"""Times the resident hot path for one workload shape (one process per configuration so
that the WFAGPU_* tuning variables, read once at device open, can differ).
usage: perf_probe.py <pairs> <length> <err> <max_error> <cigar 0|1> [steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))
import wfagpu
n, L, err, me, cigar = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
steps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
a = wfagpu.Aligner()
a.add_synthetic(0xB2000004, n, L, err, err)
a.initialize_parameters(2, 3, 1)
a.options.max_error = me
a.options.compute_cigar = bool(cigar)
rb = wfagpu.ResidentBatch(a)
rb.upload()
plan = rb.plan()
rb.align(plan); rb.wait()
ms, wf = [], []
for _ in range(steps):
    rb.align(plan)
    mp, ma = rb.wait()
    ms.append(ma)
    wf.append(rb.stats()["ms_wavefront"])
out, ops, used = rb.download()
st = rb.stats()
best = min(ms)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("WFAGPU_")}, "pairs": n, "len": L, "err": err,
                  "cigar": cigar, "align_ms": [round(m, 2) for m in ms], "wavefront_ms": round(min(wf), 2), "pairs_per_s": round(n / (best / 1e3), 1),
                  "redispatched": st["redispatched"], "cells": st["cells"], "n_cap": st["n_cap"], "cta_threads": st["cta_threads"], "ctas": st["ctas"], "d_end": st["d_end"]}))
