#!/usr/bin/env python3
"""Resident step time for the bench shards of ranks 0..7 (different seeds -> different largest scores -> ring widths
on both sides of the five-CTAs-per-SM limit).  usage: cliff_probe.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))
import wfagpu
for rank in range(8):
    a = wfagpu.Aligner()
    a.add_synthetic(0xB2000004 + 7919 * rank, 8192, 10000, 0.05, 0.05)
    a.initialize_parameters(2, 3, 1)
    a.options.max_error = 3000
    a.options.compute_cigar = True
    rb = wfagpu.ResidentBatch(a)
    rb.upload()
    plan = rb.plan()
    ms = []
    for _ in range(4):
        rb.align(plan)
        mp, ma = rb.wait()
        ms.append(mp + ma)
    st = rb.stats()
    print(json.dumps({"rank": rank, "ms": round(min(ms[1:]), 2), "n_cap": st["n_cap"], "cta_threads": st["cta_threads"], "ctas": st["ctas"], "d_end": st["d_end"]}), flush=True)
    rb.release()
    a.destroy()
