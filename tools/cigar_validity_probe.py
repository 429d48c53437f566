#!/usr/bin/env python3
"""Aligns <pairs> x <length> (err, CIGAR) twice (unhinted, then hinted) and validates every CIGAR on the host
(wfagpu_check_result): prints the pairs whose text is not an alignment of their score.
usage: cigar_validity_probe.py <pairs> <length> <err> <max_error> [batch]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))
import wfagpu
n, L, err, me = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
batch = int(sys.argv[5]) if len(sys.argv) > 5 else n
a = wfagpu.Aligner()
a.add_synthetic(0xB2000004, n, L, err, err)
a.initialize_parameters(2, 3, 1)
a.options.max_error = me
a.options.compute_cigar = True
a.set_batch_size(batch)
pen = wfagpu.AffinePenalties(2, 3, 1)
for rep in range(2):
    a.reset_results()
    a.align()
    bad = []
    for i in range(a.num_pairs):
        p, t = a.pair(i)
        if not a.L.wfagpu_check_result(p.encode(), len(p), t.encode(), len(t), pen, a.error(i), a.cigar(i).encode()):
            bad.append((i, a.error(i), a.error(i) % 15, a.error(i) % 31, len(a.cigar(i))))
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("WFAGPU_")}, "rep": rep, "pairs": n, "bad": len(bad), "first": bad[:12]}), flush=True)
