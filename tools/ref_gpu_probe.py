#!/usr/bin/env python3
"""Times the UNMODIFIED reference GPU binary (oracle/_ref/gpu) on the same synthetic pairs.
usage: ref_gpu_probe.py <pairs> <length> <err> <max_error> <cigar 0|1> [extra CLI args...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wfagpu, refgpu
n, L, err, me, cigar = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
a = wfagpu.Aligner()
a.add_synthetic(0xB2000004, n, L, err, err)
pairs = [a.pair(i) for i in range(n)]
kw = {}
if len(sys.argv) > 6: kw["batch"] = int(sys.argv[6])
if len(sys.argv) > 7: kw["threads"] = int(sys.argv[7])
res, wall, total = refgpu.run(pairs, (2, 3, 1), me, cigar=bool(cigar), **kw)
print(json.dumps({"impl": "reference-gpu", "pairs": n, "len": L, "err": err, "cigar": cigar, "tool_wall_s": wall,
                  "pairs_per_s": round(n / wall, 1) if wall else None, "total_s": round(total, 2), "kw": kw}))
