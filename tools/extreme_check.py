#!/usr/bin/env python3
"""Extreme shapes through the public API against the oracle: identical 30 kbp and 32767-base pairs (the
longest the int16 tier takes), all-mismatch 5 kbp (score 10000: wide wavefronts), half-deleted sequences,
shifted overlaps, one-base and empty sequences; budgets 50 (everything re-dispatched) and 5000."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("oracle", os.path.join("wfa-gpu_b200","python"), "tests"):
    sys.path.insert(0, os.path.join(ROOT,d))
from oracle import Oracle
import wfagpu
O=Oracle(); rng=random.Random(5)
def rnd(n): return "".join(rng.choice("ACGT") for _ in range(n))
b30=rnd(30000); b32=rnd(32767)
cases=[(b30,b30),(b32,b32),(b32[:-1],b32),("A","A"),("A",""),("","")]
p5=rnd(5000); comp={"A":"C","C":"G","G":"T","T":"A"}
cases.append((p5,"".join(comp[c] for c in p5)))          # all mismatches: score 10000
cases.append((p5,p5[:2500]))                              # half deleted
cases.append((p5[:2500],p5))
cases.append((rnd(2000)+p5, p5+rnd(2000)))                # shifted overlap
cases.append((b30[:20000], b30[:20000][:10000]+"T"+b30[:20000][10000:]))
bad=0
for cigar in (True, False):
    for me in (50, 5000):
        a=wfagpu.Aligner()
        for p,t in cases: assert a.add_sequences(p,t)
        assert a.initialize_parameters(2,3,1)
        a.options.compute_cigar=cigar; a.options.max_error=me
        a.align()
        for i,(p,t) in enumerate(cases):
            r=O.align(p,t,2,3,1,60000,cigar=cigar)
            assert r["finished"], i
            ok = a.error(i)==r["distance"] and (not cigar or a.cigar(i)==r["cigar"])
            if not ok:
                bad+=1; print("BAD",cigar,me,i,a.error(i),r["distance"])
        print("cigar",cigar,"budget",me,"stats",a.run_stats()["redispatched"], flush=True)
print("bad",bad); sys.exit(1 if bad else 0)
