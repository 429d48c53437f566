#!/usr/bin/env python3
"""Extracts the reference's own golden vectors for the hot path into small fixtures.

Run in the build container (needs /root/reference); the outputs are committed:
  utest.json.gz        tests/data/wfa.utest.seq (305 pairs) + score goldens
                       tests/data/results/test.score.affine.p{0,1,2}.alg for penalties
                       (1,2,1) (3,1,4) (5,3,2)  [tests/test-aligner.sh:11-46] and the CPU-WFA
                       CIGAR goldens external/WFA/tests/wfa.utest.check/test.affine.p{0,1,2}.alg
                       (validity reference only: CPU tie-breaks differ from the GPU path)
  api_10k.json.gz      tests/data/sequences_10K.h: 100 pairs x 10 kbp + goldens x2o3e1 / x3o5e2
  api_1000.json.gz     tests/data/sequences_1000.h: first 300 of 1000 pairs x 1 kbp + goldens
                       x2o3e1 / x5o3e2        [asserted by tests/test_api.c:59-219]
  test_hifi.{query,target}.fasta.gz + test_hifi.json
                       tests/data/test_hifi.*.fasta (50 HiFi pairs, the reference's FASTA fixture,
                       tests/test-fasta.sh:11-22) verbatim, with the record lengths and the scores of the
                       UNMODIFIED reference CPU WFA (oracle/_ref/libref_cpu.so) for penalties (2,3,1) and (5,2,5)
"""
import gzip, json, os, re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def dump(name, obj):
    with gzip.open(os.path.join(OUT, name), "wt", compresslevel=9) as f:
        json.dump(obj, f, separators=(",", ":"))


def utest():
    pats, txts = [], []
    for line in open(f"{REF}/tests/data/wfa.utest.seq"):
        line = line.rstrip("\n")
        if line.startswith(">"):
            pats.append(line[1:])
        elif line.startswith("<"):
            txts.append(line[1:])
    assert len(pats) == len(txts) == 305
    obj = {"pattern": pats, "text": txts, "penalties": [[1, 2, 1], [3, 1, 4], [5, 3, 2]], "scores": [], "cpu_cigars": []}
    for p in range(3):
        sc = [-int(l.split()[0]) for l in open(f"{REF}/tests/data/results/test.score.affine.p{p}.alg") if l.strip()]
        assert len(sc) == 305
        obj["scores"].append(sc)
        cg = [l.split()[1] for l in open(f"{REF}/external/WFA/tests/wfa.utest.check/test.affine.p{p}.alg") if l.strip()]
        assert len(cg) == 305
        obj["cpu_cigars"].append(cg)
    dump("utest.json.gz", obj)


def header(path, seq_name, golden_names, limit=None):
    src = open(path).read()
    m = re.search(r"%s\[\d+\]\s*=\s*\{(.*?)\};" % seq_name, src, re.S)
    seqs = re.findall(r'"([ACGTN]*)"', m.group(1))
    obj = {"pattern": seqs[0::2], "text": seqs[1::2], "goldens": {}}
    for g in golden_names:
        m = re.search(r"%s\[\d+\]\s*=\s*\{(.*?)\};" % g, src, re.S)
        obj["goldens"][g] = [-int(v) for v in re.findall(r"-?\d+", m.group(1))]
        assert len(obj["goldens"][g]) == len(obj["pattern"]), (g, len(obj["goldens"][g]))
    if limit:
        obj["pattern"] = obj["pattern"][:limit]
        obj["text"] = obj["text"][:limit]
        for g in golden_names:
            obj["goldens"][g] = obj["goldens"][g][:limit]
    return obj


def hifi():
    import shutil, sys
    sys.path.insert(0, os.path.join(os.path.dirname(OUT), "..", "oracle"))
    from oracle import RefCPU
    recs = {}
    for side in ("query", "target"):
        src = f"{REF}/tests/data/test_hifi.{side}.fasta"
        with open(src, "rb") as f, gzip.open(os.path.join(OUT, f"test_hifi.{side}.fasta.gz"), "wb", compresslevel=9) as g:
            shutil.copyfileobj(f, g)
        seqs, cur = [], None
        for line in open(src):
            line = line.strip()
            if not line:
                continue
            if line.lstrip().startswith(">"):
                if cur is not None:
                    seqs.append("".join(cur))
                cur = []
            else:
                cur.append(line)
        seqs.append("".join(cur))
        recs[side] = seqs
    assert len(recs["query"]) == len(recs["target"]) == 50
    r = RefCPU()
    obj = {"lengths": [[len(q), len(t)] for q, t in zip(recs["query"], recs["target"])], "scores": {}}
    for pen in ((2, 3, 1), (5, 2, 5)):
        errs, _ = r.align_batch(recs["query"], recs["target"], *pen, cigar=False)
        obj["scores"]["%d,%d,%d" % pen] = list(errs)
    with open(os.path.join(OUT, "test_hifi.json"), "w") as f:
        json.dump(obj, f)


if __name__ == "__main__":
    utest()
    hifi()
    dump("api_10k.json.gz", header(f"{REF}/tests/data/sequences_10K.h", "sequences_10K_n100",
                                   ["results_10K_n100_x2o3e1", "results_10K_n100_x3o5e2"]))
    dump("api_1000.json.gz", header(f"{REF}/tests/data/sequences_1000.h", "sequences_1000_n1000",
                                    ["results_1000_n1000_x2o3e1", "results_1000_n1000_x5o3e2"], limit=300))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
