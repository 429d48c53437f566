"""The CLI bin/wfa.affine.gpu end to end (tests/test-aligner.sh:11-48 and tests/test-fasta.sh:11-42
of the reference): golden scores on wfa.utest.seq for three penalty sets, the `-e 25` run that the
reference resolves with its CPU fallback (here: GPU re-dispatch), FASTA input with -c, banded mode."""
import gzip
import json
import os
import subprocess

import pytest

from util import synth_aligner

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin", "wfa.affine.gpu")


def utest():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "utest.json.gz"), "rt") as f:
        return json.load(f)


def run_cli(args):
    pr = subprocess.run([BIN] + args, capture_output=True, text=True, timeout=600)
    assert pr.returncode == 0, pr.stderr[-3000:]
    return pr


@pytest.mark.parametrize("pi,budget", [(0, 10000), (1, 10000), (2, 10000), (0, 25)])
def test_utest_scores(tmp_path, pi, budget):
    d = utest()
    seq = tmp_path / "utest.seq"
    with open(seq, "w") as f:
        for p, t in zip(d["pattern"], d["text"]):
            f.write(f">{p}\n<{t}\n")
    out = tmp_path / "out.alg"
    x, o, e = d["penalties"][pi]
    pr = run_cli(["-i", str(seq), "-o", str(out), "-g", f"{x},{o},{e}", "-e", str(budget)])
    assert "Alignment computed. Wall time:" in pr.stdout
    got = [-int(l.split("\t")[0]) for l in open(out) if l.strip()]
    assert got == d["scores"][pi]


def test_fasta_cigar_check_and_band(tmp_path, oracle):
    a = synth_aligner([(50, 12000, 0.02, 0.04)], 0xB2005000)
    q, t = tmp_path / "q.fasta", tmp_path / "t.fasta"
    with open(q, "w") as fq, open(t, "w") as ft:
        for i in range(a.num_pairs):
            p, s = a.pair(i)
            fq.write(f">q{i}\n" + "\n".join(p[j:j + 70] for j in range(0, len(p), 70)) + "\n")
            ft.write(f">t{i}\n" + "\n".join(s[j:j + 70] for j in range(0, len(s), 70)) + "\n")
    out = tmp_path / "o.txt"
    pr = run_cli(["-Q", str(q), "-T", str(t), "-b", "50", "-o", str(out), "-x", "-c"])
    assert "correct=50 Incorrect=0" in pr.stderr
    rows = [l.rstrip("\n").split("\t") for l in open(out)]
    for i in (0, 17, 49):
        p, s = a.pair(i)
        r = oracle.align(p, s, 2, 3, 1, 4000)
        assert (-int(rows[i][0]), rows[i][1]) == (r["distance"], r["cigar"])
    out2 = tmp_path / "o2.txt"
    pr = run_cli(["-Q", str(q), "-T", str(t), "-b", "50", "-o", str(out2), "-x", "-c", "-B", "auto", "-t", "512", "-O"])
    assert "Banded execution. Band width: 512. Band re-centering every 25 steps" in pr.stderr
    assert "correct=50 Incorrect=0" in pr.stderr
    rows2 = [l.rstrip("\n").split("\t") for l in open(out2)]
    assert len(rows2[0]) == 4 and rows2[0][2:] == list(a.pair(0))
    p, s = a.pair(3)
    r = oracle.align(p, s, 2, 3, 1, 3000, band=25, window=512)
    assert (-int(rows2[3][0]), rows2[3][1]) == (r["distance"], r["cigar"])


def test_streaming_windows_give_the_same_output(tmp_path):
    # -S N: the input is read / aligned / written N pairs at a time, the next window being read into page-locked
    # memory while the GPU aligns the current one; same lines, same order
    d = utest()
    seq = tmp_path / "utest.seq"
    with open(seq, "w") as f:
        for p, t in zip(d["pattern"], d["text"]):
            f.write(f">{p}\n<{t}\n")
    whole, windows, capped = tmp_path / "whole.alg", tmp_path / "win.alg", tmp_path / "cap.alg"
    run_cli(["-i", str(seq), "-o", str(whole), "-x", "-e", "400"])
    pr = run_cli(["-i", str(seq), "-o", str(windows), "-x", "-e", "400", "-S", "64", "-c"])
    assert "Streaming in windows of 64 pairs" in pr.stderr
    assert "correct=305 Incorrect=0" in pr.stderr
    assert open(whole).read() == open(windows).read()
    run_cli(["-i", str(seq), "-o", str(capped), "-x", "-e", "400", "-S", "64", "-n", "100"])
    assert open(capped).read().splitlines() == open(whole).read().splitlines()[:100]
