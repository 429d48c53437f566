"""Pins the oracle (restatement of the reference GPU path) against the reference's
own golden vectors (tests/golden/, made by make_golden.py from the reference's
tests/data).  CPU only."""
import gzip
import json
import os

import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with gzip.open(os.path.join(G, name), "rt") as f:
        return json.load(f)


@pytest.mark.parametrize("pi", [0, 1, 2])
def test_utest_scores_and_cigars(oracle, pi):
    # tests/test-aligner.sh:11-46: wfa.utest.seq, -g 1,2,1 / 3,1,4 / 5,3,2, -e 10000
    d = load("utest.json.gz")
    x, o, e = d["penalties"][pi]
    same_text = 0
    for p, t, gold, cpu_cigar in zip(d["pattern"], d["text"], d["scores"][pi], d["cpu_cigars"][pi]):
        budget = min(10000, 2 * gold + 64)          # any budget that finishes gives the same answer
        r = oracle.align(p, t, x, o, e, budget)
        assert r["finished"]
        assert r["distance"] == gold
        assert oracle.cigar_score(p, t, r["cigar"], x, o, e) == gold
        # the CPU library breaks ties differently (X > D > I): its CIGAR is only a validity reference
        assert oracle.cigar_score(p, t, cpu_cigar, x, o, e) == gold
        same_text += r["cigar"] == cpu_cigar
    assert same_text > 100  # informational: most short pairs have no D/X tie


def test_utest_budget_rule(oracle):
    # tests/test-aligner.sh "test CPU recovery" runs with -e 25: pairs above the budget must come
    # back unfinished (finished <=> enough MDI steps), everything else keeps its golden score.
    d = load("utest.json.gz")
    x, o, e = d["penalties"][0]
    n_unfinished = 0
    for p, t, gold in zip(d["pattern"], d["text"], d["scores"][0]):
        r = oracle.align(p, t, x, o, e, 25, cigar=False)
        if r["finished"]:
            assert r["distance"] == gold
        else:
            n_unfinished += 1
            assert gold > 20
    assert n_unfinished > 0


@pytest.mark.parametrize("golden,pen", [("results_10K_n100_x2o3e1", (2, 3, 1)), ("results_10K_n100_x3o5e2", (3, 5, 2))])
def test_api_10k_goldens(oracle, golden, pen):
    # tests/test_api.c:59-135; a 20-pair slice keeps the CPU suite short (score-only kernels)
    d = load("api_10k.json.gz")
    for p, t, gold in list(zip(d["pattern"], d["text"], d["goldens"][golden]))[:20]:
        r = oracle.align(p, t, *pen, gold + 8, cigar=False)
        assert r["finished"] and r["distance"] == gold


def test_api_10k_cigar_golden_score(oracle):
    d = load("api_10k.json.gz")
    for p, t, gold in list(zip(d["pattern"], d["text"], d["goldens"]["results_10K_n100_x2o3e1"]))[:3]:
        r = oracle.align(p, t, 2, 3, 1, 3000)
        assert r["finished"] and r["distance"] == gold
        assert oracle.cigar_score(p, t, r["cigar"], 2, 3, 1) == gold


@pytest.mark.parametrize("golden,pen", [("results_1000_n1000_x2o3e1", (2, 3, 1)), ("results_1000_n1000_x5o3e2", (5, 3, 2))])
def test_api_1000_goldens(oracle, golden, pen):
    # tests/test_api.c:137-217
    d = load("api_1000.json.gz")
    for p, t, gold in zip(d["pattern"], d["text"], d["goldens"][golden]):
        r = oracle.align(p, t, *pen, gold + 8)
        assert r["finished"] and r["distance"] == gold
        assert oracle.cigar_score(p, t, r["cigar"], *pen) == gold
