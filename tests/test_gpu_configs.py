"""BASELINE configurations 1-3 at their STATED sizes through wfagpu_align (parity was tested at the right
shapes but at 24-4000 pairs): every score equals the unmodified reference CPU WFA (oracle/_ref/libref_cpu.so, all host
threads), every CIGAR is an alignment of exactly that cost (host validator), a sample equals the oracle's CIGAR text.
Config 4 at 8192 pairs is tests/test_gpu_parity.py::test_headline_batch_at_full_size, config 5 on 64 pairs
tests/test_gpu_large_tier.py; bench.py's `configs` block runs all five at full size with parity samples."""
import os

import pytest

import wfagpu
from util import synth_aligner

pytestmark = pytest.mark.gpu
PEN = (2, 3, 1)


def cpu_scores(refcpu, a, lo, hi):
    pairs = [a.pair(i) for i in range(lo, hi)]
    errs, _ = refcpu.align_batch([p for p, _ in pairs], [t for _, t in pairs], *PEN, cigar=False,
                                 threads=len(os.sched_getaffinity(0)))
    return list(errs)


def check_cigars(a, idx):
    pen = wfagpu.AffinePenalties(*PEN)
    bad = 0
    for i in idx:
        p, t = a.pair(i)
        bad += 0 if a.L.wfagpu_check_result(p.encode(), len(p), t.encode(), len(t), pen, a.error(i), a.cigar(i).encode()) else 1
    return bad


def test_config1_10k_pairs_150bp_cigar(oracle, refcpu):
    a = synth_aligner([(10000, 150, 0.02, 0.02)], 0xB2000001)
    assert a.initialize_parameters(*PEN)
    a.options.compute_cigar = True
    a.align()
    assert a.errors() == cpu_scores(refcpu, a, 0, a.num_pairs)
    assert check_cigars(a, range(a.num_pairs)) == 0
    for i in range(0, a.num_pairs, 250):
        p, t = a.pair(i)
        r = oracle.align(p, t, *PEN, 1000)
        assert (a.error(i), a.cigar(i)) == (r["distance"], r["cigar"])


def test_config2_1M_pairs_150bp_score_only(refcpu):
    a = synth_aligner([(1000000, 150, 0.05, 0.05)], 0xB2000002)
    assert a.initialize_parameters(*PEN)
    a.options.compute_cigar = False
    a.align()
    st = a.run_stats()
    assert st["failed_pairs"] == 0
    got = a.errors()
    for lo in range(0, a.num_pairs, 250000):                 # the CPU reference in four slices (memory of the pair lists)
        assert got[lo:lo + 250000] == cpu_scores(refcpu, a, lo, min(lo + 250000, a.num_pairs))


def test_config3_100k_pairs_1kbp_cigar_with_redispatch(oracle, refcpu, monkeypatch):
    # -e 300: ~4 % of the pairs score above it.  First with the budget lift off (every chunk re-dispatches them on the GPU),
    # then with the default policy (a fresh process-wide hint: the first chunks re-dispatch, later chunks and the second call
    # run their first pass with the budget the last batch needed) -- same results either way
    a = synth_aligner([(100000, 1000, 0.10, 0.10)], 0xB2000003)
    assert a.initialize_parameters(*PEN)
    a.options.compute_cigar = True
    a.options.max_error = 300
    monkeypatch.setenv("WFAGPU_HINT_LIFT_PM", "0")
    wfagpu.load().wfagpu_device_close_all()                  # forget hints, contexts re-read the environment
    try:
        a.align()
        st = a.run_stats()
        assert st["redispatched"] > 1000 and st["failed_pairs"] == 0
        want = cpu_scores(refcpu, a, 0, a.num_pairs)
        assert a.errors() == want
        assert check_cigars(a, range(0, a.num_pairs, 7)) == 0
        for i in range(0, a.num_pairs, 5003):
            p, t = a.pair(i)
            r = oracle.align(p, t, *PEN, 100000)
            assert (a.error(i), a.cigar(i)) == (r["distance"], r["cigar"])
        unlifted = [(a.error(i), a.cigar(i)) for i in range(0, a.num_pairs, 11)]
        monkeypatch.delenv("WFAGPU_HINT_LIFT_PM")
        wfagpu.load().wfagpu_device_close_all()
        a.reset_results()
        a.align()
        first = a.run_stats()["redispatched"]
        a.reset_results()
        a.align()
        st = a.run_stats()
        assert st["redispatched"] < first and st["redispatched"] < 200 and st["failed_pairs"] == 0
        assert a.errors() == want
        assert [(a.error(i), a.cigar(i)) for i in range(0, a.num_pairs, 11)] == unlifted
    finally:
        monkeypatch.undo()
        wfagpu.load().wfagpu_device_close_all()
