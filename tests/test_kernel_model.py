"""The CPU model of the B200 kernel's data flow (oracle/kernel_model.c: penalty-only step
table, clean rings with NULL guards, decision bytes + traceback, and the checkpointed
traceback that recomputes offsets from ring snapshots) must give exactly what the faithful
restatement gives.  CPU only."""
import pytest

from test_oracle_ref import make_pairs, PENS


@pytest.mark.parametrize("pen", PENS + [(1, 0, 1), (3, 5, 2), (7, 11, 3)])
def test_model_equals_oracle(oracle, pen):
    pairs = make_pairs(3, [(150, 0.05, 80), (700, 0.1, 10), (25, 0.35, 120), (0, 0, 1)])
    pairs += [("", "ACGT"), ("ACGT", ""), ("ACGT", "ACGT"), ("A", "C"), ("ACGTACGTAC", "TTTTTTTTTT")]
    for p, t in pairs:
        r = oracle.align(p, t, *pen, 1200)
        m = oracle.model_align(p, t, *pen, 1200)
        assert r["finished"]
        assert (m["finished"], m["distance"], m["cigar"]) == (r["finished"], r["distance"], r["cigar"])


@pytest.mark.parametrize("period", [4, 7, 16, 32])
@pytest.mark.parametrize("pen", PENS + [(1, 0, 1), (3, 5, 2), (7, 11, 3), (2, 24, 9)])
def test_checkpointed_traceback_equals_oracle(oracle, pen, period):
    # ring snapshots every `period` scores + recomputation on the dependency cone
    # (the exact kernel's CIGAR path, wfa_traceback_kernel) -- same CIGARs as the reference walk
    pairs = make_pairs(7, [(150, 0.05, 30), (700, 0.1, 6), (25, 0.35, 40), (0, 0, 1)])
    pairs += [("", "ACGT"), ("ACGT", ""), ("ACGT", "ACGT"), ("A", "C"), ("ACGTACGTAC", "TTTTTTTTTT")]
    for p, t in pairs:
        r = oracle.align(p, t, *pen, 1200)
        m = oracle.model_align_ckpt(p, t, *pen, 1200, period)
        assert r["finished"]
        assert (m["finished"], m["distance"], m["cigar"]) == (r["finished"], r["distance"], r["cigar"])


@pytest.mark.parametrize("pen", PENS + [(1, 0, 1), (3, 5, 2), (7, 11, 3), (2, 24, 9)])
def test_score_bound_pruning_is_exact(oracle, pen):
    # Cells with d + e*|k - k_target| > dmax are never computed (the model poisons them, so a read of
    # one would corrupt the result): every pair that finishes with score s <= dmax must come out with
    # the reference's score and CIGAR, down to the tightest bound dmax = s; and with dmax = s - 1 it
    # must not finish at all (the GPU re-dispatches it with a larger budget).
    pairs = make_pairs(11, [(150, 0.05, 12), (700, 0.1, 3), (25, 0.35, 20), (300, 0.2, 4)])
    pairs += [("", "ACGT"), ("ACGT", ""), ("ACGT", "ACGT"), ("A", "C"), ("ACGTACGTAC", "TTTTTTTTTT"),
              ("ACGT" * 30, "ACGT" * 30 + "T" * 20), ("GATTACA" * 20 + "C" * 37, "GATTACA" * 20)]
    for p, t in pairs:
        r = oracle.align(p, t, *pen, 1200)
        assert r["finished"]
        s = r["distance"]
        for dmax in (s, s + 1, s + 3, int(1.2 * s) + 2, 3 * s + 7):
            for period in (7, 16):
                m = oracle.model_align_ckpt(p, t, *pen, 1200, period, dmax)
                assert (m["finished"], m["distance"], m["cigar"]) == (True, s, r["cigar"])
        if s > 0:
            assert not oracle.model_align_ckpt(p, t, *pen, 1200, 16, s - 1)["finished"]


@pytest.mark.parametrize("pen", [(2, 3, 1), (4, 6, 2), (5, 3, 2)])
def test_model_budget_rule_equals_oracle(oracle, pen):
    # which pairs are over budget must match the reference rule (steps < max_steps - 1)
    pairs = make_pairs(5, [(300, 0.1, 60)])
    for budget in (12, 20, 31):
        for p, t in pairs:
            r = oracle.align(p, t, *pen, budget, cigar=False)
            m = oracle.model_align(p, t, *pen, budget, cigar=False)
            assert m["finished"] == r["finished"]
            if r["finished"]:
                assert m["distance"] == r["distance"]


@pytest.mark.parametrize("pen,band,window", [((2, 3, 1), 10, 64), ((2, 3, 1), 25, 128), ((2, 3, 1), 5, 32), ((2, 3, 1), 25, 30),
                                             ((4, 6, 2), 25, 96), ((5, 3, 2), 10, 64), ((1, 0, 1), 7, 48), ((3, 5, 2), 50, 128)])
def test_banded_quad_layout_equals_oracle(oracle, pen, band, window):
    # wfa_bandq_kernel's data layout (rows at k - base in whole quads with NULL outside the window, reads outside the stored
    # part of a row yield NULL, window records shared by the I and D rows, decisions from "which operand won") against the
    # restatement of the reference's banded kernels: same finished flag, score and CIGAR -- also where the band loses the
    # alignment, where windows jump at a re-centre, and with null steps (gcd > 1)
    pairs = make_pairs(13, [(600, 0.08, 14), (1500, 0.04, 5), (300, 0.15, 12), (90, 0.05, 10), (0, 0, 1)])
    pairs += [("", "ACGT"), ("ACGT", ""), ("ACGT", "ACGT"), ("A", "C"), ("ACGT" * 60, "ACGT" * 60 + "T" * 45),
              ("GATTACA" * 40 + "C" * 77, "GATTACA" * 40)]
    n_fin = 0
    for p, t in pairs:
        for budget in (1500, 60):
            r = oracle.align(p, t, *pen, budget, band=band, window=window)
            m = oracle.model_align_bandq(p, t, *pen, budget, band, window)
            assert m["finished"] == r["finished"]
            if r["finished"]:
                assert (m["distance"], m["cigar"]) == (r["distance"], r["cigar"])
                n_fin += 1
    assert n_fin > len(pairs) // 2
