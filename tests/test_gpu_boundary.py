"""The drop-in boundary on the GPU: what an unmodified reference caller gets (page-locked aligner buffer, staged
pageable buffers), several workers / host threads on one GPU, check_correctness, per-pair failures, the FASTA fixture."""
import ctypes as C
import os
import subprocess
import sys
import threading

import pytest

import wfagpu
from util import synth_aligner, check_against_oracle
from test_host import hifi_fixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(specs, pen, cigar, max_error, batch=None, seed=0xB2003000):
    a = synth_aligner(specs, seed)
    assert a.initialize_parameters(*pen)
    a.options.max_error = max_error
    a.options.compute_cigar = cigar
    if batch:
        a.set_batch_size(batch)
    a.align()
    return a


def test_aligner_buffer_is_page_locked_without_any_extension_call():
    # the reference's TODO (utils/sequence_reader.c:73): wfagpu_initialize_aligner hands out pinned memory, also
    # after the buffer grew, so an unmodified caller uploads by asynchronous DMA
    a = synth_aligner([(600, 1000, 0.05, 0.05)])
    assert a.s.sequences_buffer_len > (1 << 20)
    assert a.host_buffer_pinned()
    assert a.initialize_parameters(2, 3, 1)
    a.options.compute_cigar = True
    a.align()
    assert a.run_stats()["staged"] == 0


def test_pageable_caller_buffer_is_staged_and_gives_the_same_results(lib):
    # launch_alignments with a buffer the caller malloc'ed itself (what the reference's readers do): staged chunk by
    # chunk through the slots' page-locked buffers; byte-identical results
    specs = [(1500, 300, 0.05, 0.05), (120, 3000, 0.03, 0.06)]
    a = run(specs, (2, 3, 1), True, 600, batch=300)
    b = synth_aligner(specs, 0xB2003000)
    assert b.initialize_parameters(2, 3, 1)
    b.options.max_error = 600
    b.options.compute_cigar = True
    b.set_batch_size(300)
    n = b.s.sequences_buffer_len
    copy = C.create_string_buffer(n)                         # pageable
    C.memmove(copy, b.s.sequences_buffer, n)
    assert not lib.wfagpu_host_is_pinned(copy)
    lib.launch_alignments(copy, n, b.s.sequences_metadata, b.s.results, b.s.alignment_options, False)
    st = b.run_stats()
    assert st["staged"] == 1 and st["failed_pairs"] == 0
    assert a.errors() == b.errors() and a.cigars() == b.cigars()


def test_two_workers_on_one_gpu_match_one_worker(oracle):
    # driver.c's multi-worker path (shared job, chunk hand-out, one leased context per worker) on a single GPU:
    # the device list "0,0" starts two host threads with their own streams, slots and staging buffers
    specs = [(3000, 300, 0.05, 0.05), (200, 3000, 0.05, 0.05)]
    wfagpu.set_devices("0")
    a = run(specs, (2, 3, 1), True, 600, batch=400)
    wfagpu.set_devices("0,0")
    try:
        b = run(specs, (2, 3, 1), True, 600, batch=400)
        st = b.run_stats()
        c = run(specs, (2, 3, 1), False, 600, batch=400)
    finally:
        wfagpu.set_devices("0")
    assert st["devices"] == 2
    assert a.errors() == b.errors() == c.errors()
    assert a.cigars() == b.cigars()
    assert check_against_oracle(oracle, b, 2, 3, 1, 600, True, sample=list(range(0, b.num_pairs, 97))) == []


def test_two_host_threads_with_their_own_aligners_share_a_gpu(oracle):
    # the reference keeps no global state (lib/align.cu:63-162); here every call leases its own device context
    specs = [[(1200, 300, 0.05, 0.05), (60, 2500, 0.05, 0.05)], [(900, 500, 0.02, 0.1), (40, 4000, 0.03, 0.05)]]
    want = [run(s, (2, 3, 1), True, 500, batch=256, seed=0xB2003100 + i) for i, s in enumerate(specs)]
    got, errs = [None, None], []

    def work(i):
        try:
            for _ in range(3):
                got[i] = run(specs[i], (2, 3, 1), True, 500, batch=256, seed=0xB2003100 + i)
        except Exception as e:                                # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert errs == []
    for i in range(2):
        assert got[i].errors() == want[i].errors() and got[i].cigars() == want[i].cigars()


@pytest.mark.parametrize("cigar", [True, False])
def test_check_correctness_validates_against_an_independent_gpu_pass(cigar):
    # launch_alignments(..., check_correctness = true): CIGAR validated on the host, score compared with the
    # one-diagonal-per-thread kernels without bounds / hints (the reference compares with its CPU WFA, lib/align.cu:300)
    a = synth_aligner([(800, 150, 0.05, 0.05), (300, 1000, 0.1, 0.1), (24, 6000, 0.05, 0.05)], 0xB2003200)
    assert a.initialize_parameters(2, 3, 1)
    a.options.max_error = 300                                 # some 1 kbp pairs exceed it: re-dispatched, then checked
    a.options.compute_cigar = cigar
    a.set_batch_size(500)
    checked, incorrect = a.align_checked()
    assert (checked, incorrect) == (a.num_pairs, 0)


def test_rescore_is_an_independent_path_with_the_same_scores(lib):
    a = synth_aligner([(64, 2000, 0.05, 0.1)], 0xB2003300)
    assert a.initialize_parameters(2, 3, 1)
    a.options.max_error = 800
    a.options.compute_cigar = True
    rb = wfagpu.ResidentBatch(a)
    rb.upload()
    rb.align()
    out, _, _ = rb.download()
    st = rb.stats()
    ref = rb.rescore()
    assert ref == [out[i].distance for i in range(rb.n)]
    assert rb.stats()["launches"] < st["launches"]            # no bound / traceback / text kernels in the check pass
    rb.release()


def test_a_pair_the_gpu_cannot_finish_fails_alone():
    # one impossible pair must not abort the call.  The step cap (60000 wavefront steps) is lowered through
    # the test hook so that two unrelated sequences exceed it.
    code = r'''
import sys, random
sys.path.insert(0, %r); sys.path.insert(0, %r)
import wfagpu
from util import synth_aligner
random.seed(7)
a = synth_aligner([(40, 400, 0.05, 0.05)], 0xB2003400)
a.add_sequences("".join(random.choice("ACGT") for _ in range(1500)), "".join(random.choice("ACGT") for _ in range(1500)))
a.add_synthetic(0xB2003401, 40, 400, 0.05, 0.05)
assert a.initialize_parameters(2, 3, 1)
a.options.max_error = 100
a.options.compute_cigar = True
ok = True
try:
    a.align()
except RuntimeError:
    ok = False
st = a.run_stats()
b = synth_aligner([(40, 400, 0.05, 0.05)], 0xB2003400)
b.add_synthetic(0xB2003401, 40, 400, 0.05, 0.05)
assert b.initialize_parameters(2, 3, 1); b.options.max_error = 100; b.options.compute_cigar = True; b.align()
good = [i for i in range(81) if i != 40]
assert not ok and st["failed_pairs"] == 1, (ok, st)
assert a.error(40) == 0xffffffff and a.cigar(40) == ""
assert [a.error(i) for i in good] == b.errors() and [a.cigar(i) for i in good] == b.cigars()
print("OK")
''' % (os.path.join(ROOT, "wfa-gpu_b200", "python"), os.path.join(ROOT, "tests"))
    pr = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, WFAGPU_MAX_STEPS_CAP="500"),
                        capture_output=True, text=True, timeout=600)
    assert pr.returncode == 0 and "OK" in pr.stdout, pr.stdout[-2000:] + pr.stderr[-3000:]
    assert "could not be aligned on the GPU" in pr.stderr


def test_penalties_are_validated_at_the_device_boundary(lib):
    a = synth_aligner([(8, 100, 0.05, 0.05)])
    assert a.initialize_parameters(2, 3, 1)
    rb = wfagpu.ResidentBatch(a)
    rb.upload()
    for x, o, e in ((0, 3, 1), (2, 3, 0), (2, -1, 1)):
        plan = rb.plan()
        plan.x, plan.o, plan.e = x, o, e
        assert lib.wfagpu_device_align(rb.dev, rb.slot, rb.n, C.byref(plan), 1) != 0
    rb.release()


def test_hifi_fasta_fixture_scores_and_check(tmp_path):
    # the reference's FASTA fixture end to end through the CLI (tests/test-fasta.sh:11-22: `correct=50`), scores ==
    # the unmodified reference CPU WFA (tests/golden/test_hifi.json), both penalty sets of the reference's test
    q, t, gold = hifi_fixture(tmp_path)
    exe = os.path.join(ROOT, "bin", "wfa.affine.gpu")
    for pen, extra in (((2, 3, 1), []), ((5, 2, 5), ["-b", "11"])):
        out = tmp_path / ("out_%d.txt" % pen[0])
        pr = subprocess.run([exe, "-Q", q, "-T", t, "-g", "%d,%d,%d" % pen, "-x", "-c", "-o", str(out)] + extra,
                            capture_output=True, text=True, timeout=600)
        assert pr.returncode == 0, pr.stderr[-2000:]
        assert "correct=50 Incorrect=0" in pr.stderr
        scores = [-int(line.split("\t")[0]) for line in open(out)]
        assert scores == gold["scores"]["%d,%d,%d" % pen]
