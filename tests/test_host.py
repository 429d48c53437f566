"""Host-side logic of the product library, no GPU needed: ABI layout, exported symbols,
step table, CIGAR text generation, synthetic generator, API argument checking."""
import ctypes as C
import os
import re

import pytest

import wfagpu
from oracle import KmStep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    declared = set()
    for h in ("wfa_gpu.h", "wfagpu_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"static inline[^{]*\{.*?\n\}", "", src, flags=re.S)
        for m in re.finditer(r"\b([a-z_][a-zA-Z0-9_]*)\s*\([^;{]*\)\s*;", src):
            declared.add(m.group(1))
    declared -= {"defined", "sizeof"}
    assert {"wfagpu_align", "launch_alignments", "wfagpu_device_align", "get_cuda_SM_count"} <= declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/ but not exported"
    for name in wfagpu.EXPORTS:
        assert hasattr(lib, name)


def test_struct_layout_matches_reference_abi():
    # sizes/offsets measured on the reference headers (x86-64)
    assert C.sizeof(wfagpu.AlignerStruct) == 104
    assert wfagpu.AlignerStruct.alignment_options.offset == 56
    assert wfagpu.AlignerStruct.last_sequence_pair_idx.offset == 48
    assert C.sizeof(wfagpu.AlignmentOptions) == 48
    assert wfagpu.AlignmentOptions.batch_size.offset == 16
    assert wfagpu.AlignmentOptions.penalties.offset == 32
    assert wfagpu.AlignmentOptions.compute_cigar.offset == 44
    assert C.sizeof(wfagpu.AlignmentResult) == 32 and wfagpu.AlignmentResult.cigar.offset == 8
    assert C.sizeof(wfagpu.SequencePair) == 48 and wfagpu.SequencePair.has_N.offset == 40
    assert C.sizeof(wfagpu.AffinePenalties) == 12


def test_c_header_layout_with_gcc(tmp_path):
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "include/wfa_gpu.h"\n#include <stddef.h>\n#include <stdio.h>\n'
        "int main(){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(wfagpu_aligner_t),"
        "offsetof(wfagpu_aligner_t, alignment_options), sizeof(wfa_alignment_options_t),"
        "offsetof(wfa_alignment_options_t, compute_cigar), sizeof(wfa_alignment_result_t),"
        "sizeof(sequence_pair_t), sizeof(alignment_result_t), sizeof(wfa_backtrace_t));return 0;}\n")
    exe = tmp_path / "abi"
    assert os.system(f"gcc -I{ROOT} -o {exe} {src}") == 0
    out = os.popen(str(exe)).read().split()
    assert out == ["104", "56", "48", "44", "32", "48", "20", "8"]


def test_api_argument_checks(lib):
    # tests/test_api.c:30-57
    assert not lib.wfagpu_initialize_aligner(None)
    assert not lib.wfagpu_add_sequences(None, b"ACGT", b"ACGT")
    assert not lib.wfagpu_initialize_parameters(None, wfagpu.AffinePenalties(2, 3, 1))
    a = wfagpu.Aligner()
    assert a.add_sequences("ACGT", "ACGA")
    assert not a.initialize_parameters(-2, 3, 1)
    assert not a.initialize_parameters(0, 0, 0)
    assert a.initialize_parameters(2, 3, 1)
    assert a.add_sequences("A" * 32767, "ACGT")          # the reference's limit (lib/aligner.c:139-142) ...
    assert a.add_sequences("A" * 40000, "ACGT")          # ... is lifted: long pairs run on the int32 tier
    assert not a.add_sequences("A" * (1 << 22), "ACGT")


def test_buffer_layout_and_defaults():
    a = wfagpu.Aligner()
    seqs = [("ACG", "ACGT"), ("ACGTA", "AC"), ("", "A"), ("ACGTACGT", "ACGTACGA")]
    for p, t in seqs:
        assert a.add_sequences(p, t)
    off = 0
    for i, (p, t) in enumerate(seqs):
        m = a.s.sequences_metadata[i]
        assert m.pattern_offset == off and m.pattern_len == len(p)
        off = (off + len(p) + 1) + (4 - (off + len(p) + 1) % 4)
        assert m.text_offset == off and m.text_len == len(t)
        off = (off + len(t) + 1) + (4 - (off + len(t) + 1) % 4)
        assert a.pair(i) == (p, t)
    b = wfagpu.Aligner()
    b.add_synthetic(1, 25, 10000, 0.05)
    assert b.initialize_parameters(2, 3, 1)
    o = b.options
    # lib/alignment_parameters.h:83-106: 0.1 * max(len) * max(x,o,e), tpb from the wavefront width
    assert (o.max_error, o.threads_per_block, o.band, o.batch_size, o.compute_cigar) == (3000, 1024, -1, 2, False)
    assert b.set_batch_size(0) and o.batch_size == 25
    assert b.set_batch_size(99) and o.batch_size == 25


@pytest.mark.parametrize("pen", [(2, 3, 1), (5, 3, 2), (4, 6, 2), (1, 2, 1), (3, 1, 4), (2, 10, 5)])
def test_step_table_equals_model(lib, oracle, pen):
    x, o, e = pen
    ms = 300
    md = ms * (max(x, o + e) + 1) + 16
    t1 = (wfagpu.Step * (md + 1))(); u1 = C.c_uint64()
    d1 = lib.wfagpu_build_step_table(x, o, e, ms, md, 0, t1, C.byref(u1))
    t2 = (KmStep * (md + 1))(); u2 = C.c_uint64()
    d2 = oracle.L.km_build_steps(x, o, e, ms, md, t2, C.byref(u2))
    assert d1 == d2 and u1.value * 16 == u2.value
    for d in range(d1):
        assert (t1[d].kind, t1[d].n) == (t2[d].kind, t2[d].n)
        if t1[d].kind == 2:
            assert t1[d].row_off * 16 == t2[d].row_off


def _model_ops(oracle, p, t, pen, budget):
    x, o, e = pen
    pb, tb = p.encode(), t.encode()
    md = budget * (max(x, o + e) + 2) + 16
    tab = (KmStep * (md + 1))(); w = C.c_uint64()
    dend = oracle.L.km_build_steps(x, o, e, budget, md, tab, C.byref(w))
    fin, dist, nops, cells = C.c_int(), C.c_int(), C.c_int(), C.c_long()
    cap = 2 * dend + 16
    ops = (C.c_uint8 * cap)()
    assert oracle.L.km_align_pair(pb, len(pb), tb, len(tb), x, o, e, tab, dend, budget, 1, C.byref(fin),
                                  C.byref(dist), ops, cap, C.byref(nops), C.byref(cells)) == 0
    return fin.value, dist.value, list(ops[: nops.value])


@pytest.mark.parametrize("pen", [(2, 3, 1), (5, 3, 2)])
def test_ops_to_cigar_equals_oracle_decoder(lib, oracle, pen):
    a = wfagpu.Aligner()
    a.add_synthetic(0xB2000001, 150, 150, 0.05)
    a.add_synthetic(0xB2000002, 8, 1000, 0.10)
    a.add_sequences("ACGT", "ACGT")
    for i in range(a.num_pairs):
        p, t = a.pair(i)
        fin, dist, ops = _model_ops(oracle, p, t, pen, 1000)
        assert fin
        n = len(ops)
        words = (C.c_uint32 * ((n + 15) // 16 + 1))()
        for j, op in enumerate(ops):
            words[j >> 4] |= op << (2 * (j & 15))
        cg = wfagpu.Cigar()
        assert lib.wfagpu_ops_to_cigar(p.encode(), len(p), t.encode(), len(t), dist, words, n, C.byref(cg))
        text = C.string_at(cg.buffer).decode() if cg.buffer else ""
        assert text == oracle.align(p, t, *pen, 1000)["cigar"]


def test_synthetic_generator_is_deterministic_and_has_the_asked_shape():
    a, b = wfagpu.Aligner(), wfagpu.Aligner()
    a.add_synthetic(42, 50, 1000, 0.10)
    b.add_synthetic(42, 20, 1000, 0.10)
    for i in range(20):
        assert a.pair(i) == b.pair(i)
    for i in range(50):
        p, t = a.pair(i)
        assert len(t) == 1000 and abs(len(p) - 1000) <= 100 and set(p + t) <= set("ACGT")
    assert a.pair(0) != a.pair(1)


def test_no_cpu_fallback_without_gpu(lib):
    # on a machine without a CUDA device the library must fail loudly, never compute on the CPU
    n = C.c_int(0)
    lib.get_num_cuda_devices(C.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    a = wfagpu.Aligner()
    a.add_sequences("ACGT", "ACGA")
    assert a.initialize_parameters(2, 3, 1)
    with pytest.raises(RuntimeError):
        a.align()


def test_readers_seq_and_fasta(tmp_path, lib):
    # utils/sequence_reader.c:137-392: .seq (blank lines skipped) and paired multi-line FASTA
    seq = tmp_path / "a.seq"
    seq.write_text(">ACGT\n<ACGA\n\n>TTTTT\n<TTT\n>GATTACA\n<GATTACA")       # no trailing newline on purpose
    a = wfagpu.Aligner()
    assert a.read_seq_file(str(seq)) == 3
    assert [a.pair(i) for i in range(3)] == [("ACGT", "ACGA"), ("TTTTT", "TTT"), ("GATTACA", "GATTACA")]
    b = wfagpu.Aligner()
    assert b.read_seq_file(str(seq), 2) == 2 and b.num_pairs == 2
    q = tmp_path / "q.fasta"
    t = tmp_path / "t.fasta"
    q.write_text(">q1 desc\nACGT\nACGT\n\n>q2\nTTTT\n")
    t.write_text(" >t1\nACGTAC\nGT\n>t2\nTT\nTA\n")
    c = wfagpu.Aligner()
    assert c.read_fasta_files(str(q), str(t)) == 2
    assert [c.pair(i) for i in range(2)] == [("ACGTACGT", "ACGTACGT"), ("TTTT", "TTTA")]
    bad = tmp_path / "bad.seq"
    bad.write_text("ACGT\n<ACGT\n")
    d = wfagpu.Aligner()
    assert d.read_seq_file(str(bad)) == -1


def test_check_result(lib, oracle):
    pen = wfagpu.AffinePenalties(2, 3, 1)
    p, t = b"ACGTACGTAC", b"ACGTTCGAC"
    r = oracle.align(p.decode(), t.decode(), 2, 3, 1, 100)
    assert lib.wfagpu_check_result(p, len(p), t, len(t), pen, r["distance"], r["cigar"].encode())
    assert not lib.wfagpu_check_result(p, len(p), t, len(t), pen, r["distance"] + 1, r["cigar"].encode())
    assert not lib.wfagpu_check_result(p, len(p), t, len(t), pen, 0, b"9M")
    assert not lib.wfagpu_check_result(p, len(p), t, len(t), pen, 0, b"garbage")


def test_reference_validator_names(lib, oracle):
    # check_cigar_edit / check_affine_distance: the reference library's generic validators, exported under
    # the same names with the same argument order (text first; utils/verification.h:37-49)
    import ctypes as C
    lib.check_cigar_edit.restype = C.c_bool
    lib.check_cigar_edit.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p]
    lib.check_affine_distance.restype = C.c_bool
    lib.check_affine_distance.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, wfagpu.AffinePenalties,
                                          C.c_char_p]
    pen = wfagpu.AffinePenalties(2, 3, 1)
    p, t = b"ACGTACGTAC", b"ACGTTCGAC"
    r = oracle.align(p.decode(), t.decode(), 2, 3, 1, 100)
    cg = r["cigar"].encode()
    assert lib.check_cigar_edit(t, p, len(t), len(p), cg)
    assert not lib.check_cigar_edit(p, t, len(p), len(t), cg)            # text / pattern swapped: I and D trade places
    assert not lib.check_cigar_edit(t, p, len(t), len(p), b"9M")
    assert lib.check_affine_distance(t, p, len(t), len(p), r["distance"], pen, cg)
    assert not lib.check_affine_distance(t, p, len(t), len(p), r["distance"] + 1, pen, cg)


def test_validators_take_the_reference_unrolled_format(lib, oracle):
    # the reference calls check_cigar_edit / check_affine_distance with the UNROLLED op string of recover_cigar
    # (lib/align.cu:284-293, utils/verification.c:27-146); the run-length text of results[i].cigar is accepted too
    import ctypes as C
    pen = wfagpu.AffinePenalties(2, 3, 1)
    p, t = "GATTACAGATTACAGGATCCA", "GATTACGATTTACAGGTCCA"
    r = oracle.align(p, t, 2, 3, 1, 100)
    rle = r["cigar"].encode()
    ptr = lib.wfagpu_unroll_cigar(rle)
    unrolled = C.string_at(ptr)
    assert re.fullmatch(rb"[MXID]+", unrolled) and len(unrolled) >= max(len(p), len(t))
    pb, tb = p.encode(), t.encode()
    for cg in (rle, unrolled):
        assert lib.check_cigar_edit(tb, pb, len(tb), len(pb), cg)
        assert lib.check_affine_distance(tb, pb, len(tb), len(pb), r["distance"], pen, cg)
        assert not lib.check_affine_distance(tb, pb, len(tb), len(pb), r["distance"] + 1, pen, cg)
    assert not lib.check_cigar_edit(tb, pb, len(tb), len(pb), unrolled[:-1])
    assert not lib.check_cigar_edit(tb, pb, len(tb), len(pb), unrolled.replace(b"X", b"M", 1))
    assert lib.wfagpu_unroll_cigar(b"3M1") is None and lib.wfagpu_unroll_cigar(b"3M0X") is None


def test_recover_cigar_decodes_reference_format_chains(lib, oracle):
    # recover_cigar (utils/verification.h:52-58) on backtrace chains in the reference's own layout (produced by the
    # restatement of its kernels): the unrolled string must be the unrolled form of the reference decoder's text
    import ctypes as C

    class Bt(C.Structure):
        _fields_ = [("backtrace", C.c_uint32), ("prev", C.c_uint32)]

    class Res(C.Structure):
        _fields_ = [("finished", C.c_bool), ("distance", C.c_int), ("bt", Bt), ("num_bt_blocks", C.c_int)]

    lib.recover_cigar.restype = C.c_void_p
    lib.recover_cigar.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_size_t, Bt, C.POINTER(Bt), Res]
    a = wfagpu.Aligner()
    a.add_synthetic(0xB2007700, 6, 400, 0.08, 0.12)
    a.add_synthetic(0xB2007701, 6, 60, 0.0, 0.3)
    for i in range(a.num_pairs):
        p, t = a.pair(i)
        fin, dist, final_word, words = oracle.align_chain(p, t, 2, 3, 1, 400)
        assert fin
        arr = (Bt * max(1, len(words)))(*[Bt(w, 0) for w in words])
        res = Res(True, dist, Bt(final_word, 0), len(words))
        got = C.string_at(lib.recover_cigar(t.encode(), p.encode(), len(t), len(p), Bt(final_word, 0), arr, res))
        want = C.string_at(lib.wfagpu_unroll_cigar(oracle.align(p, t, 2, 3, 1, 400)["cigar"].encode()))
        assert got == want, i


def test_device_list_syntax(lib):
    import ctypes as C
    devs = (C.c_int * 16)()
    f = lib.wfagpu_parse_devices
    assert f(None, 4, devs, 16) == 1 and devs[0] == 0
    assert f(b"all", 4, devs, 16) == 4 and list(devs[:4]) == [0, 1, 2, 3]
    assert f(b"n:2", 4, devs, 16) == 2 and list(devs[:2]) == [0, 1]
    assert f(b"n:8", 4, devs, 16) == 4
    assert f(b"0,2,3", 4, devs, 16) == 3 and list(devs[:3]) == [0, 2, 3]
    assert f(b"0,0", 1, devs, 16) == 2 and list(devs[:2]) == [0, 0]        # two workers, own contexts, one GPU
    assert f(b"0,7", 4, devs, 16) == -1                                      # not a visible device
    assert f(b"0;1", 4, devs, 16) == -1 and f(b"x", 4, devs, 16) == -1


def test_chunk_plan(lib):
    import ctypes as C
    chunk, n = C.c_size_t(), C.c_size_t()
    f = lambda *a: (lib.wfagpu_plan_chunks(*a, C.byref(chunk), C.byref(n)), (chunk.value, n.value))[1]
    assert f(100, 100, 1, 10**6) == (100, 1)                                 # small call: one chunk
    assert f(8192, 8192, 1, 8192 * 20000) == (4096, 2)                       # default batch size: two chunks per worker
    assert f(8192, 1000, 1, 8192 * 20000) == (1000, 9)                       # the caller's batch size is an upper bound
    assert f(8192, 8192, 2, 8192 * 20000) == (2048, 4)
    assert f(100000, 100000, 1, 100000 * 20000) == (12500, 8)                # long streams: eight chunks per worker
    assert f(3000, 3000, 2, 3000 * 300) == (750, 4)                          # several GPUs: at least two chunks each
    assert f(4, 4, 8, 4000)[0] >= 1
    c, k = f(10**6, 10**6, 1, 10**6 * 20000)                                 # a chunk's ASCII stays below 3 GiB
    assert c * 20000 <= 3 * 2**30 and c * k >= 10**6


def test_shares_of_a_multi_worker_stream(lib):
    # in-library multi-GPU: every worker walks its own contiguous share (first chunk, chunks, remainder); a worker whose share
    # is used up steals half chunks from the back of the fullest one.  Every pair is handed out exactly once.
    import ctypes as C
    def deal(n, workers, chunk, first, order):
        nxt, end = (C.c_size_t * workers)(), (C.c_size_t * workers)()
        lib.wfagpu_plan_shares(n, workers, nxt, end)
        assert nxt[0] == 0 and end[workers - 1] == n and all(end[i] == nxt[i + 1] for i in range(workers - 1))
        got, taken_first = [[] for _ in range(workers)], [False] * workers
        frm, cnt = C.c_size_t(), C.c_size_t()
        for w in order:
            want = chunk if taken_first[w] else first
            if lib.wfagpu_share_take(nxt, end, workers, w, want, chunk // 2, C.byref(frm), C.byref(cnt)):
                got[w].append((frm.value, cnt.value))
                taken_first[w] = True
        covered = sorted(x for g in got for x in g)
        at = 0
        for f0, c0 in covered:
            assert f0 == at and c0 > 0
            at += c0
        assert at == n
        return got
    # eight equal GPUs, 8 x 8192 pairs, chunks of 4096 with a 2048-pair first chunk: 2048 + 4096 + 2048 each, nothing stolen
    got = deal(65536, 8, 4096, 2048, [w for _ in range(4) for w in range(8)])
    assert all([c for _, c in g] == [2048, 4096, 2048] for g in got)
    assert all(g[0][0] == 8192 * w for w, g in enumerate(got))
    # one worker is four times faster than the other: it finishes its share and helps with half chunks from the back
    got = deal(16384, 2, 4096, 4096, [0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0])
    assert sum(c for _, c in got[0]) > sum(c for _, c in got[1]) > 0
    assert any(f >= 8192 for f, _ in got[0])
    # ragged: more workers than pairs, and a share that is not a multiple of the chunk
    deal(5, 8, 4, 4, list(range(8)) * 2)
    deal(10001, 3, 1024, 512, [0, 1, 2] * 8)
    assert not lib.wfagpu_share_take((C.c_size_t * 2)(4, 9), (C.c_size_t * 2)(4, 9), 2, 5, 1, 1, C.byref(C.c_size_t()), C.byref(C.c_size_t()))


def test_metadata_must_be_word_aligned(lib):
    import ctypes as C
    a = wfagpu.Aligner()
    a.add_sequences("ACGTACGTAC", "ACGTTCGAC")
    a.add_sequences("ACGT", "ACGA")
    pairs = (wfagpu.DevPair * 2)()
    base, nbytes = C.c_size_t(), C.c_size_t()
    assert lib.wfagpu_pairs_from_metadata(a.s.sequences_metadata, 0, 2, a.s.sequences_buffer_len, pairs, C.byref(base), C.byref(nbytes)) == 0
    a.s.sequences_metadata[1].pattern_offset += 1                            # the pack kernel selects 32-bit words
    assert lib.wfagpu_pairs_from_metadata(a.s.sequences_metadata, 0, 2, a.s.sequences_buffer_len, pairs, C.byref(base), C.byref(nbytes)) != 0
    a.s.sequences_metadata[1].pattern_offset -= 1


def test_sequences_longer_than_the_large_tier_are_rejected(lib):
    a = wfagpu.Aligner()
    assert not a.add_sequences("A" * 220000, "ACGT")
    assert a.add_sequences("A" * 1000, "ACGT")


def hifi_fixture(tmp_path):
    """The reference's FASTA fixture (tests/data/test_hifi.*.fasta), unpacked from tests/golden."""
    import gzip, json, shutil
    out = []
    for side in ("query", "target"):
        dst = tmp_path / f"test_hifi.{side}.fasta"
        with gzip.open(os.path.join(ROOT, "tests", "golden", f"test_hifi.{side}.fasta.gz"), "rb") as f, open(dst, "wb") as g:
            shutil.copyfileobj(f, g)
        out.append(str(dst))
    return out[0], out[1], json.load(open(os.path.join(ROOT, "tests", "golden", "test_hifi.json")))


def test_hifi_fasta_fixture_reads_like_the_reference(lib, tmp_path):
    # the reference's FASTA fixture (tests/test-fasta.sh:11-22: 50 pairs, `correct=50`), committed as a golden copy
    q, t, gold = hifi_fixture(tmp_path)
    a = wfagpu.Aligner()
    assert a.read_fasta_files(q, t) == 50
    lens = [[a.s.sequences_metadata[i].pattern_len, a.s.sequences_metadata[i].text_len] for i in range(50)]
    assert lens == gold["lengths"]
    for i in (0, 17, 49):
        p, tt = a.pair(i)
        assert set(p) <= set("ACGTN") and set(tt) <= set("ACGTN")
