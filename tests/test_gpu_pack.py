"""Pack kernel through the C-ABI: decode(pack(x)) == x for ragged lengths and every
4-byte misalignment (tests/test_packing_kernel.cu:33-191,225-412 pin the same property
for the reference's layout; the packed layout itself is internal)."""
import ctypes as C

import pytest

import wfagpu

pytestmark = pytest.mark.gpu

CODE = {"A": 0, "C": 1, "T": 2, "G": 3}
UNPACK = "ACTG"


def unpack(words, off, length):
    out = []
    for i in range(length):
        w = words[off + i // 8]
        out.append(UNPACK[(w >> (30 - 2 * (i % 8))) & 3])
    return "".join(out)


def test_pack_roundtrip_and_flags(lib, oracle):
    a = wfagpu.Aligner()
    a.add_synthetic(5, 40, 300, 0.01)       # the reference KAT shapes: 300/299 bp ...
    a.add_synthetic(6, 60, 150, 0.02)       # ... and 150/151/148 bp
    extra = [("", "A"), ("A", ""), ("ACGTACG", "ACGTACGT"), ("ACGTACGTA", "ACGTACGTACGTACGTA"),
             ("ACGTNACGT", "ACGTACGT"), ("ACGT", "NNNN"), ("acgtacgtac", "ACGTACGTAC"), ("ACGTRY", "ACGTRY"),
             ("A" * 33, "C" * 31), ("G" * 1025, "T" * 1023)]
    for p, t in extra:
        a.add_sequences(p, t)
    n = a.num_pairs
    rb = wfagpu.ResidentBatch(a)
    words_cap = sum(((len(s) + 7) // 8 + 1 + 3) // 4 * 4 for i in range(n) for s in a.pair(i)) + 64
    packed = (C.c_uint32 * words_cap)()
    buf = C.string_at(a.s.sequences_buffer + rb.base, rb.nbytes)
    assert lib.wfagpu_device_pack_only(rb.dev, buf, rb.nbytes, rb.pairs, n, packed, words_cap) == 0
    for i in range(n):
        p, t = a.pair(i)
        flagged = bool(rb.pairs[i].flags & 1)
        assert flagged == (oracle.has_N(p) or oracle.has_N(t)), (p[:20], t[:20])
        if flagged:
            continue
        for s, off in ((p, rb.pairs[i].p_word), (t, rb.pairs[i].t_word)):
            assert off % 4 == 0
            # (c & 6) >> 1: lower case maps like upper case, other letters are silently
            # mis-encoded exactly as in lib/kernels/sequence_packing_kernel.cu:79
            want = "".join(UNPACK[(ord(c) & 6) >> 1] for c in s)
            assert unpack(packed, off, len(s)) == want
            # every word also carries the following 8 bases (8-base stride, 16 bases per word)
            for j in range(0, max(0, len(s) - 16), 8):
                w = packed[off + j // 8]
                got = "".join(UNPACK[(w >> (30 - 2 * b)) & 3] for b in range(16))
                assert got == want[j:j + 16]
            # padding beyond the sequence is zero
            nwords = ((len(s) + 7) // 8 + 1 + 3) // 4 * 4
            last = packed[off + nwords - 1]
            assert last == 0 or len(s) > 8 * (nwords - 1)
