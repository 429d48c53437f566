"""ctypes front-end for the CPU checker (oracle/).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. The product never imports it.

  Oracle      -> oracle/liboracle.so   (wfagpu_oracle.c: restatement of the reference GPU
                                        path; kernel_model.c: CPU model of our kernel design)
  RefCPU      -> oracle/_ref/libref_cpu.so (the UNMODIFIED reference CPU path, see Makefile)
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libref_cpu.so")
REF_GPU_BIN = os.path.join(HERE, "_ref", "gpu", "wfa.affine.gpu")


def build(ref=True):
    """Compile the checker. `ref` additionally builds oracle/_ref when the
    reference tree is present (it is absent on the GPU box: prebuilt files travel)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference/external/WFA"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
        if os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.check_call(["make", "-s", "-C", HERE, "refgpu"])
        if os.path.exists(os.path.join(HERE, "..", "wfa-gpu_b200", "lib", "libwfagpu.so")):
            subprocess.check_call(["make", "-s", "-C", HERE, "dropin"])


def _b(s):
    return s if isinstance(s, bytes) else s.encode()


class KmStep(C.Structure):
    _fields_ = [("n", C.c_int32), ("row_off", C.c_uint32), ("kind", C.c_uint8)]


class Oracle:
    def __init__(self):
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(
            os.path.getmtime(os.path.join(HERE, f)) for f in ("wfagpu_oracle.c", "kernel_model.c")
        ):
            build(ref=False)
        L = self.L = C.CDLL(LIB)
        L.orc_align_cigar.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int] + [C.c_int] * 6 + [
            C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_long)]
        L.orc_align_score.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int] + [C.c_int] * 6 + [
            C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_long)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_has_N.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_cigar_score.restype = C.c_long
        L.orc_cigar_score.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p] + [C.c_int] * 3
        L.orc_decode_ops.restype = C.c_void_p
        L.orc_decode_ops.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                     C.POINTER(C.c_uint8), C.c_int]
        L.orc_align_chain.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int] + [C.c_int] * 6 + [
            C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int]
        L.km_build_steps.argtypes = [C.c_int] * 5 + [C.POINTER(KmStep), C.POINTER(C.c_uint64)]
        L.km_align_pair_ckpt.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int] + [C.c_int] * 3 + [
            C.POINTER(KmStep), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
            C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_long)]
        L.km_align_pair_bandq.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int] + [C.c_int] * 3 + [
            C.POINTER(KmStep)] + [C.c_int] * 4 + [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_int)]
        L.km_align_pair_bandq.restype = C.c_int
        L.km_align_pair.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int] + [C.c_int] * 3 + [
            C.POINTER(KmStep), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
            C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_long)]

    # -- faithful restatement of the reference GPU path -----------------------
    def align(self, pattern, text, x, o, e, max_steps, band=-1, window=0, cigar=True):
        """-> dict(finished, distance, cigar|None, cells). band<=0: exact kernels."""
        p, t = _b(pattern), _b(text)
        fin, dist, cells = C.c_int(), C.c_int(), C.c_long()
        if cigar:
            out = C.c_void_p()
            rc = self.L.orc_align_cigar(p, len(p), t, len(t), x, o, e, max_steps, band, window,
                                        C.byref(fin), C.byref(dist), C.byref(out), C.byref(cells))
            assert rc == 0
            cg = None
            if out.value:
                cg = C.string_at(out.value).decode()
                self.L.orc_free(out)
        else:
            rc = self.L.orc_align_score(p, len(p), t, len(t), x, o, e, max_steps, band, window,
                                        C.byref(fin), C.byref(dist), C.byref(cells))
            assert rc == 0
            cg = None
        return dict(finished=bool(fin.value), distance=dist.value, cigar=cg, cells=cells.value)

    def align_chain(self, pattern, text, x, o, e, max_steps, band=-1, window=0):
        """-> (finished, distance, final_word, [offloaded words newest first])"""
        p, t = _b(pattern), _b(text)
        fin, dist, fw = C.c_int(), C.c_int(), C.c_uint32()
        cap = max_steps // 4 + 64
        words = (C.c_uint32 * cap)()
        n = self.L.orc_align_chain(p, len(p), t, len(t), x, o, e, max_steps, band, window,
                                   C.byref(fin), C.byref(dist), C.byref(fw), words, cap)
        assert n >= 0
        return bool(fin.value), dist.value, fw.value, list(words[:n])

    def has_N(self, seq):
        s = _b(seq)
        return bool(self.L.orc_has_N(s, len(s)))

    def cigar_score(self, pattern, text, cigar, x, o, e):
        p, t = _b(pattern), _b(text)
        return self.L.orc_cigar_score(p, len(p), t, len(t), _b(cigar), x, o, e)

    def decode_ops(self, pattern, text, distance, ops):
        """ops: iterable of 2-bit ops, oldest first."""
        p, t = _b(pattern), _b(text)
        arr = (C.c_uint8 * max(1, len(ops)))(*ops)
        out = self.L.orc_decode_ops(p, len(p), t, len(t), distance, arr, len(ops))
        s = C.string_at(out).decode()
        self.L.orc_free(out)
        return s

    def model_align_ckpt(self, pattern, text, x, o, e, max_steps, period=32, dmax=-1):
        """CPU model of the checkpointed traceback (ring snapshots every `period` scores +
        recomputation on the dependency cone)."""
        p, t = _b(pattern), _b(text)
        max_dist = max_steps * (max(x, o + e) + 2) + 16
        tab = (KmStep * (max_dist + 1))()
        words = C.c_uint64()
        d_end = self.L.km_build_steps(x, o, e, max_steps, max_dist, tab, C.byref(words))
        fin, dist, nops, rec = C.c_int(), C.c_int(), C.c_int(), C.c_long()
        cap = 2 * d_end + 16
        ops = (C.c_uint8 * cap)()
        rc = self.L.km_align_pair_ckpt(p, len(p), t, len(t), x, o, e, tab, d_end, max_steps, period, dmax,
                                       C.byref(fin), C.byref(dist), ops, cap, C.byref(nops), C.byref(rec))
        assert rc == 0, "checkpointed traceback failed"
        cg = None
        if fin.value:
            cg = self.decode_ops(pattern, text, dist.value, list(ops[: nops.value])[::-1])
        return dict(finished=bool(fin.value), distance=dist.value, cigar=cg, recomputed=rec.value)

    def model_align_bandq(self, pattern, text, x, o, e, max_steps, band, window, cigar=True):
        """CPU model of wfa_bandq_kernel's data layout (rows at k - base in whole quads, reads outside the stored part of a
        row are NULL, window records, decisions from the max predicates) + wfa_band_traceback_kernel."""
        p, t = _b(pattern), _b(text)
        max_dist = max_steps * (max(x, o + e) + 2) + 16
        tab = (KmStep * (max_dist + 1))()
        words = C.c_uint64()
        d_end = self.L.km_build_steps(x, o, e, max_steps, max_dist, tab, C.byref(words))
        fin, dist, nops = C.c_int(), C.c_int(), C.c_int()
        cap = 2 * d_end + 16
        ops = (C.c_uint8 * cap)()
        rc = self.L.km_align_pair_bandq(p, len(p), t, len(t), x, o, e, tab, d_end, band, window, int(cigar),
                                        C.byref(fin), C.byref(dist), ops, cap, C.byref(nops))
        assert rc == 0, "banded model failed (row capacity or a decision byte that was never written)"
        cg = None
        if fin.value and cigar:
            cg = self.decode_ops(pattern, text, dist.value, list(ops[: nops.value])[::-1])
        return dict(finished=bool(fin.value), distance=dist.value, cigar=cg)

    # -- CPU model of the B200 kernel's algorithm ------------------------------
    def model_align(self, pattern, text, x, o, e, max_steps, cigar=True, n_cap=None):
        p, t = _b(pattern), _b(text)
        max_dist = max_steps * (max(x, o + e) + 2) + 16
        tab = (KmStep * (max_dist + 1))()
        words = C.c_uint64()
        d_end = self.L.km_build_steps(x, o, e, max_steps, max_dist, tab, C.byref(words))
        if n_cap is None:
            n_cap = max_steps
        fin, dist, nops, cells = C.c_int(), C.c_int(), C.c_int(), C.c_long()
        cap = 2 * d_end + 16
        ops = (C.c_uint8 * cap)()
        rc = self.L.km_align_pair(p, len(p), t, len(t), x, o, e, tab, d_end, n_cap, int(cigar),
                                  C.byref(fin), C.byref(dist), ops, cap, C.byref(nops), C.byref(cells))
        assert rc == 0, "model traceback failed"
        cg = None
        if fin.value and cigar:
            seq = list(ops[: nops.value])[::-1]
            cg = self.decode_ops(pattern, text, dist.value, seq)
        return dict(finished=bool(fin.value), distance=dist.value, cigar=cg, cells=cells.value)


class RefCPU:
    """The unmodified reference CPU path (WFA2-lib v2.3 via utils/wfa_cpu.c)."""

    def __init__(self):
        if not os.path.exists(REF_LIB):
            raise FileNotFoundError(REF_LIB)
        L = self.L = C.CDLL(REF_LIB)
        L.ref_cpu_align_batch.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                          C.POINTER(C.c_char_p), C.POINTER(C.c_int)] + [C.c_int] * 6 + [
                                              C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
        L.ref_recover_cigar.restype = C.c_void_p
        L.ref_recover_cigar.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_uint32,
                                        C.c_int, C.POINTER(C.c_uint32)]
        L.ref_free.argtypes = [C.c_void_p]

    @staticmethod
    def available():
        return os.path.exists(REF_LIB)

    def max_threads(self):
        return self.L.ref_max_threads()

    def align_batch(self, patterns, texts, x, o, e, cigar=True, adaptive=False, threads=0):
        n = len(patterns)
        ps = [_b(s) for s in patterns]
        ts = [_b(s) for s in texts]
        P = (C.c_char_p * n)(*ps)
        T = (C.c_char_p * n)(*ts)
        PL = (C.c_int * n)(*[len(s) for s in ps])
        TL = (C.c_int * n)(*[len(s) for s in ts])
        err = (C.c_int * n)()
        cg = (C.c_void_p * n)()
        done = self.L.ref_cpu_align_batch(n, P, PL, T, TL, x, o, e, int(cigar), int(adaptive), threads, err, cg)
        assert done >= 0
        cigars = None
        if cigar:
            cigars = []
            for i in range(n):
                cigars.append(C.string_at(cg[i]).decode())
                self.L.ref_free(cg[i])
        return list(err), cigars

    def recover_cigar(self, pattern, text, distance, final_word, words):
        p, t = _b(pattern), _b(text)
        arr = (C.c_uint32 * max(1, len(words)))(*words)
        out = self.L.ref_recover_cigar(p, len(p), t, len(t), distance, final_word, len(words), arr)
        s = C.string_at(out).decode()
        self.L.ref_free(out)
        return s
