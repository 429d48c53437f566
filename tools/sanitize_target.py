#!/usr/bin/env python3
"""Small workloads that touch every kernel family, for compute-sanitizer (memcheck / racecheck):

  compute-sanitizer --tool racecheck python tools/sanitize_target.py [family ...]

families: warp (150 bp CIGAR, warp per pair), cta (1 kbp CIGAR: bound + four-diagonals-per-thread CTA kernel + ring
snapshots + traceback + CIGAR text), quad_pairs (same with two scores per barrier), one_diag (the one-diagonal-per-
thread kernel), workers (two host threads, one GPU), score (1 kbp score only), banded (-B 10, W=128, CIGAR), large
(WFAGPU_FORCE_LARGE=1: rings in global memory), ascii (pairs with N: byte-compare kernels),
redispatch (budget too small on purpose).  Every result is checked against the CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

FAMILIES = {
    #           pairs, length, err, max_error, cigar, band, width, env
    "warp":       (192, 150, 0.04, 50, True, -1, 0, {}),
    "cta":        (40, 1000, 0.10, 400, True, -1, 0, {"WFAGPU_QUAD_MIN": "1"}),
    "score":      (40, 1000, 0.10, 400, False, -1, 0, {"WFAGPU_QUAD_MIN": "1"}),
    "banded":     (24, 2000, 0.05, 400, True, 10, 128, {}),                            # wfa_bandq_kernel + wfa_band_traceback_kernel
    "banded_one": (24, 2000, 0.05, 400, True, 10, 128, {"WFAGPU_NO_QUAD": "1"}),       # wfa_banded_kernel (one diagonal per thread)
    "banded_fused": (24, 2000, 0.05, 400, True, 10, 128, {"WFAGPU_NO_BAND_TB": "1"}),  # wfa_bandq_kernel, backtrace inside
    "large":      (12, 600, 0.08, 200, True, -1, 0, {"WFAGPU_FORCE_LARGE": "1"}),
    "ascii":      (24, 400, 0.05, 100, True, -1, 0, {}),
    "redispatch": (32, 800, 0.10, 40, True, -1, 0, {}),
    "prebound":   (48, 800, 0.10, 230, True, -1, 0, {"WFAGPU_FORCE_BOUND": "1", "WFAGPU_QUAD_MIN": "1"}),   # bounds first: bound-ordered queue, unbounded pairs skip the first pass
    "quad_pairs": (40, 1000, 0.10, 400, True, -1, 0, {"WFAGPU_QUAD_PAIRS": "1", "WFAGPU_QUAD_MIN": "1"}),      # two scores per barrier interval
    "one_diag":   (24, 1000, 0.10, 400, True, -1, 0, {"WFAGPU_NO_QUAD": "1"}),         # one diagonal per thread (the -c path)
    "workers":    (96, 600, 0.06, 200, True, -1, 0, {"WFAGPU_DEVICES": "0,0"}),        # two workers, one GPU
}


def run(name):
    import wfagpu
    from oracle import Oracle
    n, L, err, me, cigar, band, width, env = FAMILIES[name]
    os.environ.update(env)
    a = wfagpu.Aligner()
    a.add_synthetic(0xB2005A00 + len(name), n, L, err, err)
    if name == "ascii":
        for i in range(6):
            p, t = a.pair(i)
            a.add_sequences(p[:50] + "N" + p[51:], t[:70] + "NN" + t[72:])
    assert a.initialize_parameters(2, 3, 1)
    a.options.max_error = me
    a.options.compute_cigar = cigar
    if band > 0:
        a.options.band = band
        a.options.threads_per_block = width
    if name == "workers":
        a.set_batch_size(24)
    a.align()
    orc = Oracle()
    bad = 0
    for i in range(a.num_pairs):
        p, t = a.pair(i)
        if "N" in p or "N" in t:
            # byte-compare pairs: the packed oracle cannot take them; the CIGAR must be a valid alignment of its cost
            pen = wfagpu.AffinePenalties(2, 3, 1)
            ok = a.L.wfagpu_check_result(p.encode(), len(p), t.encode(), len(t), pen, a.error(i), a.cigar(i).encode())
            bad += 0 if ok else 1
            continue
        if band > 0:
            r = orc.align(p, t, 2, 3, 1, me, band=band, window=width)
            if not r["finished"]:
                continue
        else:
            r = orc.align(p, t, 2, 3, 1, 100000)
        if a.error(i) != r["distance"] or (cigar and a.cigar(i) != r["cigar"]):
            bad += 1
    for k in env:
        os.environ.pop(k, None)
    wfagpu.load().wfagpu_device_close_all()
    print(f"sanitize_target {name}: {a.num_pairs} pairs, mismatches vs oracle: {bad}", flush=True)
    return bad


if __name__ == "__main__":
    names = sys.argv[1:] or list(FAMILIES)
    sys.exit(1 if sum(run(n) for n in names) else 0)
