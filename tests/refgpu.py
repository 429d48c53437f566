"""Runs the UNMODIFIED reference GPU binary (oracle/_ref/gpu/wfa.affine.gpu, built by
`make -C oracle refgpu` from /root/reference for sm_100) on a set of pairs.  Test
infrastructure; only usable on a box with a GPU."""
import os
import subprocess
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "gpu", "wfa.affine.gpu")


def available():
    return os.path.exists(BIN)


def run(pairs, pen, max_error, cigar=True, band=None, threads=None, batch=None, workers=None):
    """-> (list of (score, cigar|None), wall seconds reported by the tool, total seconds)"""
    with tempfile.TemporaryDirectory() as td:
        seq = os.path.join(td, "in.seq")
        out = os.path.join(td, "out.txt")
        with open(seq, "w") as f:
            for p, t in pairs:
                f.write(">" + p + "\n<" + t + "\n")
        cmd = [BIN, "-i", seq, "-g", "%d,%d,%d" % pen, "-e", str(max_error), "-o", out]
        if cigar:
            cmd.append("-x")
        if band is not None:
            cmd += ["-B", str(band)]
        if threads is not None:
            cmd += ["-t", str(threads)]
        if batch is not None:
            cmd += ["-b", str(batch)]
        if workers is not None:
            cmd += ["-w", str(workers)]
        t0 = time.time()
        pr = subprocess.run(cmd, capture_output=True, text=True)
        total = time.time() - t0
        if pr.returncode != 0:
            raise RuntimeError("reference GPU binary failed: " + pr.stderr[-2000:])
        wall = None
        for line in (pr.stdout + pr.stderr).splitlines():
            if "Wall time" in line:
                try:
                    wall = float(line.split("Wall time:")[1].split("s")[0])
                except Exception:
                    pass
        res = []
        for line in open(out):
            parts = line.rstrip("\n").split("\t")
            if not parts or parts[0] == "":
                continue
            res.append((-int(parts[0]), parts[1] if len(parts) > 1 else None))
        return res, wall, total
