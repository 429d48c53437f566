"""Drop-in evidence: the reference's OWN tests/test_api.c (golden scores, single/multi batch,
CIGAR on/off) and examples/, compiled UNCHANGED against include/wfa_gpu.h and linked to our
libwfagpu.so (oracle/Makefile target `dropin`), run on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "oracle", "_ref", "dropin")


def run(name, timeout=600):
    exe = os.path.join(D, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs the reference tree)")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "wfa-gpu_b200", "lib"))
    return subprocess.run([exe], capture_output=True, text=True, timeout=timeout, env=env)


def test_reference_test_api_passes_on_our_library():
    pr = run("test_api")
    assert pr.returncode == 0, pr.stderr[-2000:]
    assert "FAILED" not in pr.stderr
    assert "OK" in pr.stderr


@pytest.mark.parametrize("exe", ["manual_example", "auto_example", "auto_example_cpp"])
def test_reference_examples_run_on_our_library(exe):
    pr = run(exe)
    assert pr.returncode == 0, pr.stderr[-2000:]
    assert "Score" in pr.stdout or "score" in pr.stdout.lower()
