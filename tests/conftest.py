import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def refcpu():
    from oracle import RefCPU
    if not RefCPU.available():
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    return RefCPU()


@pytest.fixture(scope="session")
def lib():
    import wfagpu
    return wfagpu.load()
