import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import _oracle
    return _oracle.load()


@pytest.fixture(scope="session")
def tagpu():
    """One GPU context shared by the gpu-marked tests; fails loudly if libtagpu.so or the GPU is missing."""
    from turingassembler_b200 import Tagpu
    t = Tagpu()
    yield t
    t.close()
