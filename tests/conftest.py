import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    import harness
    return harness.oracle()


@pytest.fixture(scope="session")
def ref():
    import harness
    if not harness.have_ref():
        pytest.skip("oracle/_ref/libtntref.so not built (needs /root/reference; see oracle/Makefile)")
    return harness.ref()


@pytest.fixture(scope="session")
def engine_lib():
    """The CUDA shared library; building is part of __graft_entry__.build()."""
    from thermonucleotideblast_b200 import build as b
    b.build()
    from thermonucleotideblast_b200.engine import load_library
    return load_library()
