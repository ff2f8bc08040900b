import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement oracle (test infrastructure), built on demand."""
    import oracle as orc

    orc.build()
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def gdt():
    """The product package; the CUDA library must already be built (no fallback)."""
    import dune_gdt_b200 as gdt

    gdt.capi.lib()
    return gdt


@pytest.fixture(scope="session")
def ctx(gdt):
    return gdt.Context(0)
