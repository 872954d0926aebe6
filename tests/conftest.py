import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built libraries (they are git-ignored): build them once, like __graft_entry__.build(), when
    the compiler is there.  On the GPU box the prebuilt files travel with the snapshot and nothing is rebuilt."""
    lib = os.path.join(ROOT, "gr-fdc_b200", "lib", "libfdc_b200.so")
    if not os.path.exists(lib) and os.path.exists("/usr/local/cuda/bin/nvcc"):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def ref():
    """The oracle: unmodified gr-FDC blocks built by oracle/Makefile (test infrastructure)."""
    from oracle import fdc_ref
    if not fdc_ref.available():
        pytest.skip("oracle/_ref/libfdc_ref.so not built")
    fdc_ref.set_fft_mode(0)
    return fdc_ref
