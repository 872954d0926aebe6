import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """The oracle: unmodified gr-FDC blocks built by oracle/Makefile (test infrastructure)."""
    from oracle import fdc_ref
    if not fdc_ref.available():
        pytest.skip("oracle/_ref/libfdc_ref.so not built")
    fdc_ref.set_fft_mode(0)
    return fdc_ref
