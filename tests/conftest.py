import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


if os.environ.get("AXR_SIMT_TESTS_ONLY") == "1":
    # tests/test_simt_kernels.py re-runs the GPU tests in a subprocess against the SIMT-interpreter build of the CUDA sources; the
    # product loader refuses that build, so the harness binds it itself (test process only)
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
    import use_simt
    use_simt.install()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def po():
    """CPU checkers (oracle/): test infrastructure only."""
    from oracle import pyoracle
    pyoracle.build(ref=True)
    return pyoracle
