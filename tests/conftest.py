import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The patched reference itself (oracle/_ref); skipped when it has not been built."""
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built (run oracle/build_ref.py where /root/reference exists)")
    return pyref.Ref()


@pytest.fixture(scope="session")
def port():
    """The plain-C oracle port (oracle/libcaaa_oracle.so)."""
    import subprocess
    from oracle import pyref
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libcaaa_oracle.so"])
    return pyref.Port()


@pytest.fixture(scope="session")
def hostsim():
    """The product's device engine compiled for the host (CPU test tier only; see tests/hostsim)."""
    import subprocess
    from oracle import pyref
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostsim")])
    return pyref.HostSim()
