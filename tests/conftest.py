"""Test configuration. `-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors
and vs the compiled reference, host logic, C-ABI symbol checks, gloo multi-process logic);
`-m gpu` is the parity suite proper: every call goes through the C ABI of include/apj_b200.h."""
import os
import sys

import pytest

# several slab ranks share the one GPU of the test box: give every handle's stream its own hardware queue
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the compiled reference; built where /root/reference exists)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checkers (oracle, and oracle/_ref where the reference tree exists) and make sure
    the CUDA library is present (nvcc cross-compiles here; on the GPU box the .so travels)."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.ORACLE_SO) or (os.path.isdir("/root/reference") and not pyoracle.have_ref()):
        pyoracle.build()
    from active_particle_jamming_b200 import _build
    if not os.path.exists(_build.LIB):
        _build.build_library()
    yield
