import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # TEST TOOL (tests/emu/mock/README.md): dry-run of -m gpu tests on host threads.  With TM_TEST_MOCK_LIB naming the mock C-ABI library
    # (and LD_PRELOAD carrying it into libtm_host.so and the driver binaries) the ctypes view is pointed at it from HERE -- the package
    # itself has no such switch.
    mock = os.environ.get("TM_TEST_MOCK_LIB")
    if mock:
        import trophic_maelstorm_b200.lattice as lat
        lat.lib_path = lambda: mock


@pytest.fixture(scope="session")
def orc():
    import oracle_lib
    return oracle_lib
