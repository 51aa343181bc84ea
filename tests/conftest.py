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


# Order of the test modules in a run: the hot path first (the C ABI, the SPB = 1 Dornic + RK4 step against the oracle at every size, its
# host layer, the noise-off path, the row slabs), then the widened rows (FKE movement, consumer steps), then the modules whose device
# paths were built last (test_zz_*), and the one module that executes a kernel which is not the product's (test_zzz_*).  Within a
# module the order is the file's.  With `-x` a failure in a widened row therefore cannot hide the result of the hot path's own tests.
_MODULE_ORDER = ["test_abi", "test_sampler_spec", "test_oracle_golden", "test_kernel_emulation", "test_gpu_parity",
                 "test_gpu_fullsize_properties", "test_gpu_host_layer", "test_host_layer_cpu", "test_dp_noise_off", "test_multi_rank_cpu",
                 "test_gpu_slab", "test_fke_oracle", "test_fke_kernels", "test_fke_host_layer", "test_consumer_oracle",
                 "test_consumer_kernels", "test_consumer_step2", "test_consumer3_oracle", "test_fpe_runner_oracle", "test_side_outputs"]


def pytest_collection_modifyitems(session, config, items):
    def rank(item):
        name = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        if name in _MODULE_ORDER:
            return (0, _MODULE_ORDER.index(name), "")
        return (1, 0, name)          # everything else after, alphabetically (test_zz_*, then test_zzz_*)
    items.sort(key=rank)             # stable: the order inside a module is kept
