import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_GPU_ORDER = {"test_gpu_parity.py": 0, "test_gpu_driver.py": 1, "test_native_driver_gpu.py": 2}


def pytest_collection_modifyitems(config, items):
    """GPU run order: the kernel parity gate (C ABI vs oracle / goldens) first, then the Python driver above it, then the native
    driver, which is held to the Python one.  Stable sort: everything else keeps its collection order."""
    items.sort(key=lambda it: _GPU_ORDER.get(os.path.basename(str(it.fspath)), -1))


@pytest.fixture(scope="session")
def cases(tmp_path_factory):
    """MOL + XDENS text files of the reference's runnable test cases, rebuilt from tests/golden."""
    import fixtures
    return fixtures.materialize(tmp_path_factory.mktemp("cases"))
