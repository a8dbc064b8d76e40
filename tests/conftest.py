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
    """GPU run order: the kernel parity gate (C ABI vs oracle / goldens) first, then the driver through its library entry (Python launcher), then the
    gimic-b200 program, which is held to the library entry.  Stable sort: everything else keeps its collection order."""
    items.sort(key=lambda it: _GPU_ORDER.get(os.path.basename(str(it.fspath)), -1))
    # a plain `pytest tests` on a host without a CUDA device (or without the built library) skips the GPU tests instead of erroring
    if any("gpu" in it.keywords for it in items) and not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device on this host: GPU parity tests need a B200 (the product has no CPU path)")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


def _have_gpu():
    """decided WITHOUT the product library: on a box that has a GPU a missing or broken libgimic_b200.so must fail the tests loudly"""
    if os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0"):
        return True
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def cases(tmp_path_factory):
    """MOL + XDENS text files of the reference's runnable test cases, rebuilt from tests/golden."""
    import fixtures
    return fixtures.materialize(tmp_path_factory.mktemp("cases"))
