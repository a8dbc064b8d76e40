import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cases(tmp_path_factory):
    """MOL + XDENS text files of the reference's runnable test cases, rebuilt from tests/golden."""
    import fixtures
    return fixtures.materialize(tmp_path_factory.mktemp("cases"))
