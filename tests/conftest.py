import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "refbin: needs the compiled reference under oracle/_ref (skipped when absent)")


@pytest.fixture(scope="session")
def tmp_cases(tmp_path_factory):
    return tmp_path_factory.mktemp("cases")
