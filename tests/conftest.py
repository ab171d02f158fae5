import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def model_dirs(tmp_path_factory):
    """case name -> model dir identical to the one the goldens were generated with."""
    from tests.cases import ALL_CASES, case_model_dir

    root = tmp_path_factory.mktemp("models")
    cache = {}

    def get(name):
        if name not in cache:
            assert name in ALL_CASES
            cache[name] = case_model_dir(name, root)
        return cache[name]

    return get
