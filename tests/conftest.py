import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


CASES = ["case_transloc_10k", "case_inversion_10k", "case_complex_10k", "case_deletion_5k", "case_transloc_25k"]


@pytest.fixture(scope="session")
def models():
    import joblib
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = joblib.load(os.path.join(GOLDEN, name))
        return cache[name]
    return get
