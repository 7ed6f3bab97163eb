import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_small_cases():
    with open(os.path.join(GOLDEN, "small_cases.json")) as f:
        return json.load(f)["cases"]


def load_big_facts():
    with open(os.path.join(GOLDEN, "big_4x1M.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def small_cases():
    return load_small_cases()


@pytest.fixture(scope="session")
def big_facts():
    return load_big_facts()
