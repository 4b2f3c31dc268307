"""pytest configuration: the ``gpu`` marker, repo root on sys.path, golden-vector loaders."""
from __future__ import annotations

import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name: str):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def fixture_vectors():
    return load_golden("fixture_vectors.json")


@pytest.fixture(scope="session")
def dealer_vectors():
    return load_golden("dealer_vectors.json")


@pytest.fixture(scope="session")
def biprime_vectors():
    return load_golden("biprime_vectors.json")


# measurement / profiling tools live under tests/tools (they may use the oracle as a checker) but are
# not test modules; coop_model.py and ref_harness.py are helpers imported by tests
collect_ignore_glob = ["tools/*"]
collect_ignore = ["coop_model.py", "ref_harness.py"]
