import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's call sites replayed on the system libzstd; tests skip when it is absent."""
    import pyoracle
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def corpus():
    return importlib.import_module("fuse-zstd_b200.corpus")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        man = json.load(fh)
    out = {}
    for name, meta in man["vectors"].items():
        with open(os.path.join(GOLDEN, name + ".zst"), "rb") as fh:
            out[name] = (fh.read(), meta)
    return out
