import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"  # only ever read by `not gpu` tests, and only when present


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


@pytest.fixture(scope="session")
def novel():
    """The config-1 text (reference examples/mrs_dalloway.txt, shipped as a gzip'd data fixture)."""
    with gzip.open(os.path.join(GOLDEN, "mrs_dalloway.txt.gz"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def golden_config1():
    return json.load(open(os.path.join(GOLDEN, "config1.json")))


@pytest.fixture(scope="session")
def golden_kats():
    return json.load(open(os.path.join(GOLDEN, "kat_small.json")))


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    from oracle import pyoracle

    pyoracle.build()


def oracle_kinds():
    from oracle import pyoracle

    return [k for k in ("ref_meyer", "ref_classic", "port") if k == "port" or pyoracle.available(k) or os.path.exists(REFERENCE)]
