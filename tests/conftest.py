import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g

    g.build()
    return g


@pytest.fixture(scope="session")
def oracle(built):
    """The CPU oracle (checker).  Only tests may load it."""
    from opencafemol_b200.capi import Library

    return Library(built.ORACLE_SO, "oracle_")


@pytest.fixture(scope="session")
def engine(built):
    """The CUDA engine through its C ABI; fails loudly when absent."""
    from opencafemol_b200.capi import product_library

    lib = product_library()
    if lib.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box")
    return lib


def load_golden_spec(name):
    from opencafemol_b200.toml_input import spec_from_dict

    with gzip.open(os.path.join(GOLDEN, name + "_spec.json.gz"), "rt") as fh:
        return spec_from_dict(json.load(fh))


def load_golden_values(name):
    with open(os.path.join(GOLDEN, name + "_oracle.json")) as fh:
        return json.load(fh)
