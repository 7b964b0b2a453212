import importlib.util
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


def load_pkg():  # same loader as __graft_entry__.load_pkg
    """Import the package directory `elections.dtree_b200/` (its name is not a Python identifier)."""
    name = "elections_dtree_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "elections.dtree_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(path, "__init__.py"),
                                                  submodule_search_locations=[path])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def port():
    from oracle import oracle
    return oracle.load("port")


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref/libdtree_ref.so not built (needs /root/reference)")
    return oracle.load("reference")


@pytest.fixture(scope="session")
def ref_vectors():
    with open(os.path.join(GOLDEN, "ref_vectors.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def wakehurst():
    with open(os.path.join(GOLDEN, "wakehurst2023.json")) as f:
        return json.load(f)
