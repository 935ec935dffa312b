import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "vecchia_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_lib():
    """ctypes handle of the UNMODIFIED reference build (oracle/_ref), or None where it was not built."""
    from oracle import ref_lib_path
    p = ref_lib_path()
    if not os.path.exists(p):
        return None
    from gpboost_b200.libpath import load_lib
    return load_lib(p)


@pytest.fixture(scope="session")
def product_lib():
    import __graft_entry__  # noqa: F401  (builds nothing; the library must already be in-tree)
    from gpboost_b200.libpath import load_lib
    return load_lib()


def case_data(spec):
    import datagen
    if spec["data"] == "r_test":
        return datagen.r_test_data()
    if spec["data"] == "lattice":
        c = datagen.lattice(spec["k"])
        rng = np.random.default_rng(spec["seed"])
        return c, rng.standard_normal(c.shape[0])
    return datagen.synth(spec["n"], spec.get("d", 2), spec["seed"])
