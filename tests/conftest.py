import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) device; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def built_libs():
    """Builds (if stale) and returns the paths of the in-tree native libraries.  nvcc cross-compiles without a GPU."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import build as build_mod
    return build_mod.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import nst_oracle
    return nst_oracle


@pytest.fixture(scope="session")
def vgg_weights(oracle):
    return oracle.vgg19_random_weights(1234, 13)


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
