import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def packaged():
    """The reference's packaged data as committed fixtures (tests/golden/make_packaged_data.py)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "packaged_data.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through its C-ABI.  No fallback: on a box without a GPU this fails loudly."""
    from dmphyclus_b200 import build
    from dmphyclus_b200.engine import Engine
    build.build()
    return Engine()
