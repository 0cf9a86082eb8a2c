import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand with the committed Makefile."""
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def c1():
    """Config-1 workload, written by tools/make_c1_input.py."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "c1_input.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def b2():
    """The product package; importing it must load the CUDA library (no fallback)."""
    import multi_sensor_slam_tookit_b200 as pkg
    from multi_sensor_slam_tookit_b200 import capi
    capi.lib()
    return pkg
