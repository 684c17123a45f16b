"""Shared fixtures.  `-m "not gpu"` tests run on CPU (oracle vs golden vectors, host logic, C-ABI symbol
check); `-m gpu` tests are the parity tests proper and call the CUDA path through the C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ltr-lowrank-sdp_b200")
ORACLE = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")
INST = os.path.join(GOLDEN, "instances")
for p in (PKG, ORACLE):
    if p not in sys.path:
        sys.path.insert(0, p)

FIXTURES = ["G11", "maxcut_torus_8x10", "maxcut_torus_20x30", "general_sparse_n60", "theta_n30",
            "dense_constraint_n24", "multiblock_sdp", "multiblock_lp", "control_like_12_6"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def inst_path(name):
    return os.path.join(INST, name + ".dat-s")


@pytest.fixture(scope="session")
def built():
    """Build the in-tree libraries if they are missing (they travel prebuilt to the GPU box)."""
    import lorads_b200 as lb
    if not (os.path.exists(lb.LIB_PATH) and os.path.exists(lb.HOST_LIB_PATH) and os.path.exists(lb.BINARY_PATH)):
        lb.build()
    return lb


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b))) / den if a.size else 0.0
