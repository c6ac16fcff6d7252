import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ikd-tree_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu` on the GPU box")


def rows(a):
    """Lexicographically sorted [n,3] float32 rows: a canonical form for comparing point SETS."""
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))] if len(a) else a


def same_set(a, b):
    a, b = rows(a), rows(b)
    return a.shape == b.shape and np.array_equal(a, b)


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ikd_golden_v1.npz")))


@pytest.fixture(scope="session")
def built_libs():
    """Make sure the oracle restatement is compiled (it is test infrastructure, built on demand)."""
    import subprocess
    import ref_ctypes as R
    if not R.oracle_available():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return True
