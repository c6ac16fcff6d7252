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


def replay_delete_by_point(D, pts):
    """Delete_by_point (ikd_Tree.cpp:713-760) replayed on a pre-order structure dump (columns: x y z axis size invalid
    flags ... has_left has_right): descend one side only (p[axis] < node[axis] -> left, else right), at every node first
    test same_point (per-axis |diff| < 1e-6 in double) && !point_deleted. Returns the dump rows that get deleted."""
    n = len(D)
    left = np.full(n, -1)
    right = np.full(n, -1)
    pos = [0]

    def rec():
        i = pos[0]
        pos[0] += 1
        if D[i, 13]:
            left[i] = pos[0]
            rec()
        if D[i, 14]:
            right[i] = pos[0]
            rec()
        return i

    sys.setrecursionlimit(10000)
    rec()
    deleted = (D[:, 6].astype(int) & 1).astype(bool).copy()
    hit = []
    for p in pts:
        cur = 0
        while cur >= 0:
            if int(D[cur, 6]) & 2:  # tree_deleted (:714)
                break
            same = all(abs(float(np.float32(D[cur, a] - p[a]))) < 1e-6 for a in range(3))
            if same and not deleted[cur]:
                deleted[cur] = True
                hit.append(cur)
                break
            a = int(D[cur, 3])
            cur = left[cur] if p[a] < D[cur, a] else right[cur]
    return hit


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
