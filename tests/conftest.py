import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    from slam_plus_plus_b200.sppio import BAGraph
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    g = BAGraph(d["g_vtype"], d["g_cams"], d["g_pts"], d["g_obs_pt"], d["g_obs_cam"], d["g_z"], d["g_info"])
    return g, d


@pytest.fixture(scope="session")
def ctx():
    from slam_plus_plus_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def lambda_to_dense(col_dims, col_ptr, row_idx, vals, symmetric=True):
    """Block structure (reference layout: upper block-triangular, column-major blocks) -> dense matrix."""
    col_dims = np.asarray(col_dims, np.int64)
    base = np.concatenate([[0], np.cumsum(col_dims)])
    n = int(base[-1])
    A = np.zeros((n, n))
    off = 0
    for c in range(len(col_dims)):
        for k in range(int(col_ptr[c]), int(col_ptr[c + 1])):
            r = int(row_idx[k])
            h, w = int(col_dims[r]), int(col_dims[c])
            A[base[r]:base[r] + h, base[c]:base[c] + w] = vals[off:off + h * w].reshape(w, h).T
            off += h * w
    if symmetric:
        A = np.triu(A) + np.triu(A, 1).T
    return A


def rel_err(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
