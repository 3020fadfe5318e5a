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


@pytest.fixture(scope="session")
def bal_rcs_pattern():
    """block pattern of the reduced camera system of the BAL-13682 shape (13 682 block columns, 2.27 M upper blocks):
    half a minute of numpy, shared by the CPU tests that need it"""
    from slam_plus_plus_b200 import graphs
    return graphs.rcs_block_pattern(graphs.ba_shape("bal13682"))


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


def load_margs_golden(name):
    """Golden marginals of the reference: the graph at the states the covariance was taken at, cam_cov, pt_cov."""
    g, d = load_golden(name)
    loc = g.vertex_local_index()
    st, off = d["states"], 0
    for v in range(g.n_vertices):
        if g.vtype[v] == 0:
            g.cams[loc[v], :6] = st[off:off + 6]
            off += 6
        else:
            g.pts[loc[v]] = st[off:off + 3]
            off += 3
    return g, d


def weakest_modes(L, m=1):
    """Eigenvectors of the m smallest eigenvalues of a symmetric matrix by inverse subspace iteration -- the gauge
    freedom the unary factor on vertex 0 leaves open: the scale of the scene when vertex 0 is a camera (m = 1), scale
    and rotation about the fixed point when it is a landmark (m = 4)."""
    import scipy.linalg
    lu = scipy.linalg.lu_factor(L)
    v = np.random.default_rng(0).normal(size=(L.shape[0], m))
    for _ in range(5):
        v, _ = np.linalg.qr(scipy.linalg.lu_solve(lu, v))
    return v


def gauge_fit_residual(n_cams, cam_a, pt_a, cam_b, pt_b, v):
    """Two block-diagonal covariances of a monocular BA system differ by V K V^T, V = the gauge modes (n x m), whose
    variances 1 / lambda are finite-difference noise (lambda ~ 1e-8 against |lambda| ~ 1e6). Fits the m (m + 1) / 2
    scalars of the symmetric K on all blocks and returns the largest remaining difference, relative to the largest
    covariance entry, for the cameras and for the points."""
    m, o = v.shape[1], 6 * n_cams

    def blocks(M):
        return (np.stack([M[6 * i:6 * i + 6, 6 * i:6 * i + 6] for i in range(n_cams)]),
                np.stack([M[o + 3 * j:o + 3 * j + 3, o + 3 * j:o + 3 * j + 3] for j in range(len(pt_a))]))
    basis = []
    for i in range(m):
        for j in range(i, m):
            B = np.outer(v[:, i], v[:, j])
            basis.append(blocks(B + B.T if i != j else B))
    dc, dp = cam_a - cam_b, pt_a - pt_b
    X = np.stack([np.concatenate([bc.ravel(), bp.ravel()]) for bc, bp in basis], 1)
    k = np.linalg.lstsq(X, np.concatenate([dc.ravel(), dp.ravel()]), rcond=None)[0]
    fc = sum(ki * bc for ki, (bc, _) in zip(k, basis))
    fp = sum(ki * bp for ki, (_, bp) in zip(k, basis))
    return float(np.abs(dc - fc).max() / np.abs(cam_b).max()), float(np.abs(dp - fp).max() / np.abs(pt_b).max()), k
