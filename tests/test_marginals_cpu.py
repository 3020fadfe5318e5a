"""Marginal covariances: the numpy restatement (oracle.ba_marginals: plain dense inverse of
lambda) against the golden block diagonal the UNMODIFIED reference recovers from the Schur-complemented system
(tests/golden/margs_*.npz, tests/golden/make_golden_margs.py).

A monocular BA system with one fixed camera has an unobservable scale: lambda has one eigenvalue that is
finite-difference noise (~1e-8 against |lambda| ~ 1e6), its reciprocal dominates every covariance block, and two
evaluations agree only up to k v v^T with v the scale mode (four such modes -- scale and rotation -- when vertex 0, which
carries the unary factor, is a landmark). The comparison fits those scalars and asserts the rest."""
import os
import sys

import numpy as np
import pytest

from conftest import gauge_fit_residual, load_margs_golden, rel_err, weakest_modes

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))

CASES = ["margs_tiny", "margs_tiny_interleaved", "margs_small"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_marginals_vs_reference(name):
    import oracle
    g, d = load_margs_golden(name)
    cc, pc, L = oracle.ba_marginals(g)
    m = 1 if g.vtype[0] == 0 else 4
    if L.shape[0] < 400:
        w = np.linalg.eigvalsh(L)
        assert w[0] > 0 and w[m - 1] < 1e-4 * w[m]        # the gauge modes are well separated from the rest
    rc, rp, k = gauge_fit_residual(g.n_cams, cc, pc, d["cam_cov"], d["pt_cov"], weakest_modes(L, m))
    assert rc < 1e-4 and rp < 1e-3, (rc, rp, k)


def test_oracle_marginals_damped_schur_identity():
    """the Schur form the reference evaluates equals the blocks of the full inverse (well conditioned with damping)"""
    import oracle
    g, _ = load_margs_golden("margs_tiny")
    alpha = 10.0
    cc, pc, L = oracle.ba_marginals(g, alpha)
    n, c = L.shape[0], 6 * g.n_cams
    A, U, D = L[:c, :c] + alpha * np.eye(c), L[:c, c:], L[c:, c:] + alpha * np.eye(n - c)
    Dinv = np.linalg.inv(D)
    Sinv = np.linalg.inv(A - U @ Dinv @ U.T)
    full = Dinv + Dinv @ U.T @ Sinv @ U @ Dinv
    assert rel_err(np.stack([Sinv[6 * i:6 * i + 6, 6 * i:6 * i + 6] for i in range(g.n_cams)]), cc) < 1e-10
    assert rel_err(np.stack([full[3 * j:3 * j + 3, 3 * j:3 * j + 3] for j in range(g.n_pts)]), pc) < 1e-10


@pytest.mark.parametrize("name,tol", [("margs_se2", 1e-9), ("margs_se3", 1e-3)])
def test_oracle_pose_marginals_vs_reference(name, tol):
    """pose graphs: the unary factor on pose 0 fixes the whole gauge, so the block diagonal of a dense inverse of the
    restated lambda meets the reference's recursive formula directly -- SE(2) (analytic Jacobians) to rounding times the
    condition number (measured 4e-11), SE(3) (forward differences, cond 1e11) at the FD noise floor"""
    import oracle
    from test_pose_cpu import load_pose_golden
    g, d = load_pose_golden(name)
    L, _ = oracle.pose_linearise_dense(g, d["states"])
    S = np.linalg.inv(L)
    n, dim = d["states"].shape
    cov = np.stack([S[dim * i:dim * i + dim, dim * i:dim * i + dim] for i in range(n)])
    from test_pose_cpu import dx_tolerance
    assert rel_err(cov, d["cov"]) < max(tol, dx_tolerance(L))  # O(cond * eps) between two stable inversions
