"""CPU tests of the SE(2) pose-graph oracle (oracle/spp_oracle.c) against the golden vectors produced by the unmodified
reference (tests/golden/se2_*.npz; oracle/_ref/ref_driver_pose = CNonlinearSolver_Lambda + CLinearSolver_UberBlock)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, GOLDEN, rel_err

sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = ["se2_tiny", "se2_small"]


def load_pose_golden(name):
    from slam_plus_plus_b200.sppio import PoseGraph
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    g = PoseGraph(int(d["g_kind"][0]), d["g_poses"], d["g_from"], d["g_to"], d["g_z"], d["g_info"])
    return g, d


def dx_tolerance(A):
    """Two backward-stable solvers of the same SPD system agree to O(cond * eps); the unary factor (identity) against
    edge information of 400 .. 2500 makes these pose graphs ill-conditioned (cond 1e8 .. 2e10), and the reference's own
    block Cholesky differs from LAPACK by 1e-10 .. 1e-8 on them. The north-star bound of 1e-9 applies where cond allows."""
    return max(1e-9, 1e-2 * np.linalg.cond(A) * np.finfo(float).eps)


def pose_lambda_to_dense(col_ptr, row_idx, vals, B):
    n = len(col_ptr) - 1
    A = np.zeros((n * B, n * B))
    off = 0
    for c in range(n):
        for k in range(int(col_ptr[c]), int(col_ptr[c + 1])):
            r = int(row_idx[k])
            A[r * B:(r + 1) * B, c * B:(c + 1) * B] = vals[off:off + B * B].reshape(B, B).T
            off += B * B
    return np.triu(A) + np.triu(A, 1).T


@pytest.mark.parametrize("name", CASES)
def test_oracle_chi2_and_linearisation(name):
    import oracle as orc
    g, d = load_pose_golden(name)
    assert abs(orc.se2_chi2(g) - d["chi2_0"][0]) <= 1e-12 * d["chi2_0"][0]
    lam, eta = orc.se2_linearise_dense(g)
    A_ref = pose_lambda_to_dense(d["L0.col_ptr"], d["L0.row_idx"], d["L0.vals"], 3)
    # analytic Jacobians: no finite-difference noise, the restatement reproduces the reference to rounding
    assert rel_err(lam, A_ref) < 1e-13
    assert rel_err(eta, d["L0.eta"]) < 1e-12
    # first Gauss-Newton increment: the reference's block Cholesky vs a dense solve of the same system
    assert rel_err(np.linalg.solve(A_ref, d["L0.eta"]), d["L0.dx"]) < dx_tolerance(A_ref)


@pytest.mark.parametrize("name", CASES)
def test_oracle_gauss_newton(name):
    import oracle as orc
    g, d = load_pose_golden(name)
    r = orc.se2_optimize(g, int(d["max_iter"][0]), 0.0)
    assert r["status"] == 0 and r["n_solves"] == int(d["n_solves"][0])
    assert abs(r["chi2_final"] - d["chi2"][0]) <= 1e-9 * d["chi2"][0]
    for k in range(r["n_solves"]):
        ref = np.linalg.norm(d[f"L{k}.dx"])
        assert abs(r["dx_norms"][k] - ref) <= 1e-6 * np.linalg.norm(d["L0.dx"])
    assert rel_err(r["poses"].ravel(), d["states"]) < 1e-7


def test_golden_structure_is_upper_with_diagonal_last():
    for name in CASES:
        g, d = load_pose_golden(name)
        cp, ri = d["L0.col_ptr"].astype(int), d["L0.row_idx"].astype(int)
        for c in range(len(cp) - 1):
            rows = ri[cp[c]:cp[c + 1]]
            assert rows[-1] == c and np.all(np.diff(rows) > 0)
