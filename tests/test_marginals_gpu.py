"""Marginal covariances on the GPU (spp_ba_marginals / spp_schur_marginals, csrc/marginals.cu)
against (a) a float64 dense inverse of the SAME lambda -- well conditioned: the reference's own damped lambda injected
through slot 1, or the device linearisation with damping -- to 1e-9, and (b) the golden block diagonal the UNMODIFIED
reference recovers from the Schur-complemented system at alpha = 0 (tests/golden/margs_*.npz), up to the variance of
the unobservable gauge modes (see tests/test_marginals_cpu.py). All calls go through the C ABI."""
import numpy as np
import pytest

from conftest import gauge_fit_residual, lambda_to_dense, load_golden, load_margs_golden, rel_err, weakest_modes

pytestmark = pytest.mark.gpu

CASES = ["ba_tiny", "ba_tiny_interleaved", "ba_small", "ba_small_hard"]


def _blocks_of_inverse(A, dims):
    """diagonal blocks of A^-1 grouped by block width (6 first, then 3), each group in block-column order"""
    Ai = np.linalg.inv(A)
    base = np.concatenate([[0], np.cumsum(dims)])
    six = [Ai[base[i]:base[i] + 6, base[i]:base[i] + 6] for i in range(len(dims)) if dims[i] == 6]
    three = [Ai[base[i]:base[i] + 3, base[i]:base[i] + 3] for i in range(len(dims)) if dims[i] == 3]
    return np.stack(six), np.stack(three)


@pytest.mark.parametrize("name", CASES)
def test_slot_marginals_vs_dense_inverse(ctx, name):
    """the reference's own (damped) lambda through slot 1: block diagonal of its inverse to 1e-9"""
    g, d = load_golden(name)
    dims = d["L0.col_dims"].astype(np.int64)
    ctx.schur_symbolic(d["L0.col_dims"], d["L0.col_ptr"], d["L0.row_idx"])
    ctx.schur_solve(d["L0.vals"], d["L0.eta"])
    cc, pc = ctx.schur_marginals(0.0)
    A = lambda_to_dense(d["L0.col_dims"], d["L0.col_ptr"], d["L0.row_idx"], d["L0.vals"])
    rc, rp = _blocks_of_inverse(A, dims)
    assert rel_err(cc, rc) < 1e-9 and rel_err(pc, rp) < 1e-9
    assert np.max(np.abs(cc - cc.transpose(0, 2, 1))) <= 1e-12 * np.abs(cc).max()   # symmetric blocks
    # extra damping goes to the camera and the landmark blocks alike
    cc2, pc2 = ctx.schur_marginals(0.5)
    rc2, rp2 = _blocks_of_inverse(A + 0.5 * np.eye(A.shape[0]), dims)
    assert rel_err(cc2, rc2) < 1e-9 and rel_err(pc2, rp2) < 1e-9


@pytest.mark.parametrize("name", CASES)
def test_ba_marginals_damped_vs_dense_inverse(ctx, name):
    """device linearisation + damping (well conditioned) against the dense inverse of the device's own lambda"""
    g, d = load_golden(name)
    ctx.ba_set_graph(g)
    ctx.ba_linearise()
    col_dims, col_ptr, row_idx, vals, eta = ctx.ba_get_lambda()
    A = lambda_to_dense(col_dims, col_ptr, row_idx, vals)
    alpha = 1e-3 * float(np.diag(A).max())
    cc, pc = ctx.ba_marginals(alpha)
    rc, rp = _blocks_of_inverse(A + alpha * np.eye(A.shape[0]), np.asarray(col_dims, np.int64))
    assert rel_err(cc, rc) < 1e-9 and rel_err(pc, rp) < 1e-9
    # a solve that follows is not disturbed by what the marginals left behind
    dx = ctx.ba_solve_step(2 * alpha)
    assert rel_err(dx, np.linalg.solve(A + 2 * alpha * np.eye(A.shape[0]), eta)) < 1e-9


@pytest.mark.parametrize("name", ["margs_tiny", "margs_tiny_interleaved", "margs_small"])
def test_ba_marginals_vs_reference(ctx, name):
    """alpha = 0 at the reference's final states against the reference's Schur_Marginals output"""
    g, d = load_margs_golden(name)
    ctx.ba_set_graph(g)
    cc, pc = ctx.ba_marginals(0.0)
    assert cc.shape == d["cam_cov"].shape and pc.shape == d["pt_cov"].shape
    col_dims, col_ptr, row_idx, vals, _ = ctx.ba_get_lambda()
    A = lambda_to_dense(col_dims, col_ptr, row_idx, vals)
    # the device's lambda comes in vertex id order; the gauge fit wants the cameras first
    dims = np.asarray(col_dims, np.int64)
    base = np.concatenate([[0], np.cumsum(dims)])
    perm = np.concatenate([np.arange(base[i], base[i + 1]) for i in np.flatnonzero(dims == 6)] +
                          [np.arange(base[i], base[i + 1]) for i in np.flatnonzero(dims == 3)])
    L = A[np.ix_(perm, perm)]
    m = 1 if g.vtype[0] == 0 else 4
    rc, rp, k = gauge_fit_residual(g.n_cams, cc, pc, d["cam_cov"], d["pt_cov"], weakest_modes(L, m))
    assert rc < 1e-4 and rp < 1e-3, (rc, rp, k)
    assert np.all(np.einsum("kii->ki", cc) > 0) and np.all(np.einsum("kii->ki", pc) > 0)


def test_marginals_mid_size_and_errors(ctx):
    """a reduced system that spans several 128-wide panels (C = 60 -> ld = 384) with long tracks"""
    from slam_plus_plus_b200 import capi, graphs
    g = graphs.make_ba(60, 2000, 11, mean_extra_track=8.0)
    ctx.ba_set_graph(g)
    ctx.ba_linearise()
    col_dims, col_ptr, row_idx, vals, _ = ctx.ba_get_lambda()
    A = lambda_to_dense(col_dims, col_ptr, row_idx, vals)
    alpha = 1e-4 * float(np.diag(A).max())
    cc, pc = ctx.ba_marginals(alpha)
    rc, rp = _blocks_of_inverse(A + alpha * np.eye(A.shape[0]), np.asarray(col_dims, np.int64))
    assert rel_err(cc, rc) < 1e-9 and rel_err(pc, rp) < 1e-9
    cc2, pc2 = ctx.ba_marginals(alpha)
    assert np.array_equal(cc, cc2) and np.array_equal(pc, pc2)       # no atomics: bit-reproducible
    # not positive definite -> the status of the factorisation, as a solve reports it
    with pytest.raises(capi.NotPositiveDefinite):
        ctx.ba_marginals(-10.0 * float(np.diag(A).max()))
    # the block-sparse path has no dense inverse to read from
    ctx.schur_set_rcs_solver(capi.RCS_SPARSE)
    try:
        with pytest.raises(capi.SppError):
            ctx.ba_marginals(alpha)
    finally:
        ctx.schur_set_rcs_solver(capi.RCS_AUTO)


@pytest.mark.parametrize("name,tol", [("margs_se2", 1e-9), ("margs_se3", 1e-3)])
def test_pose_marginals_vs_reference(ctx, name, tol):
    """pose graphs (gauge fixed by the unary factor on pose 0): spp_pose_marginals at the reference's final states
    against the reference's own marginals -- SE(2) to 1e-9, SE(3) at the FD noise floor times cond(lambda) ~ 1e11 --
    and against a float64 dense inverse of the device's lambda"""
    from test_pose_cpu import load_pose_golden, pose_lambda_to_dense
    g, d = load_pose_golden(name)
    ctx.pose_set_graph(g)
    ctx.pose_set_states(d["states"])
    cov = ctx.pose_marginals()
    assert cov.shape == d["cov"].shape
    cp, ri, vals, _ = ctx.pose_get_lambda()
    dim = cov.shape[1]
    A = pose_lambda_to_dense(cp, ri, vals, dim)
    assert rel_err(cov, d["cov"]) < max(tol, 1e-2 * np.linalg.cond(A) * np.finfo(float).eps)
    Ai = np.linalg.inv(A)
    ref = np.stack([Ai[dim * i:dim * i + dim, dim * i:dim * i + dim] for i in range(cov.shape[0])])
    assert rel_err(cov, ref) < max(1e-9, 1e-2 * np.linalg.cond(A) * np.finfo(float).eps)
    assert np.array_equal(cov, ctx.pose_marginals())  # bit-reproducible


def test_pose_marginals_manhattan_size(ctx):
    """Manhattan-3500 shape (10 500 unknowns, 83 panels): consistency of the dense inverse with a solve of the same
    system -- Sigma_ii is the i-th diagonal block of lambda^-1, so lambda-solves against unit vectors must reproduce it"""
    from slam_plus_plus_b200 import graphs
    from test_pose_cpu import pose_lambda_to_dense
    import scipy.sparse
    import scipy.sparse.linalg
    g = graphs.make_manhattan(fill_loops=True)
    ctx.pose_set_graph(g)
    ctx.pose_optimize(5, 0.0)
    ctx.pose_linearise()
    cov = ctx.pose_marginals()
    n = g.poses.shape[0]
    assert cov.shape == (n, 3, 3) and np.all(np.einsum("kii->ki", cov) > 0)
    cp, ri, vals, _ = ctx.pose_get_lambda()
    # sparse lambda (upper blocks -> symmetric) and a handful of columns of its inverse
    rows, cols, v = [], [], []
    for c in range(n):
        for k in range(int(cp[c]), int(cp[c + 1])):
            r = int(ri[k])
            blk = vals[9 * k:9 * k + 9].reshape(3, 3).T
            for i in range(3):
                for j in range(3):
                    if r < c or i <= j:
                        rows.append(3 * r + i); cols.append(3 * c + j); v.append(blk[i, j])
                        if 3 * r + i != 3 * c + j:
                            rows.append(3 * c + j); cols.append(3 * r + i); v.append(blk[i, j])
    A = scipy.sparse.csc_matrix((v, (rows, cols)), shape=(3 * n, 3 * n))
    lu = scipy.sparse.linalg.splu(A)
    for p in (0, 1, n // 3, n // 2, n - 1):
        E = np.zeros((3 * n, 3))
        E[3 * p:3 * p + 3] = np.eye(3)
        X = lu.solve(E)
        # two stable inversions agree to O(cond * eps): cond(lambda) ~ 2e10 on this graph (identity prior on pose 0
        # against edge information of 400 .. 2500, see test_pose_cpu.dx_tolerance) -- measured 1.4e-6
        assert rel_err(cov[p], X[3 * p:3 * p + 3]) < 1e-5
