"""Edge cases of the C ABI on the GPU: empty and ragged inputs, isolated vertices, duplicate pose edges, call-order
errors. The library must answer with a status (never crash or hang), and where the problem is well posed the result must
agree with a dense float64 solve of the same system."""
import numpy as np
import pytest

from conftest import lambda_to_dense, rel_err

pytestmark = pytest.mark.gpu


def _ba(n_cams=5, n_pts=30, seed=3, **kw):
    from slam_plus_plus_b200 import graphs
    return graphs.make_ba(n_cams, n_pts, seed, mean_extra_track=2.0, max_track=4, max_stride=1, **kw)


def test_ba_without_observations(ctx):
    from slam_plus_plus_b200 import capi
    from slam_plus_plus_b200.sppio import BAGraph
    g = _ba()
    e = BAGraph(g.vtype, g.cams, g.pts, g.obs_pt[:0], g.obs_cam[:0], g.z[:0], g.info[:0])
    ctx.ba_set_graph(e)
    assert ctx.ba_chi2() == 0.0
    rep = ctx.ba_optimize(3, 0.0)  # "the system contains no edges": nothing to do, no failure
    assert rep["chi2_final"] == 0.0 and rep["status"] in (capi.SPP_OK, capi.SPP_NOT_POSDEF)


@pytest.mark.parametrize("rcs", ["dense", "sparse"])
def test_camera_without_observations_and_single_observation_points(ctx, rcs):
    """an isolated camera has an all-zero Hessian block: with damping the step exists and leaves that camera alone"""
    from slam_plus_plus_b200 import capi
    from slam_plus_plus_b200.sppio import BAGraph
    g = _ba(n_cams=6, n_pts=40)
    lone = 5
    keep = g.obs_cam != lone
    # drop all but one observation of a few points as well (ragged tracks of length 1)
    loc = g.vertex_local_index()
    first = np.ones(len(keep), bool)
    seen = set()
    for k in range(len(keep)):
        p = int(loc[g.obs_pt[k]])
        if p < 5:
            first[k] = p not in seen
            seen.add(p)
    keep &= first
    h = BAGraph(g.vtype, g.cams, g.pts, g.obs_pt[keep], g.obs_cam[keep], g.z[keep], g.info[keep])
    ctx.schur_set_rcs_solver(capi.RCS_SPARSE if rcs == "sparse" else capi.RCS_DENSE)
    try:
        ctx.ba_set_graph(h)
        ctx.ba_linearise()
        alpha = 10.0
        dx = ctx.ba_solve_step(alpha)
        assert np.all(np.isfinite(dx))
        col_dims, col_ptr, row_idx, vals, eta = ctx.ba_get_lambda()
        A = lambda_to_dense(col_dims, col_ptr, row_idx, vals) + alpha * np.eye(len(eta))
        assert rel_err(dx, np.linalg.solve(A, eta)) < 1e-9
        base = int(np.concatenate([[0], np.cumsum(col_dims)])[lone])
        assert np.allclose(dx[base:base + 6], 0.0)  # nothing pulls on the isolated camera
        with pytest.raises(capi.NotPositiveDefinite):  # undamped: singular, reported as the reference reports it
            ctx.ba_solve_step(0.0)
    finally:
        ctx.schur_set_rcs_solver(capi.RCS_AUTO)


def test_single_camera_graph(ctx):
    g = _ba(n_cams=1, n_pts=12)
    ctx.ba_set_graph(g)
    rep = ctx.ba_optimize(3, 0.0)
    assert np.isfinite(rep["chi2_final"]) and rep["chi2_final"] <= rep["chi2_initial"] * (1 + 1e-12)


def test_bad_references_are_refused(ctx):
    from slam_plus_plus_b200 import capi
    from slam_plus_plus_b200.sppio import BAGraph
    g = _ba()
    bad_pt = g.obs_pt.copy()
    bad_pt[0] = 0  # vertex 0 is a camera
    with pytest.raises(capi.SppError):
        ctx.ba_set_graph(BAGraph(g.vtype, g.cams, g.pts, bad_pt, g.obs_cam, g.z, g.info))
    bad_cam = g.obs_cam.copy()
    bad_cam[1] = len(g.vtype) + 7  # out of range
    with pytest.raises(capi.SppError):
        ctx.ba_set_graph(BAGraph(g.vtype, g.cams, g.pts, g.obs_pt, bad_cam, g.z, g.info))
    dup = BAGraph(g.vtype, g.cams, g.pts, np.concatenate([g.obs_pt, g.obs_pt[:1]]), np.concatenate([g.obs_cam, g.obs_cam[:1]]),
                  np.concatenate([g.z, g.z[:1]]), np.concatenate([g.info, g.info[:1]]))
    with pytest.raises(capi.SppError):  # the same landmark twice in one camera
        ctx.ba_set_graph(dup)
    ctx.ba_set_graph(g)  # the context is usable afterwards
    assert np.isfinite(ctx.ba_chi2())


def test_pose_graph_edge_cases(ctx):
    from slam_plus_plus_b200 import capi, graphs
    from slam_plus_plus_b200.sppio import PoseGraph
    g = graphs.make_manhattan(30, 8, seed=4)
    # no edges: nothing to optimise
    e = PoseGraph(g.kind, g.poses, g.e_from[:0], g.e_to[:0], g.z[:0], g.info[:0])
    ctx.pose_set_graph(e)
    rep = ctx.pose_optimize(3, 0.0)
    assert rep["n_iterations"] == 0 and rep["chi2_final"] == 0.0
    # a duplicated edge is a longer source list of the same block: twice the information
    d = PoseGraph(g.kind, g.poses, np.concatenate([g.e_from, g.e_from[:3]]), np.concatenate([g.e_to, g.e_to[:3]]),
                  np.concatenate([g.z, g.z[:3]]), np.concatenate([g.info, g.info[:3]]))
    ctx.pose_set_graph(d)
    ctx.pose_linearise()
    cp, ri, vals, eta = ctx.pose_get_lambda()
    ctx.pose_set_graph(g)
    ctx.pose_linearise()
    cp0, ri0, vals0, eta0 = ctx.pose_get_lambda()
    assert np.array_equal(cp, cp0) and np.array_equal(ri, ri0) and not np.allclose(vals, vals0)
    # an edge from a vertex to itself or out of range is refused
    with pytest.raises(capi.SppError):
        ctx.pose_set_graph(PoseGraph(g.kind, g.poses, np.array([1]), np.array([1]), g.z[:1], g.info[:1]))
    with pytest.raises(capi.SppError):
        ctx.pose_set_graph(PoseGraph(g.kind, g.poses, np.array([1]), np.array([99]), g.z[:1], g.info[:1]))


def test_call_order_errors(ctx):
    from slam_plus_plus_b200 import capi
    c = capi.Context(0)
    try:
        with pytest.raises(capi.SppError):
            c.ba_chi2()  # no graph yet
        with pytest.raises(capi.SppError):
            c.schur_solve(np.zeros(4), np.zeros(2))  # no symbolic decomposition yet
        with pytest.raises(capi.SppError):
            c.schur_get_rcs_info()
    finally:
        c.close()


def test_ba_set_states_moves_the_linearisation_point(ctx):
    """spp_ba_set_states (what the slot-3 adapter calls when only the states of an uploaded system changed): chi2,
    lambda and the marginals afterwards are those of a graph uploaded at the new states, bit for bit; get_states
    returns what was set; restore_initial goes back to the states of spp_ba_set_graph"""
    from conftest import load_golden, load_margs_golden
    g0, d = load_golden("margs_small")            # the graph at its initial states
    g1, _ = load_margs_golden("margs_small")      # the same graph at the reference's final states
    ctx.ba_set_graph(g1)
    chi2_1 = ctx.ba_chi2()
    ctx.ba_linearise()
    vals_1 = ctx.ba_get_lambda()[3]
    ctx.ba_set_graph(g0)
    chi2_0 = ctx.ba_chi2()
    assert abs(chi2_1 - d["chi2"][0]) <= 1e-11 * d["chi2"][0] and chi2_0 > chi2_1
    ctx.ba_set_states(g1.cams[:, :6], g1.pts)
    assert ctx.ba_chi2() == chi2_1
    ctx.ba_linearise()
    assert np.array_equal(ctx.ba_get_lambda()[3], vals_1)
    cams, pts = ctx.ba_get_states()
    assert np.array_equal(cams, g1.cams[:, :6]) and np.array_equal(pts, g1.pts)
    ctx.ba_set_states(None, g0.pts)               # either array may be left alone
    cams, pts = ctx.ba_get_states()
    assert np.array_equal(cams, g1.cams[:, :6]) and np.array_equal(pts, g0.pts)
    ctx.ba_restore_initial()
    assert ctx.ba_chi2() == chi2_0
