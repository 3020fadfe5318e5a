"""spp_ba_append_graph (incremental bundle adjustment on a device-resident system): appending to the
graph on the device must give, bit for bit, what spp_ba_set_graph of the concatenated graph followed by
spp_ba_set_states of the old vertices' current states gives -- the reference's own incremental BA re-reads its
append-only system the same way (NonlinearSolver_Lambda_LM.h:796-830, Lambda_Base.h:1665-1684)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _split(g, n_v1):
    """prefix = vertices [0, n_v1) with the observations among them (edge order kept); rest = everything else"""
    from slam_plus_plus_b200.sppio import BAGraph
    vt = np.asarray(g.vtype)
    cam_idx = np.cumsum(vt == 0) - 1          # vertex id -> row of g.cams
    pt_idx = np.cumsum(vt == 1) - 1
    in1 = (np.asarray(g.obs_pt) < n_v1) & (np.asarray(g.obs_cam) < n_v1)

    def part(vsel, osel):
        ids = np.flatnonzero(vsel)
        return BAGraph(vt[ids], g.cams[cam_idx[ids[vt[ids] == 0]]], g.pts[pt_idx[ids[vt[ids] == 1]]],
                       np.asarray(g.obs_pt)[osel], np.asarray(g.obs_cam)[osel], g.z[osel], g.info[osel])
    v1 = np.arange(len(vt)) < n_v1
    first, rest = part(v1, in1), part(~v1, ~in1)
    full = BAGraph(vt, g.cams, g.pts, np.concatenate([first.obs_pt, rest.obs_pt]), np.concatenate([first.obs_cam, rest.obs_cam]),
                   np.concatenate([first.z, rest.z]), np.concatenate([first.info, rest.info]))
    return first, rest, full, cam_idx, pt_idx


@pytest.mark.parametrize("interleave", [False, True])
def test_append_equals_set_graph_of_the_concatenation(interleave):
    from slam_plus_plus_b200 import capi, graphs
    g = graphs.make_ba(40, 3000, 17, interleave_ids=interleave)
    n_v = len(g.vtype)
    n_v1 = int(0.6 * n_v) if interleave else 40 + 1700   # cameras first: the prefix holds all cameras and 1700 landmarks
    first, rest, full, cam_idx, pt_idx = _split(g, n_v1)
    assert first.n_obs > 0 and rest.n_obs > 0 and len(rest.vtype) > 0

    a = capi.Context(0)
    a.ba_set_graph(first)
    ra1 = a.ba_optimize(2, 0.0)
    a.ba_append_graph(rest)
    chi_a = a.ba_chi2()
    ra2 = a.ba_optimize(3, 0.0)
    cams_a, pts_a = a.ba_get_states()

    b = capi.Context(0)
    b.ba_set_graph(first)
    rb1 = b.ba_optimize(2, 0.0)
    cams_1, pts_1 = b.ba_get_states()
    b.ba_set_graph(full)
    cs, ps = full.cams[:, :6].copy(), full.pts.copy()
    cs[:first.n_cams], ps[:first.n_pts] = cams_1, pts_1   # prefix vertices come first within their type
    b.ba_set_states(cs, ps)
    chi_b = b.ba_chi2()
    rb2 = b.ba_optimize(3, 0.0)
    cams_b, pts_b = b.ba_get_states()

    assert ra1["chi2_final"] == rb1["chi2_final"]
    assert chi_a == chi_b
    assert ra2["chi2_final"] == rb2["chi2_final"] and ra2["n_iterations"] == rb2["n_iterations"]
    assert np.array_equal(cams_a, cams_b) and np.array_equal(pts_a, pts_b)
    # restore_initial: the old vertices go back to what spp_ba_set_graph saw, the new ones to what came with the append
    a.ba_restore_initial()
    c0, p0 = a.ba_get_states()
    assert np.array_equal(c0, full.cams[:, :6]) and np.array_equal(p0, full.pts)
    a.close()
    b.close()


def test_append_in_several_steps_and_bad_input():
    from slam_plus_plus_b200 import capi, graphs
    from slam_plus_plus_b200.sppio import BAGraph
    g = graphs.make_ba(30, 1500, 5, interleave_ids=True)
    n_v = len(g.vtype)
    cuts = [int(0.4 * n_v), int(0.7 * n_v), n_v]
    ctx = capi.Context(0)
    first, _, _, _, _ = _split(g, cuts[0])
    ctx.ba_set_graph(first)
    vt = np.asarray(g.vtype)
    cam_idx, pt_idx = np.cumsum(vt == 0) - 1, np.cumsum(vt == 1) - 1
    op, oc = np.asarray(g.obs_pt), np.asarray(g.obs_cam)
    done = (op < cuts[0]) & (oc < cuts[0])
    order = [np.flatnonzero(done)]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        ids = np.arange(lo, hi)
        osel = (op < hi) & (oc < hi) & ~done
        ctx.ba_append_graph(BAGraph(vt[ids], g.cams[cam_idx[ids[vt[ids] == 0]]], g.pts[pt_idx[ids[vt[ids] == 1]]], op[osel], oc[osel],
                                    g.z[osel], g.info[osel]))
        done |= osel
        order.append(np.flatnonzero(osel))
    order = np.concatenate(order)
    chi_inc = ctx.ba_chi2()
    ref = capi.Context(0)
    ref.ba_set_graph(BAGraph(vt, g.cams, g.pts, op[order], oc[order], g.z[order], g.info[order]))
    assert chi_inc == ref.ba_chi2()
    r1, r2 = ctx.ba_optimize(3, 0.0), ref.ba_optimize(3, 0.0)
    assert r1["chi2_final"] == r2["chi2_final"]
    # an observation that references a vertex that does not exist is refused, and the context says so afterwards
    with pytest.raises(capi.SppError):
        ctx.ba_append_graph(BAGraph(np.zeros(0, np.uint8), np.zeros((0, 11)), np.zeros((0, 3)), np.array([n_v + 5], np.uint64),
                                    np.array([0], np.uint64), np.zeros((1, 2)), np.eye(2)[None]))
    with pytest.raises(capi.SppError):
        ctx.ba_chi2()
    ctx.close()
    ref.close()
    # a context without a graph cannot be appended to
    c2 = capi.Context(0)
    with pytest.raises(capi.SppError):
        c2.ba_append_graph(first)
    c2.close()
