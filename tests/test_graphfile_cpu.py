"""Graph-file ingest (slam_plus_plus_b200/graphfile.py) against the UNMODIFIED reference's own parser: the golden
tests/golden/parse_ref.npz holds what CParserTemplate + the parse primitives of include/slam_app/ParsePrimitives.h hand
to the parse loop for the committed text files (tests/golden/make_golden_parse.py, oracle/ref_driver_parse.cpp)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_err
from slam_plus_plus_b200 import graphfile, graphs


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLDEN, "parse_ref.npz"))


def test_ba_file_matches_reference_parser(ref):
    p = graphfile.parse(os.path.join(GOLDEN, "parse_ba.txt"))
    cam = np.array(p.vertex_cam).reshape(-1, 12)
    r_cam = ref["parse_ba.vertex_cam"].reshape(-1, 12)
    assert np.array_equal(cam[:, 0], r_cam[:, 0])
    # the camera state is the inverted pose: q^-1 (-t) and the axis-angle of q^-1 -- libm-level agreement
    assert np.max(np.abs(cam[:, 1:] - r_cam[:, 1:])) <= 1e-15 * max(1.0, np.abs(r_cam).max())
    assert np.array_equal(np.array(p.vertex_xyz).reshape(-1, 4), ref["parse_ba.vertex_xyz"].reshape(-1, 4))
    assert np.array_equal(np.array(p.edge_p2c).reshape(-1, 8), ref["parse_ba.edge_p2c"].reshape(-1, 8))


def test_ba_file_round_trip():
    g = graphs.ba_shape("tiny", interleave_ids=True, shuffle_edges=True, distortion=-0.03)
    g.info[:, 0, 1] = g.info[:, 1, 0] = 0.25
    h = graphfile.load_ba(os.path.join(GOLDEN, "parse_ba.txt"))
    assert np.array_equal(h.vtype, g.vtype) and np.array_equal(h.obs_pt, g.obs_pt) and np.array_equal(h.obs_cam, g.obs_cam)
    assert np.array_equal(h.pts, g.pts) and np.array_equal(h.z, g.z) and np.array_equal(h.info, g.info)
    assert rel_err(h.cams, g.cams) < 1e-14  # pose -> file -> pose goes through a quaternion twice


@pytest.mark.parametrize("name", ["parse_se2", "parse_se2_mixed"])
def test_se2_file_matches_reference_parser(ref, name):
    p = graphfile.parse(os.path.join(GOLDEN, name + ".txt"))
    assert np.array_equal(np.array(p.vertex2d).reshape(-1, 4), ref[name + ".vertex2d"].reshape(-1, 4))
    e, r = np.array(p.edge2d).reshape(-1, 14), ref[name + ".edge2d"].reshape(-1, 14)
    assert e.shape == r.shape
    assert np.array_equal(e[:, :2], r[:, :2])                      # ids, descending edges swapped
    assert np.max(np.abs(e[:, 2:5] - r[:, 2:5])) <= 1e-15          # measurements, inverted where needed
    assert np.array_equal(e[:, 5:], r[:, 5:])                      # information matrices, both storage orders


def test_se2_load_initialises_missing_poses(tmp_path):
    g = graphs.make_manhattan(30, 8, seed=2)
    path = str(tmp_path / "g.txt")
    graphfile.write_se2(path, g, with_vertices=False)
    h = graphfile.load_se2(path)
    assert h.poses.shape == g.poses.shape and np.array_equal(h.e_from, g.e_from)
    assert np.allclose(h.poses[0], 0)
    with open(path, "a") as f:
        f.write("EDGE2 40 41 1 0 0 1 0 0 1 0 1\n")  # an island
    with pytest.raises(ValueError):
        graphfile.load_se2(path)


def test_truncated_line_is_an_error(tmp_path):
    path = str(tmp_path / "bad.txt")
    with open(path, "w") as f:
        f.write("VERTEX_XYZ 0 1 2\n")
    with pytest.raises(ValueError):
        graphfile.parse(path)


@pytest.mark.parametrize("name", ["parse_se3", "parse_se3_mixed"])
def test_se3_file_matches_reference_parser(ref, name):
    p = graphfile.parse(os.path.join(GOLDEN, name + ".txt"))
    v, r = np.array(p.vertex3d).reshape(-1, 7), ref[name + ".vertex3d"].reshape(-1, 7)
    assert v.shape == r.shape and np.array_equal(v[:, :4], r[:, :4])
    # roll-pitch-yaw -> rotation matrix -> quaternion -> axis-angle, all three non-trace branches in the mixed file
    assert np.max(np.abs(v[:, 4:] - r[:, 4:])) <= 2e-15
    e, q = np.array(p.edge3d).reshape(-1, 44), ref[name + ".edge3d"].reshape(-1, 44)
    assert e.shape == q.shape                                      # the switched EDGE3 line is dropped, as the reference does
    assert np.array_equal(e[:, :5], q[:, :5]) and np.max(np.abs(e[:, 5:8] - q[:, 5:8])) <= 2e-15
    assert np.array_equal(e[:, 8:], q[:, 8:])                      # 21 upper-triangular values -> symmetric 6x6
    assert p.n_switched == (1 if name == "parse_se3_mixed" else 0)


def test_se3_file_round_trip(tmp_path):
    g = graphs.make_sphere(5, 8, seed=5, radius=5.0)
    h = graphfile.load_se3(os.path.join(GOLDEN, "parse_se3.txt"))
    assert np.array_equal(h.e_from, g.e_from) and np.array_equal(h.e_to, g.e_to)
    assert np.array_equal(h.z, g.z) and np.array_equal(h.info, g.info)
    assert np.max(np.abs(h.poses - g.poses)) < 1e-14               # vertices go through roll-pitch-yaw
    # roll-pitch-yaw edges and no vertex lines: poses are chained from the edges (Relative_to_Absolute)
    path = str(tmp_path / "rpy.txt")
    graphfile.write_se3(path, g, with_vertices=False, axis_angle_edges=False)
    k = graphfile.load_se3(path)
    assert np.max(np.abs(k.z - g.z)) < 1e-14 and np.allclose(k.poses[0], 0)
    assert k.poses.shape == g.poses.shape and np.all(np.isfinite(k.poses))


def test_ba_bulk_records_ragged_and_bad(tmp_path):
    """landmark / projection lines go through a one-pass conversion: trailing extra numbers are ignored (sscanf reads what
    it needs), a short line is reported with its line number, a non-number is an error"""
    path = str(tmp_path / "g.txt")
    with open(path, "w") as f:
        f.write("VERTEX_CAM 0 0 0 0 0 0 0 1 500 500 320 240 0\n")
        f.write("VERTEX_XYZ 1 0.5 0.25 4.0 99\n")                       # an extra number
        f.write("VERTEX_XYZ 2 -0.5 0.25 5.0\n")
        f.write("EDGE_PROJECT_P2MC 1 0 380.0 270.0 1 0 1\nEDGE_P2C 2 0 270.0 265.0 2 0.5 3 7 7\n")
    g = graphfile.load_ba(path)
    assert np.array_equal(g.vtype, [0, 1, 1]) and np.array_equal(g.pts, [[0.5, 0.25, 4.0], [-0.5, 0.25, 5.0]])
    assert np.array_equal(g.obs_pt, [1, 2]) and np.array_equal(g.obs_cam, [0, 0])
    assert np.array_equal(g.info[1], [[2, 0.5], [0.5, 3]]) and np.array_equal(g.z, [[380.0, 270.0], [270.0, 265.0]])
    with open(path, "a") as f:
        f.write("EDGE_PROJECT_P2MC 2 0 1.0 2.0 1 0\n")                  # line 6 is short
    with pytest.raises(ValueError, match="line 6"):
        graphfile.parse(path)
    with open(path, "w") as f:
        f.write("VERTEX_XYZ 1 0.5 abc 4.0\n")
    with pytest.raises(ValueError):
        graphfile.parse(path)


def test_peek_and_load_pick_the_graph_type(tmp_path):
    from slam_plus_plus_b200 import sppio
    kinds = {"parse_ba": "ba", "parse_se2": "se2", "parse_se2_mixed": "se2", "parse_se3": "se3", "parse_se3_mixed": "se3"}
    for name, kind in kinds.items():
        assert graphfile.peek(os.path.join(GOLDEN, name + ".txt")) == kind
    assert isinstance(graphfile.load(os.path.join(GOLDEN, "parse_ba.txt")), sppio.BAGraph)
    g3 = graphfile.load(os.path.join(GOLDEN, "parse_se3.txt"))
    assert g3.kind == sppio.GRAPH_SE3 and g3.poses.shape[1] == 6
    path = str(tmp_path / "x.txt")
    with open(path, "w") as f:
        f.write("# nothing here\nEQUIV 1 2\n")
    with pytest.raises(ValueError):
        graphfile.peek(path)


def test_round_trips_on_random_graphs(tmp_path):
    """write -> parse round trips on seeded random graphs of the three kinds: ids, measurements and information matrices
    exact (17 significant digits), states that go through a quaternion or roll-pitch-yaw to rounding"""
    from slam_plus_plus_b200 import sppio
    rng = np.random.default_rng(1234)
    for trial in range(6):
        # SE(2): random information matrices (symmetric), descending loop closures written as they are in memory
        g2 = graphs.make_manhattan(int(rng.integers(5, 60)), int(rng.integers(0, 20)), seed=int(rng.integers(1 << 30)))
        a = rng.normal(size=(len(g2.e_from), 3, 3))
        g2.info = a @ a.transpose(0, 2, 1) + 3 * np.eye(3)
        asc = g2.e_from < g2.e_to                                   # descending edges are inverted by the parser (tested above)
        g2 = sppio.PoseGraph(g2.kind, g2.poses, g2.e_from[asc], g2.e_to[asc], g2.z[asc], g2.info[asc])
        p2 = str(tmp_path / f"a{trial}.txt")
        graphfile.write_se2(p2, g2)
        h2 = graphfile.load(p2)
        assert np.array_equal(h2.poses, g2.poses) and np.array_equal(h2.e_from, g2.e_from) and np.array_equal(h2.e_to, g2.e_to)
        assert np.array_equal(h2.z, g2.z) and np.array_equal(h2.info, g2.info)
        # SE(3)
        g3 = graphs.make_sphere(int(rng.integers(2, 6)), int(rng.integers(3, 9)), seed=int(rng.integers(1 << 30)), radius=5.0)
        b = rng.normal(size=(len(g3.e_from), 6, 6))
        g3.info = b @ b.transpose(0, 2, 1) + 6 * np.eye(6)
        p3 = str(tmp_path / f"b{trial}.txt")
        graphfile.write_se3(p3, g3)
        h3 = graphfile.load(p3)
        assert np.array_equal(h3.e_from, g3.e_from) and np.array_equal(h3.z, g3.z) and np.array_equal(h3.info, g3.info)
        R, Rh = graphs._axis_angle_to_rotmat(g3.poses[:, 3:]), graphs._axis_angle_to_rotmat(h3.poses[:, 3:])
        assert np.array_equal(h3.poses[:, :3], g3.poses[:, :3]) and np.abs(R - Rh).max() < 1e-14   # same rotations
        # BA
        gb = graphs.make_ba(int(rng.integers(3, 9)), int(rng.integers(10, 80)), int(rng.integers(1 << 30)),
                            interleave_ids=bool(trial & 1), shuffle_edges=bool(trial & 2), distortion=float(rng.normal(0, .05)))
        pb = str(tmp_path / f"c{trial}.txt")
        graphfile.write_ba(pb, gb)
        hb = graphfile.load(pb)
        assert np.array_equal(hb.vtype, gb.vtype) and np.array_equal(hb.obs_pt, gb.obs_pt) and np.array_equal(hb.obs_cam, gb.obs_cam)
        assert np.array_equal(hb.pts, gb.pts) and np.array_equal(hb.z, gb.z) and np.array_equal(hb.info, gb.info)
        assert rel_err(hb.cams, gb.cams) < 1e-13
