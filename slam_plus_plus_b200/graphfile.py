"""Graph-file ingest: the reference's text formats -> the SoA arrays the C ABI takes.

Replaces, for the BA, SE(2) and SE(3) formats, the reference's parser front end:
    CParserTemplate / CParserBase            include/slam/Parser.h:1137-..., TEdge2D :166-205, TEdgeP2C3D :625-664
    CVertex2DParsePrimitive                  include/slam_app/ParsePrimitives.h:409-460   VERTEX2 / VERTEX_SE2 / VERTEX
    CEdge2DParsePrimitive                    include/slam_app/ParsePrimitives.h:75-258    EDGE2 / EDGE_SE2 / EDGE / ODOMETRY
    CVertex3DParsePrimitive                  include/slam_app/ParsePrimitives.h:740-787   VERTEX3 / VERTEX_SE3
    CEdge3DParsePrimitive                    include/slam_app/ParsePrimitives.h:463-538   EDGE3 / EDGE_SE3
    CEdge3DParsePrimitiveAxisAngle           include/slam_app/ParsePrimitives.h:553-610   EDGE3:AXISANGLE / EDGE_SE3:AXISANGLE
    CVertexXYZParsePrimitive                 include/slam_app/ParsePrimitives.h:809-856   VERTEX_XYZ
    CVertexCam3DParsePrimitive               include/slam_app/ParsePrimitives.h:861-927   VERTEX_CAM
    CEdgeP2C3DParsePrimitive                 include/slam_app/ParsePrimitives.h:1123-1183 EDGE_PROJECT_P2MC / EDGE_P2MC / EDGE_P2C
including what the parser does to the numbers before the optimizer sees them:
  * VERTEX_CAM holds the camera-to-world pose as t, quaternion (x y z w); the state is the INVERSE pose
    [q^-1 (-t), axis-angle(q^-1)] (ParsePrimitives.h:898-912, Quat_to_AxisAngle 3DSolverBase.h:556-649), and the
    distortion coefficient is multiplied by (fx + fy) / 2 (TVertexCam3D, Parser.h:512-519);
  * 2D edges written with from > to (Manhattan datasets) are inverted: ids swapped, measurement replaced by
    Absolute_to_Relative(z, 0) (2DSolverBase.h:373-430) and the information matrix read in the "french" order
    |0 1 5; . 2 4; . . 3| unless its zeros say it is in the usual upper-triangular order (ParsePrimitives.h:171-246);
  * VERTEX3 / EDGE3 hold rotations as roll, pitch, yaw; the parser builds R = Rz(yaw) Ry(pitch) Rx(roll) and converts it
    to axis-angle through a quaternion (v_RotMatrix_to_AxisAngle, 3DSolverBase.h:357-366: Eigen's matrix -> quaternion,
    then Quat_to_AxisAngle); EDGE3 lines with from >= to are reported and DROPPED (ParsePrimitives.h:496-533), while
    EDGE3:AXISANGLE lines are taken as they are;
  * information matrices come as upper triangles, row by row (21 values for the 6x6 of TEdge3D, Parser.h:352-371).
Pinned against the reference's own parser: tests/golden/parse_ref.npz (made by the reference-parser driver of the test
infrastructure), tests/test_graphfile_cpu.py. Host-side plumbing only: no numerics of the hot path live here.
"""
from __future__ import annotations

import math

import numpy as np

from .sppio import GRAPH_SE2, GRAPH_SE3, BAGraph, PoseGraph

_V2 = ("VERTEX2", "VERTEX_SE2", "VERTEX")
_E2 = ("EDGE2", "EDGE_SE2", "EDGE", "ODOMETRY")
_P2C = ("EDGE_PROJECT_P2MC", "EDGE_P2MC", "EDGE_P2C")
_V3 = ("VERTEX3", "VERTEX_SE3")
_E3 = ("EDGE3", "EDGE_SE3")
_E3AA = ("EDGE3:AXISANGLE", "EDGE_SE3:AXISANGLE")


def _quat_to_axis_angle(w, x, y, z):
    """C3DJacobians::Quat_to_AxisAngle, include/slam/3DSolverBase.h:556-649 (the atan / atan2 variant)"""
    f_norm = math.sqrt(x * x + y * y + z * z)
    f_abs_w = abs(w)
    f_abs_half = math.atan(f_norm / f_abs_w) if f_abs_w > 1e-3 else math.atan2(f_norm, f_abs_w)
    f_half = math.copysign(f_abs_half, w)
    if f_norm < 1e-12:
        return 2.0 * x, 2.0 * y, 2.0 * z
    f_s = f_half * 2 / f_norm
    return x * f_s, y * f_s, z * f_s


def _quat_rotate(w, x, y, z, v):
    """Eigen::QuaternionBase::_transformVector"""
    ux, uy, uz = y * v[2] - z * v[1], z * v[0] - x * v[2], x * v[1] - y * v[0]
    ux, uy, uz = ux + ux, uy + uy, uz + uz
    return (v[0] + w * ux + (y * uz - z * uy), v[1] + w * uy + (z * ux - x * uz), v[2] + w * uz + (x * uy - y * ux))


def _rotmat_to_quat(m):
    """Eigen::Quaternion(Matrix3) (Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other, 3, 3>: the
    trace / largest-diagonal branches of Shoemake's method) -> (w, x, y, z)"""
    t = m[0][0] + m[1][1] + m[2][2]
    if t > 0:
        t = math.sqrt(t + 1.0)
        w = .5 * t
        t = .5 / t
        return w, (m[2][1] - m[1][2]) * t, (m[0][2] - m[2][0]) * t, (m[1][0] - m[0][1]) * t
    i = 0
    if m[1][1] > m[0][0]:
        i = 1
    if m[2][2] > m[i][i]:
        i = 2
    j = (i + 1) % 3
    k = (j + 1) % 3
    t = math.sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0)
    q = [0.0, 0.0, 0.0]
    q[i] = .5 * t
    t = .5 / t
    w = (m[k][j] - m[j][k]) * t
    q[j] = (m[j][i] + m[i][j]) * t
    q[k] = (m[k][i] + m[i][k]) * t
    return w, q[0], q[1], q[2]


def _rpy_to_axis_angle(roll, pitch, yaw):
    """the RPY -> axis-angle conversion of CVertex3DParsePrimitive / CEdge3DParsePrimitive (ParsePrimitives.h:509-521,
    765-777): R = Rz(yaw) Ry(pitch) Rx(roll), then C3DJacobians::v_RotMatrix_to_AxisAngle"""
    cx, sx = math.cos(yaw), math.sin(yaw)
    cy, sy = math.cos(pitch), math.sin(pitch)
    cz, sz = math.cos(roll), math.sin(roll)
    q = ((cy * cx, -cz * sx + sz * sy * cx, sz * sx + cz * sy * cx),
         (cy * sx, cz * cx + sz * sy * sx, -sz * cx + cz * sy * sx),
         (-sy, sz * cy, cz * cy))
    return _quat_to_axis_angle(*_rotmat_to_quat(q))


def _upper21_to_full(u):
    """TEdge3D's 6x6 information matrix from its 21 upper-triangular values, row-major (Parser.h:362-369)"""
    m = [[0.0] * 6 for _ in range(6)]
    k = 0
    for i in range(6):
        for j in range(i, 6):
            m[i][j] = m[j][i] = u[k]
            k += 1
    return tuple(v for row in m for v in row)


def _clamp_angle_2pi(a):
    return math.fmod(a, 2 * math.pi) if math.isfinite(a) else 0.0


def _se2_absolute_to_relative(v1, v2):
    """C2DJacobians::Absolute_to_Relative (value), include/slam/2DSolverBase.h:373-430"""
    de, dn, da = v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]
    o = -v1[2]
    co, so = math.cos(o), math.sin(o)
    return co * de - so * dn, so * de + co * dn, _clamp_angle_2pi(da)


class ParsedGraph:
    """What the reference's parse loop would receive, in file order."""

    def __init__(self):
        self.vertex2d, self.vertex_xyz, self.vertex_cam, self.vertex3d = [], [], [], []  # (id, state...)
        self.edge2d, self.edge_p2c, self.edge3d = [], [], []          # (id0, id1, z..., info row-major...)
        self.n_ignored = 0
        self.n_switched = 0  # EDGE3 lines with from >= to, which the reference's parser reports and drops


def _bulk_numbers(path, lines, line_nos, n_col):
    """the records of one type that need no per-record arithmetic (landmarks, projections: all but a few thousand lines of
    a BA file), converted in one pass; numbers beyond the first n_col of a line are ignored, as sscanf does"""
    if not lines:
        return np.zeros((0, n_col))
    a = np.array(" ".join(lines).split(), np.float64)
    if a.size == len(lines) * n_col:
        return a.reshape(-1, n_col)
    rows = np.empty((len(lines), n_col))  # ragged: the slow way, which also finds the truncated line
    for k, line in enumerate(lines):
        tok = line.split()
        if len(tok) < n_col:
            raise ValueError(f"{path}: line {line_nos[k] + 1}: line is truncated")
        rows[k] = [float(x) for x in tok[:n_col]]
    return rows


def parse(path) -> ParsedGraph:
    out = ParsedGraph()
    xyz_lines, xyz_nos, p2c_lines, p2c_nos = [], [], [], []
    with open(path) as f:
        for line_no, line in enumerate(f):
            head = line.split(None, 1)
            if not head or head[0].startswith("#") or head[0].startswith("%"):
                continue
            name = head[0].upper()
            if name == "VERTEX_XYZ":       # bulk records: converted after the loop
                xyz_lines.append(head[1] if len(head) > 1 else "")
                xyz_nos.append(line_no)
                continue
            if name in _P2C:
                p2c_lines.append(head[1] if len(head) > 1 else "")
                p2c_nos.append(line_no)
                continue
            a = head[1].split() if len(head) > 1 else []
            try:
                if name in _V2:
                    out.vertex2d.append((int(a[0]), float(a[1]), float(a[2]), float(a[3])))
                elif name == "VERTEX_CAM":
                    v = [float(x) for x in a[1:13]]
                    if len(v) != 12:
                        raise ValueError
                    qx, qy, qz, qw = v[3], v[4], v[5], v[6]
                    n = math.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
                    qx, qy, qz, qw = qx / n, qy / n, qz / n, qw / n      # quat.normalize()
                    n2 = qw * qw + qx * qx + qy * qy + qz * qz           # quat.inverse() = conjugate / squaredNorm
                    iw, ix, iy, iz = qw / n2, -qx / n2, -qy / n2, -qz / n2
                    c = _quat_rotate(iw, ix, iy, iz, (-v[0], -v[1], -v[2]))
                    ax = _quat_to_axis_angle(iw, ix, iy, iz)
                    d = v[11] * (.5 * (v[7] + v[8]))  # the distortion is scaled by the focal length internally (Parser.h:512-519)
                    out.vertex_cam.append((int(a[0]), c[0], c[1], c[2], ax[0], ax[1], ax[2], v[7], v[8], v[9], v[10], d))
                elif name in _E2:
                    i0, i1 = int(a[0]), int(a[1])
                    z = [float(x) for x in a[2:5]]
                    m = [float(x) for x in a[5:11]]
                    if len(m) != 6:
                        raise ValueError
                    if i0 < i1:
                        up = m
                    else:  # descending edge: invert it (ParsePrimitives.h:171-246)
                        g2o = False
                        if abs(m[0]) < 1e-5 or abs(m[2]) < 1e-5 or abs(m[3]) < 1e-5:
                            if abs(m[0]) > 1e-5 and abs(m[3]) > 1e-5 and abs(m[5]) > 1e-5:
                                g2o = True
                        up = m if g2o else [m[0], m[1], m[5], m[2], m[4], m[3]]
                        i0, i1 = i1, i0
                        z = list(_se2_absolute_to_relative(z, (0.0, 0.0, 0.0)))
                    info = (up[0], up[1], up[2], up[1], up[3], up[4], up[2], up[4], up[5])
                    out.edge2d.append((i0, i1, z[0], z[1], z[2]) + info)
                elif name in _V3:
                    v = [float(x) for x in a[1:7]]
                    if len(v) != 6:
                        raise ValueError
                    out.vertex3d.append((int(a[0]), v[0], v[1], v[2]) + tuple(_rpy_to_axis_angle(v[3], v[4], v[5])))
                elif name in _E3 or name in _E3AA:
                    i0, i1 = int(a[0]), int(a[1])
                    z = [float(x) for x in a[2:8]]
                    m = [float(x) for x in a[8:29]]
                    if len(m) != 21:
                        raise ValueError
                    if name in _E3:
                        if not i0 < i1:
                            out.n_switched += 1
                            continue
                        z[3:6] = _rpy_to_axis_angle(z[3], z[4], z[5])
                    out.edge3d.append((i0, i1) + tuple(z) + _upper21_to_full(m))
                else:
                    out.n_ignored += 1  # CONSISTENCY_MARKER and the primitives of other problem types
            except (ValueError, IndexError):
                raise ValueError(f"{path}: line {line_no + 1}: line is truncated") from None
    try:
        out.vertex_xyz = _bulk_numbers(path, xyz_lines, xyz_nos, 4)              # (id, x, y, z)
        e = _bulk_numbers(path, p2c_lines, p2c_nos, 7)                           # ids, z, upper triangle of the 2x2 information
    except ValueError as err:
        if "truncated" in str(err):
            raise
        raise ValueError(f"{path}: a landmark or projection line does not parse as numbers") from None
    out.edge_p2c = e[:, [0, 1, 2, 3, 4, 5, 5, 6]]
    return out


def load_ba(path) -> BAGraph:
    """A BA file (VERTEX_CAM / VERTEX_XYZ / EDGE_PROJECT_P2MC) as the arrays spp_ba_set_graph takes. Vertex ids must be
    0 .. n-1 (any interleaving of cameras and points); edges keep their file order (= edge insertion order)."""
    p = parse(path)
    cam = np.array(p.vertex_cam, np.float64).reshape(-1, 12)
    xyz = np.asarray(p.vertex_xyz, np.float64).reshape(-1, 4)
    ids = np.concatenate([cam[:, 0], xyz[:, 0]]).astype(np.int64)
    order = np.argsort(ids, kind="stable")
    if not np.array_equal(ids[order], np.arange(len(ids))):
        raise ValueError(f"{path}: vertex ids must be 0 .. n-1 without gaps or repeats")
    vtype = (order >= len(cam)).astype(np.int64)          # per vertex id: 0 camera, 1 point
    cams = cam[np.argsort(cam[:, 0], kind="stable"), 1:].copy()
    pts = xyz[np.argsort(xyz[:, 0], kind="stable"), 1:].copy()
    e = np.asarray(p.edge_p2c, np.float64).reshape(-1, 8)
    return BAGraph(vtype, cams, pts, e[:, 0].astype(np.int64), e[:, 1].astype(np.int64), e[:, 2:4].copy(),
                   e[:, 4:8].reshape(-1, 2, 2).copy())


def load_se2(path) -> PoseGraph:
    """A 2D pose graph (VERTEX2 / EDGE2 ...). Poses without a VERTEX line are initialised from the first edge that
    reaches them, as the reference does (CEdgePose2D constructor, SE2_Types.h:225-238: Relative_to_Absolute)."""
    p = parse(path)
    e = np.array(p.edge2d, np.float64).reshape(-1, 14)
    n = int(max([v[0] for v in p.vertex2d] + ([int(e[:, :2].max())] if len(e) else []) + [-1])) + 1
    poses = np.zeros((n, 3))
    known = np.zeros(n, bool)
    for v in p.vertex2d:
        poses[v[0]] = v[1:]
        known[v[0]] = True
    if n and not known[0]:
        known[0] = True  # the first vertex starts at the origin
    for k in range(len(e)):
        a, b = int(e[k, 0]), int(e[k, 1])
        if known[a] and not known[b]:
            c, s = math.cos(poses[a, 2]), math.sin(poses[a, 2])
            poses[b] = (poses[a, 0] + c * e[k, 2] - s * e[k, 3], poses[a, 1] + s * e[k, 2] + c * e[k, 3],
                        _clamp_angle_2pi(poses[a, 2] + e[k, 4]))
            known[b] = True
    if not known.all():
        raise ValueError(f"{path}: pose {int(np.flatnonzero(~known)[0])} is neither given nor reachable from an initialised pose")
    return PoseGraph(GRAPH_SE2, poses, e[:, 0].astype(np.int64), e[:, 1].astype(np.int64), e[:, 2:5].copy(),
                     e[:, 5:14].reshape(-1, 3, 3).copy())


def _quat_mul(a, b):
    """Eigen quaternion product, (w, x, y, z)"""
    return (a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3], a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1])


def _se3_relative_to_absolute(v1, v2):
    """C3DJacobians::Relative_to_Absolute (value), include/slam/3DSolverBase.h:807-850 (the quaternion branch)"""
    q1, q2 = _axis_angle_to_quat(v1[3:6]), _axis_angle_to_quat(v2[3:6])
    t = _quat_rotate(q1[0], q1[1], q1[2], q1[3], v2[0:3])
    return (v1[0] + t[0], v1[1] + t[1], v1[2] + t[2]) + tuple(_quat_to_axis_angle(*_quat_mul(q1, q2)))


def load_se3(path) -> PoseGraph:
    """A 3D pose graph (VERTEX3 / EDGE3 / EDGE3:AXISANGLE). Poses without a VERTEX line are initialised from the first
    edge that reaches them (CEdgePose3D constructor, SE3_Types.h: Relative_to_Absolute)."""
    p = parse(path)
    e = np.array(p.edge3d, np.float64).reshape(-1, 44)
    n = int(max([v[0] for v in p.vertex3d] + ([int(e[:, :2].max())] if len(e) else []) + [-1])) + 1
    poses = np.zeros((n, 6))
    known = np.zeros(n, bool)
    for v in p.vertex3d:
        poses[v[0]] = v[1:]
        known[v[0]] = True
    if n and not known[0]:
        known[0] = True
    for k in range(len(e)):
        a, b = int(e[k, 0]), int(e[k, 1])
        if known[a] and not known[b]:
            poses[b] = _se3_relative_to_absolute(poses[a], e[k, 2:8])
            known[b] = True
    if not known.all():
        raise ValueError(f"{path}: pose {int(np.flatnonzero(~known)[0])} is neither given nor reachable from an initialised pose")
    return PoseGraph(GRAPH_SE3, poses, e[:, 0].astype(np.int64), e[:, 1].astype(np.int64), e[:, 2:8].copy(),
                     e[:, 8:44].reshape(-1, 6, 6).copy())


def peek(path, max_lines: int = 1000) -> str:
    """What kind of graph a file holds -- "ba", "se2" or "se3" -- from the tokens of its first lines: the decision the
    reference's TDatasetPeeker takes before slam_app picks a system type (include/slam_app/Main.h:830-900: it parses the
    first 1000 lines and records which kinds of vertices and edges occur)."""
    seen = set()
    with open(path) as f:
        for k, line in enumerate(f):
            if k >= max_lines:
                break
            head = line.split(None, 1)
            if head and not head[0].startswith(("#", "%")):
                seen.add(head[0].upper())
    if seen & (set(_P2C) | {"VERTEX_CAM", "VERTEX_XYZ"}):
        return "ba"
    if seen & (set(_V3) | set(_E3) | set(_E3AA)):
        return "se3"
    if seen & (set(_V2) | set(_E2)):
        return "se2"
    raise ValueError(f"{path}: no vertex or edge token of a supported graph type in the first {max_lines} lines")


def load(path):
    """load_ba / load_se2 / load_se3, chosen by peek()"""
    return {"ba": load_ba, "se2": load_se2, "se3": load_se3}[peek(path)](path)


# ---- writers (synthetic graphs -> the reference's text formats) ------------------------------------------------------

def _axis_angle_to_quat(a):
    """C3DJacobians::AxisAngle_to_Quat, include/slam/3DSolverBase.h:476-519 -> (w, x, y, z)"""
    ang = math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])
    if ang < 1e-12:
        return 1.0, a[0] * .5, a[1] * .5, a[2] * .5
    s = math.sin(ang * .5) / ang
    return math.cos(ang * .5), a[0] * s, a[1] * s, a[2] * s


def write_ba(path, g: BAGraph):
    """Writes the file whose parse gives g back (camera states are stored inverted in the file, as in the datasets)."""
    ci = pi = 0
    with open(path, "w") as f:
        for vid, t in enumerate(g.vtype):
            if t == 0:
                c = g.cams[ci]
                ci += 1
                w, x, y, z = _axis_angle_to_quat(c[3:6])     # state rotation q^-1 -> file rotation q
                tw = _quat_rotate(w, -x, -y, -z, (-c[0], -c[1], -c[2]))  # t = -(q c) with q = conj(state quaternion)
                f.write("VERTEX_CAM %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n"
                        % ((vid,) + tuple(tw) + (-x, -y, -z, w) + tuple(c[6:10]) + (c[10] / (.5 * (c[6] + c[7])),)))
            else:
                q = g.pts[pi]
                pi += 1
                f.write("VERTEX_XYZ %d %.17g %.17g %.17g\n" % (vid, q[0], q[1], q[2]))
        for k in range(len(g.obs_pt)):
            m = g.info[k]
            f.write("EDGE_PROJECT_P2MC %d %d %.17g %.17g %.17g %.17g %.17g\n"
                    % (g.obs_pt[k], g.obs_cam[k], g.z[k, 0], g.z[k, 1], m[0, 0], m[0, 1], m[1, 1]))


def write_se2(path, g: PoseGraph, with_vertices: bool = True):
    with open(path, "w") as f:
        if with_vertices:
            for i, v in enumerate(g.poses):
                f.write("VERTEX2 %d %.17g %.17g %.17g\n" % (i, v[0], v[1], v[2]))
        for k in range(len(g.e_from)):
            m = g.info[k]
            f.write("EDGE2 %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n"
                    % (g.e_from[k], g.e_to[k], g.z[k, 0], g.z[k, 1], g.z[k, 2], m[0, 0], m[0, 1], m[0, 2], m[1, 1], m[1, 2], m[2, 2]))


def _axis_angle_to_rpy(a):
    """inverse of _rpy_to_axis_angle away from pitch = +-pi/2: R = Rz(yaw) Ry(pitch) Rx(roll)"""
    w, x, y, z = _axis_angle_to_quat(a)
    r00, r10, r20 = 1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)
    r21, r22 = 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)
    return math.atan2(r21, r22), math.atan2(-r20, math.hypot(r00, r10)), math.atan2(r10, r00)


def write_se3(path, g: PoseGraph, with_vertices: bool = True, axis_angle_edges: bool = True):
    """VERTEX3 lines carry roll, pitch, yaw (the only 3D vertex format the reference reads); edges are written as
    EDGE3:AXISANGLE (exact round trip) or, with axis_angle_edges=False, as EDGE3 with roll, pitch, yaw."""
    iu = [(i, j) for i in range(6) for j in range(i, 6)]
    with open(path, "w") as f:
        if with_vertices:
            for i, v in enumerate(g.poses):
                f.write(("VERTEX3 %d" + " %.17g" * 6 + "\n") % ((i, v[0], v[1], v[2]) + _axis_angle_to_rpy(v[3:6])))
        for k in range(len(g.e_from)):
            z = tuple(g.z[k, :3]) + (tuple(g.z[k, 3:6]) if axis_angle_edges else _axis_angle_to_rpy(g.z[k, 3:6]))
            f.write((("EDGE3:AXISANGLE" if axis_angle_edges else "EDGE3") + " %d %d" + " %.17g" * 27 + "\n")
                    % ((g.e_from[k], g.e_to[k]) + z + tuple(g.info[k][i, j] for i, j in iu)))
