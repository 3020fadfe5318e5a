"""Binary graph / dump containers shared by the product's Python host layer and the oracle tools.

Graph file ``SPPGRAF1`` (all little endian, 8-byte words)::

    char[8]  "SPPGRAF1"
    u64      kind        0 = BA (cameras + points), 1 = SE(2) pose graph, 2 = SE(3) pose graph
    u64      n_vertices
    u64      n_edges
    BA only: u64 vtype[n_vertices]   0 = camera (11 doubles: t, axis-angle, fx fy cx cy d), 1 = point (3 doubles)
    f64      vdata[...]              vertex states in id order
    u64      e0[n_edges]             BA: point vertex id;  pose graphs: "from" vertex
    u64      e1[n_edges]             BA: camera vertex id; pose graphs: "to" vertex
    f64      z[n_edges * zdim]       zdim = 2 / 3 / 6
    f64      info[n_edges * zdim^2]  full symmetric information matrices

The vertex/edge semantics follow the reference's text formats (``VERTEX_CAM`` / ``VERTEX_XYZ`` /
``EDGE_PROJECT_P2MC``, ``EDGE2``/``EDGE_SE2``, ``EDGE3``/``EDGE_SE3:AXISANGLE``; /root/reference/data/Readme.txt)
after the parser's conversion to the internal representation (camera poses already inverted, distortion
already scaled; include/slam_app/ParsePrimitives.h:861-927).

Dump file: sequence of records ``char[32] name, u64 dtype (0 f64 / 1 u64), u64 count, payload``
(written by oracle/spp_dump.h).
"""
from __future__ import annotations

import dataclasses
import struct
from typing import Dict

import numpy as np

GRAPH_BA, GRAPH_SE2, GRAPH_SE3 = 0, 1, 2
_ZDIM = {GRAPH_BA: 2, GRAPH_SE2: 3, GRAPH_SE3: 6}
_MAGIC = b"SPPGRAF1"


@dataclasses.dataclass
class BAGraph:
    """Bundle-adjustment graph in the reference's internal representation.

    ``vtype[v]``: 0 camera / 1 point, indexed by vertex id (ids are shared, as in CFlatSystem).
    ``cams``: (C, 11) rows ``[t(3), axis_angle(3), fx, fy, cx, cy, d]`` in id order of the camera vertices.
    ``pts``: (P, 3). ``obs_pt`` / ``obs_cam``: *vertex ids* of each observation's point and camera,
    in edge insertion order. ``z``: (O, 2). ``info``: (O, 2, 2).
    """
    vtype: np.ndarray
    cams: np.ndarray
    pts: np.ndarray
    obs_pt: np.ndarray
    obs_cam: np.ndarray
    z: np.ndarray
    info: np.ndarray

    @property
    def n_cams(self) -> int:
        return int(self.cams.shape[0])

    @property
    def n_pts(self) -> int:
        return int(self.pts.shape[0])

    @property
    def n_obs(self) -> int:
        return int(self.obs_pt.shape[0])

    @property
    def n_vertices(self) -> int:
        return int(self.vtype.shape[0])

    def vertex_local_index(self) -> np.ndarray:
        """vertex id -> index into ``cams`` or ``pts`` (by type)."""
        is_cam = self.vtype == 0
        loc = np.empty(self.n_vertices, np.int64)
        loc[is_cam] = np.arange(int(is_cam.sum()))
        loc[~is_cam] = np.arange(int((~is_cam).sum()))
        return loc

    def vertex_offsets(self) -> np.ndarray:
        """Scalar offset of each vertex in the full state / dx vector (vertex id order: 6 per cam, 3 per point)."""
        dims = np.where(self.vtype == 0, 6, 3)
        return np.concatenate([[0], np.cumsum(dims)]).astype(np.int64)


@dataclasses.dataclass
class PoseGraph:
    """SE(2) / SE(3) pose graph: ``poses`` (N, 3|6) ``[t, angle|axis-angle]``; edges from->to."""
    kind: int
    poses: np.ndarray
    e_from: np.ndarray
    e_to: np.ndarray
    z: np.ndarray
    info: np.ndarray

    @property
    def dim(self) -> int:
        return _ZDIM[self.kind]


def write_graph(path: str, g) -> None:
    with open(path, "wb") as f:
        if isinstance(g, BAGraph):
            f.write(_MAGIC)
            f.write(struct.pack("<3Q", GRAPH_BA, g.n_vertices, g.n_obs))
            np.ascontiguousarray(g.vtype, np.uint64).tofile(f)
            vdata = np.empty(int(np.where(g.vtype == 0, 11, 3).sum()), np.float64)
            off = np.concatenate([[0], np.cumsum(np.where(g.vtype == 0, 11, 3))])[:-1]
            cam_off = off[g.vtype == 0]
            pt_off = off[g.vtype == 1]
            vdata[(cam_off[:, None] + np.arange(11)[None, :]).ravel()] = np.asarray(g.cams, np.float64).ravel()
            vdata[(pt_off[:, None] + np.arange(3)[None, :]).ravel()] = np.asarray(g.pts, np.float64).ravel()
            vdata.tofile(f)
            np.ascontiguousarray(g.obs_pt, np.uint64).tofile(f)
            np.ascontiguousarray(g.obs_cam, np.uint64).tofile(f)
            np.ascontiguousarray(g.z, np.float64).tofile(f)
            np.ascontiguousarray(g.info, np.float64).tofile(f)
        else:
            f.write(_MAGIC)
            f.write(struct.pack("<3Q", g.kind, g.poses.shape[0], g.e_from.shape[0]))
            np.ascontiguousarray(g.poses, np.float64).tofile(f)
            np.ascontiguousarray(g.e_from, np.uint64).tofile(f)
            np.ascontiguousarray(g.e_to, np.uint64).tofile(f)
            np.ascontiguousarray(g.z, np.float64).tofile(f)
            np.ascontiguousarray(g.info, np.float64).tofile(f)


def read_graph(path: str):
    with open(path, "rb") as f:
        if f.read(8) != _MAGIC:
            raise ValueError(f"{path}: not an SPPGRAF1 file")
        kind, nv, ne = struct.unpack("<3Q", f.read(24))
        zd = _ZDIM[kind]
        if kind == GRAPH_BA:
            vtype = np.fromfile(f, np.uint64, nv).astype(np.int64)
            dims = np.where(vtype == 0, 11, 3)
            vdata = np.fromfile(f, np.float64, int(dims.sum()))
            off = np.concatenate([[0], np.cumsum(dims)])[:-1]
            cams = vdata[(off[vtype == 0][:, None] + np.arange(11)[None, :])].reshape(-1, 11)
            pts = vdata[(off[vtype == 1][:, None] + np.arange(3)[None, :])].reshape(-1, 3)
            e0 = np.fromfile(f, np.uint64, ne).astype(np.int64)
            e1 = np.fromfile(f, np.uint64, ne).astype(np.int64)
            z = np.fromfile(f, np.float64, ne * zd).reshape(ne, zd)
            info = np.fromfile(f, np.float64, ne * zd * zd).reshape(ne, zd, zd)
            return BAGraph(vtype, cams, pts, e0, e1, z, info)
        poses = np.fromfile(f, np.float64, nv * zd).reshape(nv, zd)
        e0 = np.fromfile(f, np.uint64, ne).astype(np.int64)
        e1 = np.fromfile(f, np.uint64, ne).astype(np.int64)
        z = np.fromfile(f, np.float64, ne * zd).reshape(ne, zd)
        info = np.fromfile(f, np.float64, ne * zd * zd).reshape(ne, zd, zd)
        return PoseGraph(kind, poses, e0, e1, z, info)


def read_dump(path: str) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        while True:
            hdr = f.read(48)
            if len(hdr) < 48:
                break
            name = hdr[:32].split(b"\0", 1)[0].decode()
            dtype, count = struct.unpack("<2Q", hdr[32:])
            out[name] = np.fromfile(f, np.float64 if dtype == 0 else np.uint64, count)
    return out


def write_dump(path: str, arrays: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            if a.dtype == np.float64:
                dt = 0
            else:
                a = a.astype(np.uint64)
                dt = 1
            nm = name.encode()[:31]
            f.write(nm + b"\0" * (32 - len(nm)))
            f.write(struct.pack("<2Q", dt, a.size))
            a.tofile(f)
