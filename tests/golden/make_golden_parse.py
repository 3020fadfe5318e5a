#!/usr/bin/env python
"""Generates tests/golden/parse_*.txt / parse_ref.npz: small graph files in the reference's text formats and what the
UNMODIFIED reference's parser (oracle/_ref/ref_driver_parse: CParserTemplate + the parse primitives of
include/slam_app/ParsePrimitives.h) hands to its parse loop for them. Needs /root/reference (build container only).

usage: python tests/golden/make_golden_parse.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import graphfile, graphs, sppio  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver_parse")


def main():
    rng = np.random.default_rng(5)
    # BA: a tiny graph with interleaved ids and distortion, full 2x2 information
    g = graphs.ba_shape("tiny", interleave_ids=True, shuffle_edges=True, distortion=-0.03)
    g.info[:, 0, 1] = g.info[:, 1, 0] = 0.25
    graphfile.write_ba(os.path.join(HERE, "parse_ba.txt"), g)
    # SE(2): ascending edges, then a hand-written file with descending ("Manhattan order") edges, both matrix orders,
    # alternative tokens, a comment and an unknown token
    m = graphs.make_manhattan(40, 15, seed=5)
    graphfile.write_se2(os.path.join(HERE, "parse_se2.txt"), m)
    with open(os.path.join(HERE, "parse_se2_mixed.txt"), "w") as f:
        f.write("# descending edges are inverted by the parser\n")
        f.write("VERTEX_SE2 0 0 0 0\nVERTEX 1 1.0 0.1 0.05\n")
        f.write("EDGE 1 0 1.02 0.11 0.049 44.7 0 1000.0 44.7 0 0\n")          # french order |0 1 5; . 2 4; . . 3|
        f.write("ODOMETRY 2 1 0.98 -0.07 -0.3 50 0 0 60 0 700\n")              # descending, usual order
        f.write("EDGE_SE2 2 3 1.0 0.0 1.57 10 1 2 20 3 30\n")
        f.write("EDGE2 5 3 -0.5 0.25 -3.0 11 0.5 333 22 0.25 0.125\n")
        f.write("EQUIV 1 2\nEDGE2 3 4 0.3 0.2 0.1 1 0 0 1 0 1\n")
    # SE(3): a small sphere (VERTEX3 with roll-pitch-yaw, EDGE3:AXISANGLE), then a hand-written file with EDGE3 in
    # roll-pitch-yaw, rotations near pi about each axis (the three non-trace branches of the matrix -> quaternion
    # conversion), a full information matrix, a switched edge (reported and dropped) and the alternative tokens
    s3 = graphs.make_sphere(5, 8, seed=5, radius=5.0)
    graphfile.write_se3(os.path.join(HERE, "parse_se3.txt"), s3)
    diag = "100 0 0 0 0 0 100 0 0 0 0 100 0 0 0 400 0 0 400 0 400"
    full = " ".join("%.17g" % v for v in (np.arange(21) * 0.01 + np.array([5 if k in (0, 6, 11, 15, 18, 20) else 0 for k in range(21)])))
    with open(os.path.join(HERE, "parse_se3_mixed.txt"), "w") as f:
        f.write("VERTEX_SE3 0 0 0 0 0 0 0\nVERTEX3 1 1.0 0.1 -0.2 0.3 -0.4 0.5\n")
        f.write("VERTEX3 2 2.0 0.0 0.5 3.1 0.02 -0.01\nVERTEX3 3 2.5 1.0 0.5 0.03 3.12 0.01\n")
        f.write("VERTEX3 4 3.0 1.5 0.7 -0.02 0.01 -3.13\nVERTEX3 5 3.5 2.5 0.9 2.5 -1.2 -2.8\n")
        f.write("EDGE3 0 1 1.01 0.09 -0.21 0.31 -0.39 0.49 " + diag + "\n")
        f.write("EDGE_SE3 1 2 1.0 -0.1 0.7 2.8 0.4 -0.5 " + full + "\n")
        f.write("EDGE3 3 2 0.5 1.0 0.0 0.1 0.2 0.3 " + diag + "\n")           # switched order: dropped by the parser
        f.write("EDGE3 2 3 0.5 1.0 0.0 -3.0 3.0 0.1 " + diag + "\n")
        f.write("EDGE3:AXISANGLE 3 4 0.5 0.5 0.2 0.1 -3.0 0.2 " + full + "\n")
        f.write("EDGE_SE3:AXISANGLE 5 4 -0.5 -1.0 -0.2 1.5 0.5 -2.5 " + diag + "\n")  # kept as it is
        f.write("EDGE3 4 5 0.5 1.0 0.2 1e-9 -2e-9 1e-10 " + diag + "\n")     # tiny rotation
    out = {}
    for name in ("parse_ba", "parse_se2", "parse_se2_mixed", "parse_se3", "parse_se3_mixed"):
        with tempfile.TemporaryDirectory() as td:
            dp = os.path.join(td, "d.dump")
            subprocess.run([REF, os.path.join(HERE, name + ".txt"), dp], check=True)
            d = sppio.read_dump(dp)
        for k, v in d.items():
            out[name + "." + k] = v
    np.savez_compressed(os.path.join(HERE, "parse_ref.npz"), **out)


if __name__ == "__main__":
    main()
