#!/usr/bin/env python
"""Development helper: a pose graph that grows between two solves on the SAME context (what the slot-3 adapter does in an
incremental run) against a fresh context per solve."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from slam_plus_plus_b200 import capi, graphs  # noqa: E402
from slam_plus_plus_b200.sppio import PoseGraph  # noqa: E402


def prefix(g, n):
    m = (np.maximum(g.e_from, g.e_to) < n)
    return PoseGraph(g.kind, g.poses[:n].copy(), g.e_from[m], g.e_to[m], g.z[m], g.info[m])


g = graphs.make_manhattan(fill_loops=True)
sizes = [int(a) for a in sys.argv[1:]] or [50, 150, 220, 280]
ctx = capi.Context(0)
for n in sizes:
    sub = prefix(g, n)
    res = {}
    for name, c in (("reused", ctx), ("fresh", capi.Context(0))):
        c.pose_set_graph(sub)
        try:
            rep = c.pose_optimize(5, 0.01)
            res[name] = (rep["status"], rep["n_iterations"], rep["chi2_initial"], rep["chi2_final"], [round(x, 4) for x in rep["trace_dx_norm"]])
        except Exception as e:  # noqa: BLE001
            res[name] = repr(e)
    print(n, "vertices,", len(sub.e_from), "edges:")
    for k, v in res.items():
        print("   ", k, v)
