#!/usr/bin/env python
"""Development helper: relative residual of the reduced-camera-system solve on the BAL-13682 shape (block-sparse path),
first solves of fresh contexts (where a race between the streams of the supernodal factorisation would show)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from slam_plus_plus_b200 import capi, graphs  # noqa: E402

g = graphs.ba_shape("bal13682")
worst = 0.0
for trial in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    ctx = capi.Context(0)
    ctx.ba_set_graph(g)
    ctx.ba_linearise()
    res = []
    for alpha in (50.0, 1.0, 2.0, 3.0):
        ctx.ba_solve_step(alpha)
        res.append(ctx.schur_get_rcs_residual())
    worst = max(worst, max(res))
    print("trial", trial, "RCS relative residuals", res, flush=True)
    ctx.close()
print("worst", worst)
