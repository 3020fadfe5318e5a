#!/usr/bin/env python
"""Development helper: run the device LM solver on one of the synthetic BA shapes and print the report."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_plus_plus_b200 import capi, graphs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="venice871")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--repeat", type=int, default=3)
ap.add_argument("--analytic", action="store_true")
a = ap.parse_args()

t = time.time()
g = graphs.ba_shape(a.shape)
print(f"graph {a.shape}: C={g.n_cams} P={g.n_pts} O={g.n_obs} generated in {time.time() - t:.2f}s", flush=True)
ctx = capi.Context(0)
print(ctx.describe())
t = time.time()
ctx.ba_set_graph(g)
print(f"set_graph (upload + symbolic): {time.time() - t:.3f}s", flush=True)
if a.analytic:
    ctx.ba_set_jacobian_mode(capi.JAC_ANALYTIC)
for r in range(a.repeat):
    ctx.ba_set_states(g.cams[:, :6], g.pts)
    l0 = ctx.kernel_launches
    t = time.time()
    rep = ctx.ba_optimize(a.iters, 0.0)
    wall = time.time() - t
    print(json.dumps(dict(run=r, wall_s=round(wall, 4), launches=ctx.kernel_launches - l0,
                          **{k: v for k, v in rep.items() if not k.startswith("trace_alpha")})), flush=True)
