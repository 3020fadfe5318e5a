#!/usr/bin/env python
"""Development helper: the block-sparse (supernodal) reduced-camera-system solver against the dense one, then timing."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_plus_plus_b200 import capi, graphs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="tiny,small,mid,venice871")
ap.add_argument("--big", default="")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()

ctx = capi.Context(0)
print(ctx.describe(), flush=True)
for name in [s for s in a.shapes.split(",") if s]:
    g = graphs.ba_shape(name)
    ctx.schur_set_rcs_solver(capi.RCS_DENSE)
    ctx.ba_set_graph(g)
    ctx.ba_linearise()
    dx_d = ctx.ba_solve_step(1e-3 * 1000)
    ctx.schur_set_rcs_solver(capi.RCS_SPARSE)
    t = time.time()
    dx_s = ctx.ba_solve_step(1e-3 * 1000)
    t1 = time.time() - t
    t = time.time()
    dx_s = ctx.ba_solve_step(1e-3 * 1000)
    t2 = time.time() - t
    info = ctx.schur_get_rcs_info()
    err = np.linalg.norm(dx_s - dx_d) / np.linalg.norm(dx_d)
    print(f"{name}: C={g.n_cams} sparse vs dense rel err {err:.3e}; first solve {t1:.3f}s (symbolic incl.), second {t2:.4f}s; "
          + json.dumps({k: v for k, v in info.items() if k != 'order'}), flush=True)
for name in [s for s in a.big.split(",") if s]:
    t = time.time()
    g = graphs.ba_shape(name)
    print(f"{name}: C={g.n_cams} P={g.n_pts} O={g.n_obs} generated in {time.time() - t:.1f}s", flush=True)
    ctx.schur_set_rcs_solver(capi.RCS_AUTO)
    t = time.time()
    ctx.ba_set_graph(g)
    print(f"set_graph {time.time() - t:.2f}s", flush=True)
    t = time.time()
    rep = ctx.ba_optimize(1, 0.0)
    print(f"first Optimize(1) {time.time() - t:.2f}s (symbolic incl.)", json.dumps(rep["ms"]), flush=True)
    info = ctx.schur_get_rcs_info()
    print(json.dumps({k: v for k, v in info.items() if k != 'order'}), flush=True)
    for r in range(2):
        ctx.ba_restore_initial()
        t = time.time()
        rep = ctx.ba_optimize(a.iters, 0.0)
        wall = time.time() - t
        ms = rep["ms"]
        n = rep["n_iterations"]
        print(json.dumps(dict(run=r, wall_s=round(wall, 3), n=n, chi2=[rep["chi2_initial"], rep["chi2_final"]],
                              schur_solve_ms_per_iter=(ms["schur"] + ms["factor"] + ms["backsubst"]) / n,
                              factor_tflops=info["factor_flops"] / (ms["factor"] / n * 1e-3) / 1e12,
                              ms_per_iter={k: v / n for k, v in ms.items()})), flush=True)
