#!/usr/bin/env python
"""Development helper: run the device Gauss-Newton solver on the Manhattan-3500-shape SE(2) graph, print the report
and, when oracle/_ref is present, the reference's time for the same Optimize(5, 0) on this machine."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import capi, graphs, sppio  # noqa: E402

g = graphs.make_manhattan(fill_loops=True)
ctx = capi.Context(0)
t = time.time()
ctx.pose_set_graph(g)
print(f"manhattan: N={g.poses.shape[0]} E={g.e_from.shape[0]}; set_graph {time.time() - t:.4f}s", flush=True)
for r in range(4):
    ctx.pose_restore_initial()
    l0 = ctx.kernel_launches
    t = time.time()
    rep = ctx.pose_optimize(5, 0.0)
    wall = time.time() - t
    print(json.dumps(dict(run=r, wall_s=round(wall, 5), launches=ctx.kernel_launches - l0, n=rep["n_iterations"],
                          chi2=(rep["chi2_initial"], rep["chi2_final"]), ms=rep["ms"])), flush=True)
ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver_pose")
if os.path.exists(ref):
    with tempfile.TemporaryDirectory() as td:
        sppio.write_graph(td + "/g.bin", g)
        for thr in ("1", str(os.cpu_count())):
            out = subprocess.run([ref, "time", td + "/g.bin", td + "/d.dump", "5", "0"], capture_output=True, text=True,
                                 env=dict(os.environ, OMP_NUM_THREADS=thr)).stdout
            print("reference:", out.strip())

# SE(3): sphere2500 shape (radius / noise chosen so that the reference's Gauss-Newton converges, see tests/golden/make_golden.py)
g3 = graphs.make_sphere(n_rings=50, n_per_ring=50, seed=2500, sigma_t=0.004, sigma_r=0.0004, radius=5.0)
t = time.time()
ctx.pose_set_graph(g3)
print(f"sphere2500: N={g3.poses.shape[0]} E={g3.e_from.shape[0]}; set_graph {time.time() - t:.4f}s", flush=True)
for r in range(4):
    ctx.pose_restore_initial()
    l0 = ctx.kernel_launches
    t = time.time()
    rep = ctx.pose_optimize(5, 0.0)
    wall = time.time() - t
    print(json.dumps(dict(run=r, wall_s=round(wall, 5), launches=ctx.kernel_launches - l0, n=rep["n_iterations"],
                          chi2=(rep["chi2_initial"], rep["chi2_final"]), ms=rep["ms"])), flush=True)
if os.path.exists(ref):
    with tempfile.TemporaryDirectory() as td:
        sppio.write_graph(td + "/g.bin", g3)
        for thr in ("1", str(os.cpu_count())):
            out = subprocess.run([ref, "time", td + "/g.bin", td + "/d.dump", "5", "0"], capture_output=True, text=True,
                                 env=dict(os.environ, OMP_NUM_THREADS=thr)).stdout
            print("reference:", out.strip())
