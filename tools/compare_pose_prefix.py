#!/usr/bin/env python
"""Development helper: lambda / eta / dx of the first solve on a prefix of the sphere2500 graph (the system of the first
nonlinear solve of the incremental run), library against the reference's dump (oracle/_ref/ref_driver_pose dump)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import capi, graphs, sppio  # noqa: E402
from slam_plus_plus_b200.sppio import PoseGraph  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 59
g = graphs.make_sphere(n_rings=50, n_per_ring=50, seed=2500, sigma_t=0.004, sigma_r=0.0004, radius=5.0)
m = np.maximum(g.e_from, g.e_to) < n
# the order in which the incremental driver adds the edges: by the later pose, odometry first
idx = np.nonzero(m)[0]
key = np.maximum(g.e_from[idx], g.e_to[idx]) * 2 + (g.e_to[idx] != g.e_from[idx] + 1)
idx = idx[np.argsort(key, kind="stable")]
sub = PoseGraph(g.kind, g.poses[:n].copy(), g.e_from[idx], g.e_to[idx], g.z[idx], g.info[idx])
td = tempfile.mkdtemp()
sppio.write_graph(f"{td}/g.bin", sub)
subprocess.run([os.path.join(ROOT, "oracle/_ref/ref_driver_pose"), "dump", f"{td}/g.bin", f"{td}/d.dump", "3", "0"], check=True,
               stdout=subprocess.DEVNULL)
d = sppio.read_dump(f"{td}/d.dump")
ctx = capi.Context(0)
ctx.pose_set_graph(sub)
print("chi2", ctx.pose_chi2())
ctx.pose_linearise()
cp, ri, vals, eta = ctx.pose_get_lambda()
print("pattern equal:", np.array_equal(cp, d["L0.col_ptr"]) and np.array_equal(ri, d["L0.row_idx"]))
rv = d["L0.vals"]
print("lambda rel err %.3e, eta rel err %.3e" % (np.linalg.norm(vals - rv) / np.linalg.norm(rv), np.linalg.norm(eta - d["L0.eta"]) / np.linalg.norm(d["L0.eta"])))
B = 6
blk_err = np.linalg.norm((vals - rv).reshape(-1, B * B), axis=1) / (np.linalg.norm(rv.reshape(-1, B * B), axis=1) + 1e-300)
worst = np.argsort(-blk_err)[:8]
cols = np.searchsorted(cp, worst, side="right") - 1
print("worst blocks (row, col, rel err):", [(int(ri[b]), int(c), float("%.3g" % blk_err[b])) for b, c in zip(worst, cols)])
eta_err = np.abs(eta - d["L0.eta"]).reshape(-1, B).max(axis=1)
print("eta worst vertices:", np.argsort(-eta_err)[:8], np.sort(eta_err)[::-1][:4])
dx = ctx.pose_solve_step()
print("dx norm lib %.6f ref %.6f, rel err %.3e" % (np.linalg.norm(dx), np.linalg.norm(d["L0.dx"]), np.linalg.norm(dx - d["L0.dx"]) / np.linalg.norm(d["L0.dx"])))
