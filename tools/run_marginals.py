#!/usr/bin/env python
"""Development helper: time the marginal-covariance recovery (spp_ba_marginals) on the Venice-871-shape graph after
Optimize(5), and check the camera blocks against solves with the same system (S^-1 e_i through spp_ba_solve_step is not
available with a custom right-hand side, so the check is symmetry / positivity / reproducibility)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import capi, graphs  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "venice871"
g = graphs.ba_shape(shape)
ctx = capi.Context(0)
ctx.ba_set_graph(g)
rep = ctx.ba_optimize(5, 0.0)
print(f"{shape}: C={g.n_cams} P={g.n_pts} O={g.n_obs} chi2 {rep['chi2_initial']:.6g} -> {rep['chi2_final']:.6g}", flush=True)
for alpha in (0.0, 1.0):
    for r in range(3):
        l0 = ctx.kernel_launches
        t = time.time()
        cc, pc = ctx.ba_marginals(alpha)
        wall = time.time() - t
        print(json.dumps(dict(alpha=alpha, run=r, wall_ms=round(wall * 1e3, 2), launches=ctx.kernel_launches - l0,
                              cam_var_max=float(np.einsum("kii->ki", cc).max()), pt_var_max=float(np.einsum("kii->ki", pc).max()),
                              min_diag=float(min(np.einsum("kii->ki", cc).min(), np.einsum("kii->ki", pc).min())),
                              asym=float(np.abs(cc - cc.transpose(0, 2, 1)).max() / np.abs(cc).max()))), flush=True)
    cc2, pc2 = ctx.ba_marginals(alpha)
    print("bitwise reproducible:", bool(np.array_equal(cc, cc2) and np.array_equal(pc, pc2)), flush=True)

# pose graphs: block diagonal of lambda^-1 through the dense inverse (spp_pose_marginals)
for name, gp in (("manhattan3500", graphs.make_manhattan(fill_loops=True)),
                 ("sphere2500", graphs.make_sphere(n_rings=50, n_per_ring=50, seed=2500, sigma_t=0.004, sigma_r=0.0004, radius=5.0))):
    ctx.pose_set_graph(gp)
    ctx.pose_optimize(5, 0.0)
    ctx.pose_linearise()
    for r in range(3):
        l0 = ctx.kernel_launches
        t = time.time()
        cov = ctx.pose_marginals()
        wall = time.time() - t
        print(json.dumps(dict(graph=name, run=r, wall_ms=round(wall * 1e3, 2), launches=ctx.kernel_launches - l0,
                              var_max=float(np.einsum("kii->ki", cov).max()), var_min=float(np.einsum("kii->ki", cov).min()))), flush=True)
