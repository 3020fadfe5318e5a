#!/usr/bin/env python
"""Development helper: wall-clock split of the end-to-end call sequence bench.py times (Venice-871 shape, page-locked
host buffers): spp_ba_set_graph / spp_ba_optimize(5) / spp_ba_get_states."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from slam_plus_plus_b200 import capi, graphs  # noqa: E402
from slam_plus_plus_b200.sppio import BAGraph  # noqa: E402

g = graphs.ba_shape(sys.argv[1] if len(sys.argv) > 1 else "venice871")
ctx = capi.Context(0)
vtype = np.ascontiguousarray(g.vtype, np.uint8)
host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in
        (g.cams, g.pts, g.obs_pt.astype(np.uint64).view(np.int64), g.obs_cam.astype(np.uint64).view(np.int64), g.z, g.info)]
gh = BAGraph(vtype, host[0].numpy(), host[1].numpy(), host[2].numpy().view(np.uint64), host[3].numpy().view(np.uint64),
             host[4].numpy(), host[5].numpy())
out_c = torch.empty((g.n_cams, 6), dtype=torch.float64).pin_memory()
out_p = torch.zeros((g.n_pts, 3), dtype=torch.float64).pin_memory()
for step in range(6):
    t0 = time.perf_counter()
    ctx.ba_set_graph(gh)
    t1 = time.perf_counter()
    r = ctx.ba_optimize(5, 0.0)
    t2 = time.perf_counter()
    if step % 2:
        ctx.ba_get_states(out_c.numpy(), out_p.numpy())
    else:
        ctx.ba_get_states()
    t3 = time.perf_counter()
    print("step %d: set_graph %.2f ms, optimize %.2f ms (%d iterations), get_states (%s) %.2f ms" %
          (step, 1e3 * (t1 - t0), 1e3 * (t2 - t1), r["n_iterations"], "page-locked" if step % 2 else "pageable", 1e3 * (t3 - t2)), flush=True)
ctx.close()
