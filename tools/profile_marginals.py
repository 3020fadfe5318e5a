#!/usr/bin/env python
"""Development helper for `ncu --metrics gpu__time_duration.sum`: one linearisation and one spp_ba_marginals call on the
Venice-871-shape graph (the launch list of the marginals path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import capi, graphs  # noqa: E402

g = graphs.ba_shape("venice871")
ctx = capi.Context(0)
ctx.ba_set_graph(g)
ctx.ba_linearise()
cc, pc = ctx.ba_marginals(0.0)
print("marginals done", cc.shape, pc.shape, ctx.kernel_launches)
