#!/usr/bin/env python
"""Development helper: marker-driven incremental BA (BASELINE config 5 shape: cameras streamed in batches of 10, Optimize()
at every marker) through the slot-3 adapter -- the unmodified reference application logic with
CNonlinearSolver_Lambda_LM_B200 (oracle/_ref/ref_driver_dropin_lm b200 incremental) -- and, on a smaller sample, the
reference's own solver for comparison."""
import argparse
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import graphs, sppio  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="venice871")
ap.add_argument("--ref-cams", type=int, default=100, help="cameras of the sub-sequence the reference solver is timed on")
a = ap.parse_args()
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_driver_dropin_lm")
g = graphs.ba_shape(a.shape)
with tempfile.TemporaryDirectory() as td:
    sppio.write_graph(td + "/g.bin", g)
    t = time.time()
    out = subprocess.run([BIN, "b200", "incremental", td + "/g.bin", td + "/d.dump", "5", "0", "10"], capture_output=True, text=True, env=dict(os.environ, SPP_REF_DUMP_TIMING="1"))
    print(f"{a.shape} (C={g.n_cams} P={g.n_pts} O={g.n_obs}), b200, whole process {time.time() - t:.1f}s:", out.stdout.strip(), flush=True)
    d = sppio.read_dump(td + "/d.dump")
    print("  markers", len(d["chi2_trace"]), "chi2 first/last", d["chi2_trace"][0], d["chi2_trace"][-1])
    gs = graphs.make_ba(a.ref_cams, int(g.n_pts * a.ref_cams / g.n_cams), 871, mean_extra_track=3.35, max_track=60, max_stride=11)
    sppio.write_graph(td + "/s.bin", gs)
    for impl in ("b200", "ref"):
        out = subprocess.run([BIN, impl, "incremental", td + "/s.bin", td + "/s.dump", "5", "0", "10"], capture_output=True, text=True)
        print(f"sample C={gs.n_cams} P={gs.n_pts} O={gs.n_obs}, {impl}:", out.stdout.strip(), flush=True)
