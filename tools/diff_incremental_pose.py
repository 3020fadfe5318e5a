#!/usr/bin/env python
"""Development helper: where do the incremental pose-graph runs of the slot-3 pose adapter and of the reference's own
CNonlinearSolver_Lambda part ways? Runs oracle/_ref/ref_driver_dropin_gn (b200 | ref) incremental with the solver in
verbose mode and compares the sequences of printed step norms ("residual norm: %.4f", one per linear solve).
    python tools/diff_incremental_pose.py [sphere2500|manhattan3500] [min_dx=0.01] [period=10]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import graphs, sppio  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "ref_driver_dropin_gn")
name = sys.argv[1] if len(sys.argv) > 1 else "sphere2500"
min_dx = sys.argv[2] if len(sys.argv) > 2 else "0.01"
period = sys.argv[3] if len(sys.argv) > 3 else "10"
g = graphs.make_manhattan(fill_loops=True) if name.startswith("manhattan") else graphs.make_sphere(
    n_rings=50, n_per_ring=50, seed=2500, sigma_t=0.004, sigma_r=0.0004, radius=5.0)
norms = {}
with tempfile.TemporaryDirectory() as td:
    sppio.write_graph(f"{td}/g.bin", g)
    for impl in ("b200", "ref"):
        out = subprocess.run([BIN, impl, "incremental", f"{td}/g.bin", f"{td}/d.dump", "5", min_dx, period],
                             capture_output=True, text=True, env=dict(os.environ, SPP_REF_VERBOSE="1"), cwd=td)
        seq, solve = [], -1
        for l in out.stdout.splitlines():
            if l.startswith("residual norm:"):
                seq.append(l.split(":")[1].strip())
        norms[impl] = seq
        print(impl, len(seq), "linear solves;", [l for l in out.stdout.splitlines() if "final chi2" in l])
a, b = norms["b200"], norms["ref"]
k = next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), None)
if k is None:
    print("the common prefix of", min(len(a), len(b)), "step norms is identical")
else:
    print("first difference at linear solve", k)
    print("  b200:", a[max(0, k - 6):k + 8])
    print("  ref :", b[max(0, k - 6):k + 8])
