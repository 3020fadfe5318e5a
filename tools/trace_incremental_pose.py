#!/usr/bin/env python
"""Development helper: chi2 after every edge (SPP_TRACE_STEPS) and the step norms of the incremental pose-graph run, slot-3
pose adapter next to the reference's own solver, side by side around the first differing line."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import graphs, sppio  # noqa: E402

g = graphs.make_manhattan(fill_loops=True) if (len(sys.argv) < 3 or sys.argv[2].startswith("manhattan")) else graphs.make_sphere(
    n_rings=50, n_per_ring=50, seed=2500, sigma_t=0.004, sigma_r=0.0004, radius=5.0)
td = tempfile.mkdtemp()
sppio.write_graph(f"{td}/g.bin", g)
out = {}
for impl in ("b200", "ref"):
    r = subprocess.run([os.path.join(ROOT, "oracle/_ref/ref_driver_dropin_gn"), impl, "incremental", f"{td}/g.bin", f"{td}/d.dump", "5", "0.01", "10"],
                       capture_output=True, text=True, env=dict(os.environ, SPP_TRACE_STEPS=sys.argv[1] if len(sys.argv) > 1 else "400", SPP_REF_VERBOSE="1"))
    out[impl] = [l for l in r.stdout.splitlines() if l.startswith("step ") or l.startswith("residual")]
    err = [l for l in r.stderr.splitlines() if l.strip()]
    print(impl, "stderr:", len(err), "lines;", err[:5])
a, b = out["b200"], out["ref"]
def differs(x, y):
    if x.startswith("step ") and y.startswith("step "):
        if x.split("chi2")[0] != y.split("chi2")[0]:
            return True
        u, v = float(x.split("chi2")[1]), float(y.split("chi2")[1])
        return abs(u - v) > 1e-6 * max(abs(u), abs(v)) + 1e-12
    return x != y


k = next((i for i in range(min(len(a), len(b))) if differs(a[i], b[i])), None)
print("first differing line", k)
if k is not None:
    for i in range(max(0, k - 12), k + 14):
        print("%-70s | %s" % (a[i] if i < len(a) else "", b[i] if i < len(b) else ""))
