#!/usr/bin/env python
"""Development helper: incremental pose-graph SLAM (BASELINE config 3 / config 1 shapes fed edge by edge, a nonlinear
solve every 10 new vertices once a loop has closed -- slam_app's "-nsp 10") through the slot-3 adapter for pose graphs:
the unmodified reference's system and feeding logic with CNonlinearSolver_Lambda_B200 (oracle/_ref/ref_driver_dropin_gn
b200 incremental), next to the reference's own CNonlinearSolver_Lambda on the same machine."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import graphs, sppio  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "ref_driver_dropin_gn")
GRAPHS = (("manhattan3500", graphs.make_manhattan(fill_loops=True)),
          ("sphere2500", graphs.make_sphere(n_rings=50, n_per_ring=50, seed=2500, sigma_t=0.004, sigma_r=0.0004, radius=5.0)))
with tempfile.TemporaryDirectory() as td:
    for name, g in GRAPHS:
        sppio.write_graph(f"{td}/{name}.bin", g)
        for impl in ("b200", "ref"):
            t = time.time()
            out = subprocess.run([BIN, impl, "incremental", f"{td}/{name}.bin", f"{td}/d.dump", "5", "0.01", "10"],
                                 capture_output=True, text=True, env=dict(os.environ, SPP_REF_DUMP_TIMING="1"), cwd=td)
            print(f"{name} {impl} (whole process {time.time() - t:.1f}s):", flush=True)
            print("   " + "\n   ".join(l for l in out.stdout.splitlines() if "took" in l or "host side" in l or "device" in l
                                       or "ref_driver" in l), flush=True)
            if out.returncode:
                print(out.stderr[-500:])
