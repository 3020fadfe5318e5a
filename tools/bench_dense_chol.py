#!/usr/bin/env python
"""Dense FP64 Cholesky solve through the C ABI (spp_dense_posdef_solve): the persistent dataflow kernel
(SPP_CHOL_DATAFLOW=1, default) against the stream-scheduled right-looking factorisation (=0) on random SPD matrices.
Prints residuals, the difference between the two solutions and (SPP_CHOL_TIMING=1) the device times.
    python tools/bench_dense_chol.py [n ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SPP_CHOL_TIMING", "1")
from slam_plus_plus_b200 import capi  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [128, 130, 1000, 5226]
    ctx = capi.Context(0)
    rng = np.random.default_rng(5)
    ok = True
    for n in sizes:
        m = rng.standard_normal((n, 96))
        a = m @ m.T + np.diag(1.0 + rng.random(n) * 10)
        b = rng.standard_normal(n)
        xs = {}
        for mode in ("1", "0", "1"):
            os.environ["SPP_CHOL_DATAFLOW"] = mode
            t0 = time.time()
            x = ctx.dense_posdef_solve(a, b)
            dt = time.time() - t0
            res = np.linalg.norm(a @ x - b) / np.linalg.norm(b)
            print("n %5d dataflow=%s: relative residual %.3e (call %.1f ms)" % (n, mode, res, dt * 1e3), flush=True)
            ok &= bool(res < 1e-10)
            if mode in xs:
                same = np.array_equal(xs[mode], x)
                print("        repeated run bit-identical: %s" % same, flush=True)
                ok &= same
            xs[mode] = x
        d = np.linalg.norm(xs["1"] - xs["0"]) / np.linalg.norm(xs["0"])
        print("        |x_dataflow - x_streams| / |x| = %.3e" % d, flush=True)
        ok &= bool(d < 1e-9)
    # not positive definite: both paths must say so
    n = 700
    m = rng.standard_normal((n, 32))
    a = m @ m.T + np.eye(n)
    a[300, 300] = -5.0
    for mode in ("1", "0"):
        os.environ["SPP_CHOL_DATAFLOW"] = mode
        try:
            ctx.dense_posdef_solve(a, np.ones(n))
            print("not-PD matrix, dataflow=%s: NOT detected" % mode)
            ok = False
        except Exception as e:  # noqa: BLE001
            print("not-PD matrix, dataflow=%s: %s" % (mode, e))
    print("OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
