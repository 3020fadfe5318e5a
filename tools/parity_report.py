#!/usr/bin/env python
"""Prints the MEASURED parity errors of the CUDA path (what DESIGN.md quotes): slot-1 increments against the golden
vectors of the unmodified reference, dense solves against LAPACK at several condition numbers, and the residual of
the full-size dense solve. Development helper; the asserting versions live in tests/test_ba_gpu.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, lambda_to_dense, rel_err  # noqa: E402
from slam_plus_plus_b200 import capi  # noqa: E402

ctx = capi.Context(0)
for name in ["ba_tiny", "ba_tiny_interleaved", "ba_small", "ba_small_hard"]:
    g, d = load_golden(name)
    ctx.schur_symbolic(d["L0.col_dims"], d["L0.col_ptr"], d["L0.row_idx"])
    dx = ctx.schur_solve(d["L0.vals"], d["L0.eta"])
    A = lambda_to_dense(d["L0.col_dims"], d["L0.col_ptr"], d["L0.row_idx"], d["L0.vals"])
    x = np.linalg.solve(A, d["L0.eta"])
    print(f"{name:22s} slot-1 dx vs reference {rel_err(dx, d['L0.dx']):.2e}  vs LAPACK {rel_err(dx, x):.2e}  "
          f"(reference vs LAPACK {rel_err(d['L0.dx'], x):.2e}, cond {np.linalg.cond(A):.1e})")
rng = np.random.default_rng(1)
for n in (200, 777, 2600):
    for cond in (1e2, 1e6, 1e10):
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        A = (Q * np.geomspace(1, cond, n)) @ Q.T
        A = 0.5 * (A + A.T)
        b = rng.normal(size=n)
        x = ctx.dense_posdef_solve(A, b)
        xr = np.linalg.solve(A, b)
        print(f"dense n={n:5d} cond {cond:.0e}: vs LAPACK {rel_err(x, xr):.2e}  residual {np.linalg.norm(A @ x - b) / np.linalg.norm(b):.2e} "
              f"(LAPACK residual {np.linalg.norm(A @ xr - b) / np.linalg.norm(b):.2e})")
ctx.close()
