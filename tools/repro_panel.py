#!/usr/bin/env python
"""Development helper: the dense front primitive on panels with many columns right of the triangle (a supernode's row
structure), dataflow kernel against the stream path and against numpy, repeated."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from slam_plus_plus_b200 import capi  # noqa: E402

ctx = capi.Context(0)
rng = np.random.default_rng(3)
shapes = [(1152, 1152 + 2048), (1024, 1024 + 4096), (2048, 2048 + 1024), (1280, 1280 + 128)]
if len(sys.argv) > 1:
    shapes = [(int(sys.argv[1]), int(sys.argv[2]))]
for n, m in shapes:
    g = rng.standard_normal((n, 64))
    a11 = g @ g.T + np.diag(1.0 + 10 * rng.random(n))
    a12 = rng.standard_normal((n, m - n))
    panel = np.asfortranarray(np.hstack([np.triu(a11), a12]))
    os.environ["SPP_CHOL_DATAFLOW"] = "0"
    ref = ctx.dense_panel_factor(panel)
    r11 = np.triu(ref[:, :n])
    print(n, m, "stream path: |R^T R - A| %.2e, |R^T X - A12| %.2e" % (np.abs(r11.T @ r11 - a11).max() / np.abs(a11).max(),
                                                                      np.abs(r11.T @ ref[:, n:] - a12).max() / np.abs(a12).max()), flush=True)
    os.environ["SPP_CHOL_DATAFLOW"] = "1"
    first, bad = None, 0
    for k in range(int(os.environ.get("REPS", "30"))):
        out = ctx.dense_panel_factor(panel)
        out[:, :n] = np.triu(out[:, :n])
        if first is None:
            first = out
            err = max(np.abs(np.triu(out[:, :n]) - r11).max(), np.abs(out[:, n:] - ref[:, n:]).max())
            print("   dataflow vs stream path: max abs diff %.2e" % err, flush=True)
        elif np.isnan(out).any():
            bad += 1
            nn = np.isnan(out)
            cols = np.nonzero(nn.any(axis=0))[0]
            rows = np.nonzero(nn.any(axis=1))[0]
            print("   run %d has NaN: columns %d.. rows %d.." % (k, cols[0], rows[0]), flush=True)
        elif not np.array_equal(out, first):
            bad += 1
            d = np.abs(out - first)
            cols = np.nonzero(d.max(axis=0))[0]
            rows = np.nonzero(d.max(axis=1))[0]
            print("   run %d differs from run 0: max %.2e, columns %d..%d (%d), rows %d..%d (%d)" % (k, d.max(), cols[0], cols[-1], len(cols), rows[0], rows[-1], len(rows)), flush=True)
            # the first wrong tile in dependency order: smallest (tile row + tile column), then the pattern inside it
            tiles = [(i, j) for i in range(n // 128) for j in range(i, m // 128) if d[i * 128:(i + 1) * 128, j * 128:(j + 1) * 128].max() > 0]
            ti, tj = min(tiles, key=lambda t: (t[0], t[1]))
            blk = d[ti * 128:(ti + 1) * 128, tj * 128:(tj + 1) * 128] > 0
            lr, lc = np.nonzero(blk.any(axis=1))[0], np.nonzero(blk.any(axis=0))[0]
            print("      first wrong tile (%d, %d): local rows %s, local cols %s, %d entries; next wrong tiles %s" % (ti, tj, (lr[0], lr[-1], len(lr)), (lc[0], lc[-1], len(lc)), blk.sum(), sorted(tiles)[:6]), flush=True)
            if tj == ti + 1:
                # forensic: the partial sums T' the wrong tile implies (T' = R(ti,ti)^T W) against the true ones: is a slab missing?
                sl = lambda a, b: (slice(a * 128, (a + 1) * 128), slice(b * 128, (b + 1) * 128))
                Rii = np.triu(first[sl(ti, ti)])
                dT = Rii.T @ (out[sl(ti, tj)] - first[sl(ti, tj)])
                print("         implied error of the partial sums: max %.3e in local cols %s" % (np.abs(dT).max(), np.nonzero(np.abs(dT).max(axis=0) > 1e-9)[0][[0, -1]]))
                c = np.nonzero(np.abs(dT).max(axis=0) > 1e-9)[0]
                best = []
                for k in range(ti):
                    Ak, Bk = first[sl(k, ti)], first[sl(k, tj)]
                    for ch in range(8):
                        rows = slice(16 * ch, 16 * ch + 16)
                        term = Ak[rows].T @ Bk[rows][:, c]           # this chunk's contribution to the columns in question
                        best.append((np.abs(dT[:, c] - term).max(), "chunk (%d, %d) missing" % (k, ch)))
                        q = 8 * k + ch - 6                            # the chunk that occupied the pipeline stage before
                        if q >= 0:
                            kp, cp = divmod(q, 8)
                            Bp = first[sl(kp, tj)][16 * cp:16 * cp + 16][:, c]
                            best.append((np.abs(dT[:, c] - Ak[rows].T @ (Bk[rows][:, c] - Bp)).max(), "chunk (%d, %d) computed with the B rows of chunk (%d, %d)" % (k, ch, kp, cp)))
                            Ap = first[sl(kp, ti)][16 * cp:16 * cp + 16]
                            best.append((np.abs(dT[:, c] - (Ak[rows] - Ap).T @ Bk[rows][:, c]).max(), "chunk (%d, %d) computed with the A rows of chunk (%d, %d)" % (k, ch, kp, cp)))
                sv = np.linalg.svd(dT[:, c], compute_uv=False)
                print("         singular values of dT: %s (numerical rank %d)" % (np.array2string(sv[:20], precision=2), int((sv > 1e-9 * sv[0]).sum())), flush=True)
                proj = []
                for k in range(ti):
                    Ak = first[sl(k, ti)]
                    for ch in range(8):
                        Q, _ = np.linalg.qr(Ak[16 * ch:16 * ch + 16].T)   # span of this chunk's A columns
                        proj.append((np.abs(dT[:, c] - Q @ (Q.T @ dT[:, c])).max(), "A chunk (%d, %d)" % (k, ch)))
                proj.sort()
                print("         residual after projecting dT on one A chunk's span:", proj[:3], flush=True)
                # a few wrong rows of dT: is each one a combination of the rows of ONE B chunk (a wrong A line in that chunk)?
                wr = np.nonzero(np.abs(dT[:, c]).max(axis=1) > 1e-9)[0]
                print("         wrong rows of dT:", wr[:16], flush=True)
                for r in wr[:3]:
                    cand = []
                    for k in range(ti):
                        Bk = first[sl(k, tj)]
                        for ch in range(8):
                            Bc = Bk[16 * ch:16 * ch + 16][:, c]              # 16 x len(c)
                            coef, *_ = np.linalg.lstsq(Bc.T, dT[r, c], rcond=None)
                            cand.append((np.abs(Bc.T @ coef - dT[r, c]).max(), k, ch, coef))
                    cand.sort(key=lambda t: t[0])
                    res, k, ch, coef = cand[0]
                    print("         row %d: best B chunk (%d, %d) residual %.3e (next best %.3e)" % (r, k, ch, res, cand[1][0]), flush=True)
                    a_true = first[sl(k, ti)][16 * ch:16 * ch + 16, r]
                    a_used = a_true - coef                                     # T = A - sum a b: dT = -(a_used - a_true) b
                    print("            a_true", np.array2string(a_true, precision=4))
                    print("            a_used", np.array2string(a_used, precision=4))
                    for dq in (-12, -6, 6, 12):
                        q = 8 * k + ch + dq
                        if 0 <= q < 8 * ti:
                            kp, cp = divmod(q, 8)
                            alt = first[sl(kp, ti)][16 * cp:16 * cp + 16, r]
                            print("            chunk %+d (%d, %d) same line: max diff to a_used %.3e" % (dq, kp, cp, np.abs(alt - a_used).max()))
                    print("            wrong k positions:", np.nonzero(np.abs(coef) > 1e-9 * max(1e-300, np.abs(a_true).max()))[0], flush=True)
                best.sort()
                print("         |dT| max %.3e over %d columns; best explanations:" % (np.abs(dT[:, c]).max(), len(c)), best[:3], flush=True)
    print("   %d of the repeated runs differ" % bad)
