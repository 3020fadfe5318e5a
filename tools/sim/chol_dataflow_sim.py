#!/usr/bin/env python
"""Discrete-event model of the persistent dataflow Cholesky (csrc/dense_chol_mega.cu): left-looking tile tasks handed
out in row-major order to W worker CTAs, one potrf CTA, one inverse CTA and a helper group for the two tiles next to the
diagonal. Used to choose tile shapes / role counts before spending GPU time. Times in microseconds."""
import heapq
import sys


def simulate(NB=41, n_rhs_tiles=1, W=136, BN=64, t_slab128=21.0, t_potrf=14.0, t_inv=10.0, t_flag=0.8, t_hslab=3.0, t_htrsm=3.0,
             trsm_frac=0.5, verbose=False):
    """BN: columns per regular task (128 or 64). t_slab128: one 128x128x128 slab on one CTA."""
    sub = 128 // BN
    t_slab = t_slab128 / sub
    t_trsm = t_slab * trsm_frac + 1.0
    NJ = NB + n_rhs_tiles
    INF = float("inf")
    ready = {}      # (i, j, s) -> time R(i, j sub-tile s) final
    F1 = [INF] * NB
    F2 = [INF] * NB
    diag_ready = [INF] * (NB + 1)
    diag_ready[0] = 0.0
    # chain tiles: (i, i) and (i, i+1): workers accumulate slabs <= i-2 (diag: <= i-1 needs R(i-1,i) -> helper), helpers finish.
    tasks = []
    for i in range(NB):
        for j in range(i, NJ):
            for s in range(sub):
                tasks.append((i, j, s))
    workers = [0.0] * W
    heapq.heapify(workers)
    part = {}       # chain tiles: time the worker's partial sum is in memory
    busy = 0.0

    def rdy(k, col, s=None):
        # R(k, col) all sub-tiles
        if s is None:
            return max(ready.get((k, col, q), INF) for q in range(sub))
        return ready.get((k, col, s), INF)

    # process panels in order; tasks in row-major order are popped as workers free up. Because readiness of row i-1 is
    # needed to time row i we iterate rows, and inside a row first resolve the chain.
    ti = 0
    for i in range(NB):
        # --- chain for panel i
        F1[i] = diag_ready[i] + 1.5 + t_potrf + 1.0   # load, factor, store
        F2[i] = F1[i] + t_flag + 1.5 + t_inv + 1.0
        # --- row i tasks
        row_tasks = [(i, j, s) for j in range(i, NJ) for s in range(sub)]
        for (ii, j, s) in row_tasks:
            t = heapq.heappop(workers)
            t0 = t
            chain_tile = (j == i) or (j == i + 1 and j < NB)
            kmax = i if not chain_tile else (i - 1 if j == i + 1 else i - 1)
            kmax = max(kmax, 0)
            for k in range(kmax):
                t = max(t, rdy(k, i) if j != i else rdy(k, i, s), rdy(k, j, s)) + t_slab
            if chain_tile:
                t += 1.0   # store the partial sum
                part[(i, j, s)] = t + t_flag
            else:
                t = max(t, F2[i] + t_flag) + t_trsm + 0.5
                ready[(i, j, s)] = t + t_flag
            busy += t - t0
            heapq.heappush(workers, t)
        # --- helpers: finish T(i, i+1): last slab (i-1), then trsm against R(i,i)
        if i + 1 < NB:
            p = max(part[(i, i + 1, s)] for s in range(sub))
            if i >= 1:
                p = max(p, rdy(i - 1, i), rdy(i - 1, i + 1)) + t_hslab
            t = max(p, F1[i] + t_flag) + t_htrsm
            for s in range(sub):
                ready[(i, i + 1, s)] = t + t_flag
            # diag (i+1, i+1): the worker's partial holds slabs <= i-1 ... it is a row i+1 task: handled below
        # diag tile of the next panel: its partial (slabs <= i-1) comes from a row-(i+1) worker task, which is timed in the
        # next loop iteration -- but the chain needs it now. Approximate: the partial of (i+1, i+1) is produced by the helpers'
        # previous idle time if its inputs R(k<=i-1, i+1) are there (they are: row i-1 is complete long before).
        if i + 1 < NB:
            inputs = max([rdy(k, i + 1) for k in range(i)] + [0.0])
            diag_ready[i + 1] = max(inputs + t_flag, ready[(i, i + 1, 0)]) + t_hslab + t_flag
        # the diagonal "ready" entries for rdy(k, i) lookups: R(i, i) itself is never a GEMM operand
    total = max(max(workers), F2[NB - 1])
    if verbose:
        print("potrf chain end %.0f us" % F1[NB - 1])
    return total, busy / (W * total)


if __name__ == "__main__":
    for BN in (128, 64):
        for W in (138,):
            for tp in (33.0, 20.0, 14.0):
                for slab in (18.0, 21.0, 24.0):
                    tot, util = simulate(BN=BN, W=W, t_potrf=tp, t_slab128=slab)
                    print("BN %3d W %d potrf %4.0f slab %4.0f -> total %6.0f us, worker utilisation %.2f, %.1f TF/s" % (BN, W, tp, slab, tot, util, 5226 ** 3 / 3 / tot / 1e6))
