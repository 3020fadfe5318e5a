#!/usr/bin/env python
"""Generates tests/golden/order_ref.npz: block patterns (reduced camera systems of seeded BA graphs, a pose graph) and
the fill-reducing ordering the UNMODIFIED reference computes for them -- CMatrixOrdering::p_BlockOrdering
(src/slam/OrderingMagic.cpp:701-1033, SuiteSparse amd_l2 on the block graph) through oracle/_ref/ref_driver_order.
Needs /root/reference (build container only); the .npz is committed.

usage: python tests/golden/make_golden_order.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import graphs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ORDER = os.path.join(ROOT, "oracle", "_ref", "ref_driver_order")
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ref_order(col_ptr, row_idx):
    with tempfile.TemporaryDirectory() as td:
        pin, pout = os.path.join(td, "p.bin"), os.path.join(td, "o.bin")
        with open(pin, "wb") as f:
            np.array([len(col_ptr) - 1, len(row_idx)], np.uint64).tofile(f)
            col_ptr.astype(np.uint64).tofile(f)
            row_idx.astype(np.uint64).tofile(f)
        subprocess.run([REF_ORDER, pin, pout], check=True, stdout=subprocess.DEVNULL)
        return np.fromfile(pout, np.uint64)


def random_upper(n, density, rng, hubs=0):
    """a random symmetric pattern as upper block CSC with the diagonal; hubs = rows that touch most of the others (AMD
    sets rows with more than max(16, 10 sqrt(n)) entries aside)"""
    a = rng.random((n, n)) < density
    for _ in range(hubs):
        a[rng.integers(n), :] = rng.random(n) < 0.6
    a = np.triu(a | a.T, 1) | np.eye(n, dtype=bool)
    col_ptr, row_idx = [0], []
    for c in range(n):
        row_idx.extend(np.nonzero(a[:c + 1, c])[0].tolist())
        col_ptr.append(len(row_idx))
    return np.array(col_ptr, np.uint64), np.array(row_idx, np.uint64)


def main():
    from conftest import load_golden
    rng = np.random.default_rng(1)
    cases = {
        "rcs_small": graphs.rcs_block_pattern(load_golden("ba_small")[0]),       # the graph of ba_small.npz
        "rcs_mid": graphs.rcs_block_pattern(graphs.ba_shape("mid")),
        "rcs_seq400": graphs.rcs_block_pattern(graphs.make_ba(400, 30000, 400, mean_extra_track=3.0, max_track=20, max_stride=4, loops=2)),
        "pose_manhattan800": graphs.pose_block_pattern(graphs.make_manhattan(800, 450, seed=800)),
        # round 2: the library's ordering must BE the reference's; corner cases of the tie-breaking
        "rand_n1": random_upper(1, 0.5, rng), "rand_n2_full": random_upper(2, 1.0, rng), "rand_n5_empty": random_upper(5, 0.0, rng),
        "rand_n30": random_upper(30, 0.1, rng), "rand_n64_hubs": random_upper(64, 0.3, rng, 2),
        "rand_n200": random_upper(200, 0.02, rng), "rand_n200_hubs": random_upper(200, 0.05, rng, 3),
        "rand_n500_hubs": random_upper(500, 0.01, rng, 5), "rand_n1000_hubs": random_upper(1000, 0.003, rng, 8),
        "rand_n300_dense": random_upper(300, 0.5, rng), "rand_n2000_hubs": random_upper(2000, 0.0015, rng, 20),
        "pose_manhattan3500": graphs.pose_block_pattern(graphs.make_manhattan(3500, 1953, seed=3500, fill_loops=True)),
        "rcs_venice871": graphs.rcs_block_pattern(graphs.ba_shape("venice871")),
    }
    # BAL-13682: the pattern is large (2.3 M blocks) but regenerates from the seeded graph; only the order is stored
    bal_cp, bal_ri = graphs.rcs_block_pattern(graphs.ba_shape("bal13682"))
    out = {}
    for name, (cp, ri) in cases.items():
        o = ref_order(cp, ri)
        assert sorted(o.tolist()) == list(range(len(cp) - 1))
        out[name + ".col_ptr"], out[name + ".row_idx"], out[name + ".order"] = cp, ri, o
        print(f"{name}: n={len(cp) - 1} nnzb={len(ri)}")
    out["rcs_bal13682.order"] = ref_order(bal_cp, bal_ri).astype(np.uint16)
    np.savez_compressed(os.path.join(HERE, "order_ref.npz"), **out)


if __name__ == "__main__":
    main()
