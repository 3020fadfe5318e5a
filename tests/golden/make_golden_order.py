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


def main():
    from conftest import load_golden
    cases = {
        "rcs_small": graphs.rcs_block_pattern(load_golden("ba_small")[0]),       # the graph of ba_small.npz
        "rcs_mid": graphs.rcs_block_pattern(graphs.ba_shape("mid")),
        "rcs_seq400": graphs.rcs_block_pattern(graphs.make_ba(400, 30000, 400, mean_extra_track=3.0, max_track=20, max_stride=4, loops=2)),
        "pose_manhattan800": graphs.pose_block_pattern(graphs.make_manhattan(800, 450, seed=800)),
    }
    out = {}
    for name, (cp, ri) in cases.items():
        o = ref_order(cp, ri)
        assert sorted(o.tolist()) == list(range(len(cp) - 1))
        out[name + ".col_ptr"], out[name + ".row_idx"], out[name + ".order"] = cp, ri, o
        print(f"{name}: n={len(cp) - 1} nnzb={len(ri)}")
    np.savez_compressed(os.path.join(HERE, "order_ref.npz"), **out)


if __name__ == "__main__":
    main()
