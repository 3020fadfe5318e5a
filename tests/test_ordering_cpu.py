"""Host-side integer work of the block-sparse Cholesky (csrc/block_ordering.cpp) -- CPU tests, no GPU needed.

Pins: (1) the symbolic factorisation (column counts, elimination tree) against a brute-force boolean elimination;
(2) the library's own fill-reducing ordering against the ordering the UNMODIFIED reference computes for the same
patterns (tests/golden/order_ref.npz, made by tests/golden/make_golden_order.py with CMatrixOrdering::p_BlockOrdering,
src/slam/OrderingMagic.cpp:701-1033): it must be a permutation with fill within 15 % of the reference's AMD."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from slam_plus_plus_b200 import capi


def brute_force_symbolic(n, col_ptr, row_idx, order):
    inv = np.empty(n, np.int64)
    inv[np.asarray(order, np.int64)] = np.arange(n)
    A = np.zeros((n, n), bool)
    for c in range(n):
        for k in range(int(col_ptr[c]), int(col_ptr[c + 1])):
            r = int(row_idx[k])
            A[inv[r], inv[c]] = A[inv[c], inv[r]] = True
    L = np.tril(A)
    for j in range(n):
        rows = np.flatnonzero(L[j + 1:, j]) + j + 1
        for a in rows:
            L[rows[rows >= a], a] = True
    counts = L.sum(0)
    parent = np.array([(np.flatnonzero(L[j + 1:, j])[0] + j + 1) if L[j + 1:, j].any() else -1 for j in range(n)])
    return counts, parent


def random_pattern(n, density, rng):
    cols = [[] for _ in range(n)]
    for c in range(n):
        for r in range(c):
            if rng.random() < density:
                cols[c].append(r)
        cols[c].append(c)
    col_ptr = np.concatenate([[0], np.cumsum([len(c) for c in cols])]).astype(np.uint64)
    row_idx = np.array([r for c in cols for r in c], np.uint64)
    return col_ptr, row_idx


@pytest.mark.parametrize("n,density,seed", [(1, 0.0, 0), (7, 0.0, 1), (12, 0.3, 2), (40, 0.08, 3), (60, 0.05, 4), (25, 1.0, 5)])
def test_symbolic_against_brute_force(n, density, seed):
    rng = np.random.default_rng(seed)
    col_ptr, row_idx = random_pattern(n, density, rng)
    for order in (None, rng.permutation(n).astype(np.uint64), capi.block_ordering(col_ptr, row_idx)):
        st = capi.block_symbolic_stats(col_ptr, row_idx, order)
        o = np.arange(n) if order is None else order
        counts, parent = brute_force_symbolic(n, col_ptr, row_idx, o)
        assert np.array_equal(st["col_count"].astype(np.int64), counts)
        ref_parent = np.where(parent < 0, np.iinfo(np.uint64).max, parent).astype(np.uint64)
        assert np.array_equal(st["parent"], ref_parent)
        assert st["nnzb_factor"] == int(counts.sum())


def test_postorder_property():
    """the library's ordering is a postorder of its elimination tree: parent[j] > j and subtrees are contiguous"""
    rng = np.random.default_rng(11)
    col_ptr, row_idx = random_pattern(80, 0.04, rng)
    order = capi.block_ordering(col_ptr, row_idx)
    assert sorted(order.tolist()) == list(range(80))
    parent = capi.block_symbolic_stats(col_ptr, row_idx, order)["parent"].astype(np.int64)
    size = np.ones(80, np.int64)
    for j in range(80):
        if parent[j] >= 0:
            assert parent[j] > j
            size[parent[j]] += size[j]
    for j in range(80):  # the subtree of j is exactly the columns j - size + 1 .. j
        lo = j - size[j] + 1
        k = j
        for c in range(lo, j):
            p = c
            while p < j and p >= 0:
                p = parent[p]
            assert p == j, (c, j)
        assert k == j


@pytest.mark.parametrize("name", ["rcs_small", "rcs_mid", "rcs_seq400", "pose_manhattan800"])
def test_fill_against_reference_amd(name):
    d = np.load(os.path.join(GOLDEN, "order_ref.npz"))
    col_ptr, row_idx, ref = d[name + ".col_ptr"], d[name + ".row_idx"], d[name + ".order"]
    n = len(col_ptr) - 1
    own = capi.block_ordering(col_ptr, row_idx)
    assert sorted(own.tolist()) == list(range(n))
    f_ref = capi.block_symbolic_stats(col_ptr, row_idx, ref)
    f_own = capi.block_symbolic_stats(col_ptr, row_idx, own)
    f_nat = capi.block_symbolic_stats(col_ptr, row_idx, None)
    print(f"{name}: factor blocks natural {f_nat['nnzb_factor']}, reference AMD {f_ref['nnzb_factor']}, own {f_own['nnzb_factor']}; "
          f"sum count^2 reference {f_ref['sum_count_sq']:.4g}, own {f_own['sum_count_sq']:.4g}")
    assert f_own["nnzb_factor"] <= 1.15 * f_ref["nnzb_factor"]
    assert f_own["sum_count_sq"] <= 1.3 * f_ref["sum_count_sq"]


def test_invalid_input():
    lib = capi.load_library()
    cp = np.array([0, 1, 2], np.uint64)
    ri = np.array([0, 5], np.uint64)  # row out of range
    o = np.zeros(2, np.uint64)
    assert lib.spp_block_ordering(2, capi._u64p(cp), capi._u64p(ri), capi._u64p(o)) == capi.SPP_ERR_INVALID
    assert lib.spp_block_ordering(2, None, None, None) == capi.SPP_ERR_INVALID
