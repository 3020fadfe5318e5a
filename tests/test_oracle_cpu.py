"""CPU tests (no GPU needed): the C oracle (oracle/spp_oracle.c) against the golden vectors of the unmodified
reference, host-side logic, and the C ABI surface of libspp_b200.so (load + exported symbols only)."""
import ctypes
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, lambda_to_dense, rel_err

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402

CASES = ["ba_tiny", "ba_tiny_interleaved", "ba_small", "ba_small_hard"]
FD_NOISE_TOL = 2e-5  # see tests/test_ba_gpu.py


@pytest.mark.parametrize("name", CASES)
def test_oracle_chi2_matches_reference(name):
    g, d = load_golden(name)
    assert abs(orc.ba_chi2(g) - d["chi2_0"][0]) <= 1e-12 * d["chi2_0"][0]


@pytest.mark.parametrize("name", CASES)
def test_oracle_lambda_matches_reference(name):
    g, d = load_golden(name)
    U, V, W, gc, gp, md = orc.ba_linearise(g)
    cd, cp, ri, vals = orc.lambda_blocks_to_reference_layout(g, U, V, W)
    # block structure of lambda: bit-exact
    assert np.array_equal(cd, d["L0.col_dims"]) and np.array_equal(cp, d["L0.col_ptr"]) and np.array_equal(ri, d["L0.row_idx"])
    A = lambda_to_dense(cd, cp, ri, vals)
    A_ref = lambda_to_dense(d["L0.col_dims"], d["L0.col_ptr"], d["L0.row_idx"], d["L0.vals"])
    A_ref -= d["alpha0"][0] * np.eye(len(A_ref))
    assert rel_err(A, A_ref) < FD_NOISE_TOL
    eta = np.empty(len(d["L0.eta"]))
    off, loc = g.vertex_offsets(), g.vertex_local_index()
    for v in range(g.n_vertices):
        eta[off[v]:off[v + 1]] = gc[loc[v]] if g.vtype[v] == 0 else gp[loc[v]]
    assert rel_err(eta, d["L0.eta"]) < FD_NOISE_TOL
    assert abs(md * 1e-3 - d["alpha0"][0]) <= FD_NOISE_TOL * d["alpha0"][0]


def _split_reference_lambda(d):
    """Reference lambda dump -> (obs_c, obs_p, U, V, W, gc, gp) in Schur order."""
    dims = d["L0.col_dims"].astype(np.int64)
    n = len(dims)
    is_cam = dims == 6
    loc = np.empty(n, np.int64)
    loc[is_cam] = np.arange(is_cam.sum())
    loc[~is_cam] = np.arange((~is_cam).sum())
    base = np.concatenate([[0], np.cumsum(dims)])
    U = np.zeros((is_cam.sum(), 36))
    V = np.zeros(((~is_cam).sum(), 9))
    W, oc, op = [], [], []
    off = 0
    vals = d["L0.vals"]
    for c in range(n):
        for k in range(int(d["L0.col_ptr"][c]), int(d["L0.col_ptr"][c + 1])):
            r = int(d["L0.row_idx"][k])
            sz = int(dims[r] * dims[c])
            blk = vals[off:off + sz]
            off += sz
            if r == c:
                (U if is_cam[c] else V)[loc[c]] = blk
            else:
                m = blk.reshape(int(dims[c]), int(dims[r])).T  # rows x cols
                if not is_cam[r]:
                    m = m.T  # stored 3x6 (point row, camera column) -> 6x3
                W.append(m.T.ravel())
                oc.append(loc[r] if is_cam[r] else loc[c])
                op.append(loc[c] if is_cam[r] else loc[r])
    eta = d["L0.eta"]
    gc = np.stack([eta[base[i]:base[i] + 6] for i in np.flatnonzero(is_cam)])
    gp = np.stack([eta[base[i]:base[i] + 3] for i in np.flatnonzero(~is_cam)])
    return np.array(oc), np.array(op), U, V, np.array(W), gc, gp, is_cam, base


@pytest.mark.parametrize("name", CASES)
def test_oracle_schur_solve_matches_reference(name):
    """The reference's own lambda / eta through the oracle's Schur path -> the reference's dx at 1e-11."""
    g, d = load_golden(name)
    oc, op, U, V, W, gc, gp, is_cam, base = _split_reference_lambda(d)
    rc, dxc, dxp = orc.schur_solve(oc, op, U, V, W, gc, gp, 0.0)
    assert rc == 0
    dx = np.empty(len(d["L0.dx"]))
    for k, i in enumerate(np.flatnonzero(is_cam)):
        dx[base[i]:base[i] + 6] = dxc[k]
    for k, i in enumerate(np.flatnonzero(~is_cam)):
        dx[base[i]:base[i] + 3] = dxp[k]
    assert rel_err(dx, d["L0.dx"]) < 1e-11


@pytest.mark.parametrize("name", CASES)
def test_oracle_lm_matches_reference(name):
    g, d = load_golden(name)
    r = orc.ba_optimize(g, int(d["max_iter"][0]), 0.0)
    tr = d["lm_trace"].reshape(-1, 6)
    assert abs(r["chi2_final"] - d["chi2"][0]) <= 1e-7 * d["chi2"][0]
    for k in range(min(len(tr), r["n_solves"])):
        if abs(tr[k, 1] - tr[k, 2]) <= 1e-6 * tr[k, 1]:
            break  # the accept / reject decision is inside the FD noise from here on
        assert int(r["trace"][k, 4]) == int(tr[k, 4])
        assert abs(r["trace"][k, 2] - tr[k, 2]) <= FD_NOISE_TOL * tr[k, 2]


def test_oracle_dense_llt():
    rng = np.random.default_rng(1)
    M = rng.normal(size=(40, 40))
    A = M @ M.T + 40 * np.eye(40)
    b = rng.normal(size=40)
    rc, x = orc.dense_llt_solve(A, b)
    assert rc == 0 and rel_err(x, np.linalg.solve(A, b)) < 1e-12
    A[7, 7] = -1
    rc, _ = orc.dense_llt_solve(A, b)
    assert rc == 1


def test_oracle_pose_composition_identity():
    v = np.array([1.0, -2.0, 0.5, 0.3, -0.2, 0.9])
    assert rel_err(orc.relative_to_absolute(v, np.zeros(6)), v) < 1e-15
    # composing with a rotation about the own axis adds the angles
    w = np.array([0, 0, 0, 0.3, -0.2, 0.9]) * 0.1
    out = orc.relative_to_absolute(v, w)
    assert rel_err(out[3:], v[3:] * 1.1) < 1e-14


# ---- host-side logic -------------------------------------------------------------------------------------

def test_graph_io_roundtrip(tmp_path):
    from slam_plus_plus_b200 import graphs, sppio
    g = graphs.ba_shape("tiny", interleave_ids=True, shuffle_edges=True)
    p = str(tmp_path / "g.bin")
    sppio.write_graph(p, g)
    h = sppio.read_graph(p)
    for a, b in ((g.vtype, h.vtype), (g.cams, h.cams), (g.pts, h.pts), (g.obs_pt, h.obs_pt), (g.obs_cam, h.obs_cam),
                 (g.z, h.z), (g.info, h.info)):
        assert np.array_equal(a, b)
    d = {"a": np.arange(5, dtype=np.float64), "b.c": np.arange(3, dtype=np.uint64)}
    sppio.write_dump(str(tmp_path / "d.dump"), d)
    e = sppio.read_dump(str(tmp_path / "d.dump"))
    assert all(np.array_equal(d[k], e[k]) for k in d)


def test_generators_are_seeded_and_shaped():
    from slam_plus_plus_b200 import graphs
    a, b = graphs.ba_shape("small"), graphs.ba_shape("small")
    assert np.array_equal(a.z, b.z) and np.array_equal(a.obs_cam, b.obs_cam)
    # no landmark is observed twice by one camera, every track has at least two observations
    key = a.obs_pt.astype(np.int64) * a.n_vertices + a.obs_cam
    assert len(np.unique(key)) == len(key)
    assert np.bincount(a.obs_pt - a.n_cams).min() >= 2
    assert graphs.BA_SHAPES["venice871"][:2] == (871, 530304)


# ---- C ABI surface -----------------------------------------------------------------------------------------

def test_library_exports_every_declared_symbol():
    from slam_plus_plus_b200 import capi
    lib = capi.load_library()
    header = open(os.path.join(ROOT, "include", "spp_b200.h")).read()
    import re
    declared = sorted(set(re.findall(r"\b(spp_[a-z0-9_]+)\s*\(", header)) - {"spp_allreduce_fn"})
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/spp_b200.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == declared


def test_no_cpu_fallback_without_gpu():
    """Without a usable sm_100 device the library must refuse to create a context (no silent CPU path)."""
    from slam_plus_plus_b200 import capi
    lib = capi.load_library()
    h = ctypes.c_void_p()
    rc = lib.spp_create(0, ctypes.byref(h))
    if rc == 0:  # a GPU is present (GPU box): nothing to check here
        lib.spp_destroy(h)
        pytest.skip("GPU present")
    assert rc == capi.SPP_ERR_CUDA
    assert b"no CPU fallback" in lib.spp_last_error(None) or b"sm_100" in lib.spp_last_error(None)


def test_product_does_not_import_the_oracle():
    """The product package must not reference oracle/ (the oracle is test infrastructure only)."""
    pkg = os.path.join(ROOT, "slam_plus_plus_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "spp_oracle" not in text and "import oracle" not in text and "oracle/" not in text.replace("oracle/spp_dump.h", ""), f
