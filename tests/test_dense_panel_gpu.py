"""The dense front primitive (spp_dense_panel_factor) through the C-ABI: the dataflow kernel against numpy and against the
stream path, and -- the regression test for the stage-release race of round 2 (a consumer warp's last ld.shared of a
pipeline stage overtaken by the stage's refill, wrong factors in about one run in twenty) -- many repetitions of the same
factorisation, which must agree bit for bit: the kernel adds in a fixed order whatever the schedule of its CTAs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _panel(n, m, seed):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, 64))
    a11 = g @ g.T + np.diag(1.0 + 10 * rng.random(n))
    a12 = rng.standard_normal((n, m - n))
    return a11, a12, np.asfortranarray(np.hstack([np.triu(a11), a12]))


@pytest.fixture()
def ctx():
    from slam_plus_plus_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n,m", [(128, 128), (256, 384), (1152, 1152 + 2048), (1280, 1408)])
def test_panel_factor_matches_numpy_and_stream_path(ctx, n, m, monkeypatch):
    a11, a12, panel = _panel(n, m, 3)
    monkeypatch.setenv("SPP_CHOL_DATAFLOW", "1")
    out = ctx.dense_panel_factor(panel)
    r11 = np.triu(out[:, :n])
    assert np.abs(r11.T @ r11 - a11).max() <= 1e-12 * np.abs(a11).max()      # FP64 factor: tolerance 1e-12 relative
    if m > n:
        assert np.abs(r11.T @ out[:, n:] - a12).max() <= 1e-12 * max(1.0, np.abs(a12).max()) * n
    monkeypatch.setenv("SPP_CHOL_DATAFLOW", "0")
    ref = ctx.dense_panel_factor(panel)
    assert np.abs(np.triu(ref[:, :n]) - r11).max() <= 1e-11 * np.abs(r11).max()
    if m > n:
        assert np.abs(ref[:, n:] - out[:, n:]).max() <= 1e-10 * max(1.0, np.abs(out[:, n:]).max())


@pytest.mark.parametrize("n,m,reps", [(2048, 3072, 60), (5248, 5376, 25)])
def test_repeated_factorisations_are_bit_identical(ctx, n, m, reps, monkeypatch):
    monkeypatch.setenv("SPP_CHOL_DATAFLOW", "1")
    _, _, panel = _panel(n, m, 5)
    first = ctx.dense_panel_factor(panel)
    first[:, :n] = np.triu(first[:, :n])
    for k in range(reps):
        out = ctx.dense_panel_factor(panel)
        out[:, :n] = np.triu(out[:, :n])
        assert np.array_equal(out, first), "run %d differs from run 0 (max %.3e)" % (k + 1, np.abs(out - first).max())


def test_panel_factor_reports_a_non_positive_pivot(ctx):
    from slam_plus_plus_b200 import capi
    _, _, panel = _panel(256, 256, 7)
    panel[130, 130] = -1.0
    with pytest.raises(capi.NotPositiveDefinite):
        ctx.dense_panel_factor(panel)


def test_panel_factor_rejects_bad_shapes(ctx):
    from slam_plus_plus_b200 import capi
    with pytest.raises(capi.SppError):
        ctx.dense_panel_factor(np.eye(100))
    with pytest.raises(capi.SppError):
        ctx.dense_panel_factor(np.zeros((256, 128)))
