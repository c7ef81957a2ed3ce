"""BASELINE config 4 at FULL size (n = 50 000, d = 8, sqrexp l = 1, noise 0.01) against the committed oracle sample
tests/golden/c4_full_sample.npz (tests/golden/make_c4_golden.py: the oracle's arithmetic with SciPy dpotrf in place at
n = 50 000, first 256 test points of bench.py's generator).  Replaces R/GPRclass.R:138-164 at the size the headline
metric is quoted on; tolerances are the north star's: 1e-9 on mean / variance, 1e-8 on logp."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c4_full_sample.npz")


@pytest.fixture(scope="module")
def c4(gprc, ctx):
    import bench
    z = np.load(GOLDEN)
    n, m, d = int(z["n"]), int(z["m"]), int(z["d"])
    X, y, Xs = bench.make_inputs(n, m, d)
    np.testing.assert_array_equal(Xs[:, :int(z["M"])], z["Xs_head"])     # same generator, same bits
    model = gprc.GPR(X, y, float(z["noise"]), gprc.cov_func(gprc.sqrexp, l=float(z["l"])), ctx=ctx)
    return z, model, Xs


def test_logp_alpha_and_factor_at_full_size(c4):
    z, model, _ = c4
    assert abs(model.logp[0, 0] - float(z["logp"])) <= 1e-8 * abs(float(z["logp"]))
    a = model.alpha
    assert np.max(np.abs(a[:64] - z["alpha_head"])) <= 1e-9 * np.max(np.abs(z["alpha_head"]))


@pytest.mark.parametrize("variant", [2, 1, 64])
def test_int8_pass_at_full_size_against_the_oracle(gprc, ctx, c4, variant):
    """path 4 forced on the oracle's 256 test points padded to one full wave (the tail of the chunk repeats them)"""
    z, model, Xs = c4
    M = int(z["M"])
    pts = np.concatenate([Xs[:, :M], Xs[:, M:148 * 64]], axis=1)
    ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 4)
    ctx.set_option(gprc._lib.OPT_INT8_TILE, variant)
    try:
        out = model.predict(pts)
        assert ctx.last_predict_path() == 4
    finally:
        ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 0)
        ctx.set_option(gprc._lib.OPT_INT8_TILE, gprc._lib.INT8_TILE_DEFAULT)
    assert np.max(np.abs(out[:M, 0] - z["mean"])) <= 1e-9 * np.max(np.abs(z["mean"]))
    assert np.max(np.abs(out[:M, 1] - z["var"]) / np.maximum(np.abs(z["var"]), 1.0)) <= 1e-9


def test_fp64_substitution_at_full_size_against_the_oracle(gprc, ctx, c4):
    z, model, Xs = c4
    M = int(z["M"])
    ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 2)
    try:
        out = model.predict(Xs[:, :M])
    finally:
        ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 0)
    assert np.max(np.abs(out[:, 0] - z["mean"])) <= 1e-9 * np.max(np.abs(z["mean"]))
    assert np.max(np.abs(out[:, 1] - z["var"]) / np.maximum(np.abs(z["var"]), 1.0)) <= 1e-9
