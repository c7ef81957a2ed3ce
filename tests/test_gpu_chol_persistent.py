"""t(chol(K + noise I)) (R/GPRclass.R:142) through the persistent tile kernel (GPRC_OPT_CHOL_TILES, csrc/potrf.cuh:
chol_persistent_kernel -- one launch, tile tasks, per-row progress counters) against the oracle and against the
launch-sequence factorisation, including the not-positive-definite report and the GPC Newton loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture
def tiles(gprc, ctx):
    ctx.set_option(gprc._lib.OPT_CHOL_TILES, 128)
    yield
    ctx.set_option(gprc._lib.OPT_CHOL_TILES, gprc._lib.CHOL_TILES_DEFAULT)


@pytest.mark.parametrize("n", [129, 256, 1000, 2048, 3000])
def test_factor_alpha_logp_match_the_oracle(gprc, oracle, ctx, tiles, n):
    rng = np.random.default_rng(70 + n)
    X = rng.uniform(-2, 2, (3, n))
    y = np.sum(np.sin(X), axis=0) + rng.normal(0, 0.1, n)
    g = gprc.GPR(X, y, 0.05, gprc.cov_func(gprc.gammaexp, l=1.1, gamma=1.3), ctx=ctx)
    o = oracle.GPR(X, y, 0.05, oracle.cov_func(oracle.gammaexp, l=1.1, gamma=1.3))
    assert abs(g.logp[0, 0] - float(o.logp)) <= 1e-8 * abs(float(o.logp))
    assert np.max(np.abs(g.alpha - o.alpha)) <= 1e-9 * np.max(np.abs(o.alpha))
    L = g.L
    assert np.max(np.abs(L - np.asarray(o.L))) <= 1e-10 * np.max(np.abs(o.L))
    assert np.all(np.triu(L, 1) == 0.0)            # explicit zeros above the diagonal, like t(chol(.))
    Xs = rng.uniform(-2, 2, (3, 200))
    got, ref = g.predict(Xs), o.predict(Xs)
    assert np.max(np.abs(got[:, 0] - ref[:, 0])) <= 1e-9 * np.max(np.abs(ref[:, 0]))
    assert np.max(np.abs(got[:, 1] - ref[:, 1])) <= 1e-9


@pytest.mark.parametrize("n", [5000, 8192, 16384])
def test_agrees_with_the_launch_sequence_factorisation(gprc, ctx, n):
    rng = np.random.default_rng(71)
    X = rng.uniform(-1, 1, (6, n))
    y = np.sum(np.sin(3 * X), axis=0) + rng.normal(0, 0.1, n)
    k = gprc.cov_func(gprc.sqrexp, l=1.0)
    out = {}
    for tiles_ in (128, 0):
        ctx.set_option(gprc._lib.OPT_CHOL_TILES, tiles_)
        try:
            g = gprc.GPR(X, y, 0.01, k, ctx=ctx)
            out[tiles_] = (g.alpha.copy(), g.logp[0, 0])
        finally:
            ctx.set_option(gprc._lib.OPT_CHOL_TILES, gprc._lib.CHOL_TILES_DEFAULT)
        del g
    (a1, lp1), (a0, lp0) = out[128], out[0]
    assert abs(lp1 - lp0) <= 1e-11 * abs(lp0)
    assert np.max(np.abs(a1 - a0)) <= 1e-8 * np.max(np.abs(a0))     # cond(K + 0.01 I) ~ 1e6: two summation orders
    idx = rng.integers(0, n, 64)
    Krows = np.exp(-0.5 * np.sum((X[:, idx, None] - X[:, None, :]) ** 2, axis=0))
    Krows[np.arange(64), idx] += 0.01
    assert np.max(np.abs(Krows @ a1 - y[idx])) < 1e-8 * np.max(np.abs(a1))


def test_first_bad_pivot_is_reported(gprc, ctx, tiles):
    """LAPACK's info: index of the first non-positive pivot; the host turns it into the noise bump / the error of
    R/GPRclass.R:149"""
    rng = np.random.default_rng(72)
    n = 700
    B = rng.normal(size=(n, 5))
    K = B @ B.T                                  # rank 5: singular
    K[300, 300] -= 10.0                           # and indefinite from row 300 on at the latest
    with pytest.raises(ValueError, match="non positive definite"):
        gprc.GPR(np.arange(n, dtype=float)[None, :], rng.normal(size=n), 0.0,
                 lambda x, y: K[np.asarray(x[0], dtype=int), np.asarray(y[0], dtype=int)], ctx=ctx)


def test_gpc_newton_loop(gprc, oracle, ctx, tiles):
    c2 = oracle.make_config("C2", n=900, m=300)
    gc = gprc.GPC(c2["X"], c2["y"], gprc.cov_func(gprc.sqrexp, l=0.3), verbose=False, ctx=ctx)
    oc = oracle.GPC(c2["X"], c2["y"], oracle.cov_func(oracle.sqrexp, l=0.3))
    assert gc.iterations == oc.iterations
    np.testing.assert_allclose(gc.objective_trace, oc.objective_trace[:len(gc.objective_trace)], rtol=1e-9)
    assert np.array_equal(gc.predict_class(c2["Xs"]) >= 0.5, oc.predict_class(c2["Xs"]) >= 0.5)
