"""alpha = solve(t(L), solve(L, y)) (R/GPRclass.R:152) through the dataflow substitution kernel (GPRC_OPT_TRSV = 1,
csrc/trsv.cuh: one CTA per row block, release / acquire flags instead of grid-wide barriers) against the oracle and
against the cooperative sweeps, at sizes around the padding and grid boundaries (1 block, fewer blocks than CTAs, more
blocks than CTAs) and inside the GPC Newton loop (R/GPCclass.R:82-83)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture
def flow(gprc, ctx):
    ctx.set_option(gprc._lib.OPT_TRSV, 1)
    yield
    ctx.set_option(gprc._lib.OPT_TRSV, gprc._lib.TRSV_DEFAULT)


@pytest.mark.parametrize("n", [1, 127, 129, 1000, 5000])
def test_alpha_and_logp_match_the_oracle(gprc, oracle, ctx, flow, n):
    rng = np.random.default_rng(60 + n)
    X = rng.uniform(-2, 2, (3, n))
    y = np.sum(np.sin(X), axis=0) + rng.normal(0, 0.1, n)
    g = gprc.GPR(X, y, 0.05, gprc.cov_func(gprc.rationalquadratic, l=0.9, alpha=1.2), ctx=ctx)
    o = oracle.GPR(X, y, 0.05, oracle.cov_func(oracle.rationalquadratic, l=0.9, alpha=1.2))
    assert abs(g.logp[0, 0] - float(o.logp)) <= 1e-8 * abs(float(o.logp))
    assert np.max(np.abs(g.alpha - o.alpha)) <= 1e-9 * np.max(np.abs(o.alpha))


def test_more_row_blocks_than_ctas_and_agreement_with_the_cooperative_sweeps(gprc, ctx):
    """n = 40 000: 313 row blocks on at most 296 resident CTAs, so CTAs take a second block.  Checked through
    (K + noise I) alpha = y on sampled rows and against the cooperative sweeps (different summation order: 1e-11)."""
    rng = np.random.default_rng(61)
    n, D = 40000, 6
    X = rng.uniform(-1, 1, (D, n))
    y = np.sum(np.sin(3 * X), axis=0) + rng.normal(0, 0.1, n)
    k = gprc.cov_func(gprc.sqrexp, l=1.0)
    out = {}
    for mode in (1, 0):
        ctx.set_option(gprc._lib.OPT_TRSV, mode)
        try:
            g = gprc.GPR(X, y, 0.01, k, ctx=ctx)
            out[mode] = (g.alpha.copy(), g.logp[0, 0])
        finally:
            ctx.set_option(gprc._lib.OPT_TRSV, gprc._lib.TRSV_DEFAULT)
        del g
    a1, lp1 = out[1]
    a0, lp0 = out[0]
    assert abs(lp1 - lp0) <= 1e-11 * abs(lp0)
    assert np.max(np.abs(a1 - a0)) <= 1e-9 * np.max(np.abs(a0))
    idx = rng.integers(0, n, 64)
    Krows = np.exp(-0.5 * np.sum((X[:, idx, None] - X[:, None, :]) ** 2, axis=0))
    Krows[np.arange(64), idx] += 0.01
    assert np.max(np.abs(Krows @ a1 - y[idx])) < 1e-8 * np.max(np.abs(a1))


def test_gpc_newton_loop_on_the_dataflow_kernel(gprc, oracle, ctx, flow):
    c2 = oracle.make_config("C2", n=700, m=200)
    gc = gprc.GPC(c2["X"], c2["y"], gprc.cov_func(gprc.sqrexp, l=0.3), verbose=False, ctx=ctx)
    oc = oracle.GPC(c2["X"], c2["y"], oracle.cov_func(oracle.sqrexp, l=0.3))
    assert gc.iterations == oc.iterations
    np.testing.assert_allclose(gc.objective_trace, oc.objective_trace[:len(gc.objective_trace)], rtol=1e-9)
    assert np.array_equal(gc.predict_class(c2["Xs"]) >= 0.5, oc.predict_class(c2["Xs"]) >= 0.5)
