"""GPR parity: the reference's known-answer tests (tests/testthat/test-gpr.R) verbatim, then randomised differential
tests against the oracle at the north_star tolerances: relative 1e-9 on predictive mean/variance, 1e-8 on logp."""
import math
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MEAN_RTOL = 1e-9
LOGP_RTOL = 1e-8


def assert_mean_var(got, ref, kss):
    """|d mean| <= 1e-9 max(|mean|, scale); |d var| <= 1e-9 max(|var|, k(x*, x*)) (cancellation-limited,
    BASELINE.md section 3)."""
    scale = np.maximum(np.abs(ref[:, 0]), np.max(np.abs(ref[:, 0])))
    assert np.all(np.abs(got[:, 0] - ref[:, 0]) <= MEAN_RTOL * scale)
    assert np.all(np.abs(got[:, 1] - ref[:, 1]) <= MEAN_RTOL * np.maximum(np.abs(ref[:, 1]), kss))


def test_known_answers_test_gpr_R(gprc):
    # tests/testthat/test-gpr.R:6-9
    g1 = gprc.GPR.polynomial.new(np.array([[-0.5, 0.5]]), np.array([4.0, 4.0]), 0.5, 0.25, 1)
    np.testing.assert_allclose(g1.predict(0.0), [[2.0, 1 / 8]], rtol=0, atol=1.5e-8)
    # :12-15
    g2 = gprc.GPR.constant.new(np.array([[1.0, 2.0]]), np.array([1.0, 3.0]), 1, 1)
    np.testing.assert_allclose(g2.predict(3.0), [[4 / 3, 1 / 3]], rtol=0, atol=1.5e-8)
    # :16-19
    g3 = gprc.GPR.constant.new(np.array([[100.0, 54.0]]), np.array([5.0, 0.0]), 1, 1)
    np.testing.assert_allclose(g3.predict(math.pi), [[5 / 3, 1 / 3]], rtol=0, atol=1.5e-8)
    # :23-27
    g4 = gprc.GPR.sqrexp.new(np.array([[1.0, 2.0]]), np.array([0.0, 1.0]), 1, 1)
    e = math.exp
    cov = 1 - (2 * e(-1) - 2 * e(-3) + 2 * e(-4)) / (4 - e(-1))
    np.testing.assert_allclose(g4.predict(0.0), [[(2 * e(-2) - e(-1)) / (4 - e(-1)), cov]], rtol=0, atol=1.5e-8)
    # regression values of the restatement
    for g, lp in ((g1, -17.8378770664), (g2, -4.72051654408), (g3, -10.7205165441), (g4, -2.75810665086)):
        assert g.logp.shape == (1, 1)
        assert abs(g.logp[0, 0] - lp) < 1e-9


def test_read_only_bindings(gprc):
    g = gprc.GPR.sqrexp.new(np.array([[1.0, 2.0]]), np.array([0.0, 1.0]), 1, 1)
    for name in ("X", "k", "y", "noise", "L", "alpha", "logp"):
        with pytest.raises(AttributeError, match=r"`\$%s` is read only" % name):
            setattr(g, name, 1)


CASES = [
    ("sqrexp", dict(l=1.0), 1, 200, 1000, 0.01),        # BASELINE config 1 scale
    ("sqrexp", dict(l=1.0), 8, 900, 700, 0.01),         # config 4's kernel at oracle-sized n
    ("gammaexp", dict(l=1.0, gamma=1.5), 8, 515, 300, 0.1),
    ("rationalquadratic", dict(l=1.0, alpha=1.0), 4, 640, 129, 0.05),
    ("polynomial", dict(sigma=1.0, p=3.0), 8, 300, 257, 0.1),
    ("linear", dict(sigma=0.7), 3, 129, 50, 0.5),
    ("constant", dict(c=2.0), 2, 64, 10, 1.0),
]


@pytest.mark.parametrize("name,params,D,n,m,noise", CASES)
def test_gpr_matches_oracle(gprc, oracle, name, params, D, n, m, noise):
    rng = np.random.default_rng(n + m)
    lim = 6.0 if D == 1 else 1.0
    X = rng.uniform(-lim, lim, (D, n))
    y = np.sum(np.sin(X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-lim, lim, (D, m))
    Xs[:, 0] = X[:, 0]  # a test point on a training point: var << k**
    g = gprc.GPR(X, y, noise, gprc.cov_func(getattr(gprc, name), **params))
    o = oracle.GPR(X, y, noise, oracle.cov_func(getattr(oracle, name), **params))
    assert g.noise == o.noise
    assert abs(g.logp[0, 0] - o.logp) <= LOGP_RTOL * abs(o.logp)
    got, ref = g.predict(Xs), o.predict(Xs)
    assert got.shape == (m, 2)
    kss = oracle.cov_func(getattr(oracle, name), **params)(Xs, Xs)
    assert_mean_var(got, ref, kss)
    # unpinned by the reference, pinned by the restatement: the factor and alpha
    np.testing.assert_allclose(g.L, o.L, rtol=0, atol=1e-10 * np.max(np.abs(o.L)))
    assert np.all(np.triu(g.L, 1) == 0)
    np.testing.assert_allclose(g.alpha, o.alpha, rtol=0, atol=1e-8 * np.max(np.abs(o.alpha)))


def test_predict_vector_reshape_rule(gprc, oracle):
    # a bare vector is D x (length / D), column-major (R/GPRclass.R:157-159)
    rng = np.random.default_rng(3)
    X = rng.uniform(-1, 1, (2, 50))
    y = np.sin(X[0]) + X[1]
    g = gprc.GPR(X, y, 0.1, gprc.cov_func(gprc.sqrexp, l=1.0))
    o = oracle.GPR(X, y, 0.1, oracle.cov_func(oracle.sqrexp, l=1.0))
    flat = rng.uniform(-1, 1, 14)
    np.testing.assert_allclose(g.predict(flat), o.predict(flat), rtol=1e-9, atol=1e-12)
    with pytest.raises(ValueError):
        g.predict(np.zeros(3))


def test_predict_rejects_a_matrix_with_the_wrong_number_of_rows(gprc):
    """a (1, m) matrix with D = 2 passes length(X_star) %% nrow(X) == 0 (R/GPRclass.R:156) and must fail like the
    reference's covariance_matrix does, never reach the library (which would read D * m doubles from it)"""
    rng = np.random.default_rng(3)
    g = gprc.GPR(rng.uniform(-1, 1, (2, 30)), rng.normal(size=30), 0.1, gprc.cov_func(gprc.sqrexp, l=1.0))
    for bad in (np.zeros((1, 8)), np.zeros((3, 4))):
        with pytest.raises(ValueError, match="non-conformable"):
            g.predict(bad)
        with pytest.raises(ValueError, match="non-conformable"):
            g.predict(bad, pointwise_var=False)
    assert g.predict(np.zeros(8)).shape == (4, 2)          # a bare vector is reshaped to D rows (R/GPRclass.R:157-159)


def test_predict_full_covariance(gprc, oracle):
    rng = np.random.default_rng(4)
    X = rng.uniform(-5, 5, (1, 150))
    y = 0.1 * X[0] ** 3 + rng.normal(0, 0.1, 150)
    Xs = np.linspace(-6, 6, 203)
    g = gprc.GPR(X, y, 0.1, gprc.cov_func(gprc.sqrexp, l=1.0))
    o = oracle.GPR(X, y, 0.1, oracle.cov_func(oracle.sqrexp, l=1.0))
    mean, cov = g.predict(Xs, pointwise_var=False)
    omean, ocov = o.predict(Xs, pointwise_var=False)
    assert mean.shape == (203, 1) and cov.shape == (203, 203)
    np.testing.assert_allclose(mean, omean, rtol=0, atol=1e-9 * np.max(np.abs(omean)))
    np.testing.assert_allclose(cov, ocov, rtol=0, atol=1e-9)


def test_noise_bump_schedule(gprc, oracle):
    # duplicated points with zero noise: chol() fails, the constructor retries with noise + 0.01 i (R/GPRclass.R:141-148)
    X = np.array([[0.0, 0.0, 1.0, 1.0, 2.0]])
    y = np.array([1.0, 1.0, 2.0, 2.0, 0.5])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        g = gprc.GPR(X, y, 0.0, gprc.cov_func(gprc.sqrexp, l=1.0))
    o = oracle.GPR(X, y, 0.0, oracle.cov_func(oracle.sqrexp, l=1.0))
    assert g.noise == o.noise == 0.01
    assert any("Noise got changed to 0.01" in str(x.message) for x in w)
    np.testing.assert_allclose(g.predict(np.array([0.5])), o.predict(np.array([0.5])), rtol=1e-7)


def test_not_positive_definite_raises(gprc):
    # a negative "kernel": no noise of the schedule rescues it (R/GPRclass.R:149)
    bad = lambda x, y: -5.0 * np.ones(x.shape[1])
    with pytest.raises(ValueError, match="non positive definite"):
        gprc.GPR(np.array([[0.0, 1.0, 2.0]]), np.array([1.0, 2.0, 3.0]), 0.0, bad)


def test_closure_kernel_precomputed_path(gprc, oracle):
    kappa = lambda x, y: np.exp(-3 * np.sum((x - y) ** 2, axis=0))
    rng = np.random.default_rng(9)
    X = rng.uniform(-1, 1, (2, 140))
    y = rng.standard_normal(140)
    Xs = rng.uniform(-1, 1, (2, 77))
    g = gprc.GPR(X, y, 0.2, kappa)
    o = oracle.GPR(X, y, 0.2, kappa)
    assert abs(g.logp[0, 0] - o.logp) <= LOGP_RTOL * abs(o.logp)
    assert_mean_var(g.predict(Xs), o.predict(Xs), np.ones(77))


def test_large_size_properties(gprc):
    """Size-independent properties at a size the oracle cannot do in seconds (n = 8192, d = 8):
    L L' reproduces K + noise I on sampled entries; alpha solves the system; var in [0, k**]."""
    rng = np.random.default_rng(11)
    n, D, m = 8192, 8, 4096
    X = rng.uniform(-1, 1, (D, n))
    y = np.sum(np.sin(np.pi * X), axis=0) + rng.normal(0, 0.1, n)
    k = gprc.cov_func(gprc.sqrexp, l=1.0)
    g = gprc.GPR(X, y, 0.01, k)
    L = g.L
    idx = rng.integers(0, n, 300)
    jdx = rng.integers(0, n, 300)
    Kij = np.array([gprc.sqrexp(X[:, i], X[:, j], 1.0) + (0.01 if i == j else 0.0) for i, j in zip(idx, jdx)])
    LLt = np.einsum("ij,ij->i", L[idx, :], L[jdx, :])
    assert np.max(np.abs(LLt - Kij)) < 1e-11
    # residual of (K + noise I) alpha = y on sampled rows
    Krows = gprc.covariance_matrix(X[:, idx[:50]], X, k).copy()
    Krows[np.arange(50), idx[:50]] += 0.01
    assert np.max(np.abs(Krows @ g.alpha - y[idx[:50]])) < 1e-8 * np.max(np.abs(g.alpha))
    pred = g.predict(rng.uniform(-1, 1, (D, m)))
    assert np.all(pred[:, 1] > -1e-9) and np.all(pred[:, 1] <= 1.0 + 1e-12)
    # predicting at training points reproduces y up to the noise shrinkage: mean = y - noise * alpha
    ptrain = g.predict(X[:, :256])
    np.testing.assert_allclose(ptrain[:, 0], y[:256] - 0.01 * g.alpha[:256], rtol=0, atol=1e-8)


@pytest.mark.parametrize("path", [1, 2, 3, 4])
def test_variance_pass_paths_agree_with_oracle(gprc, oracle, ctx, path):
    """v = L^-1 K_star either with the explicit inverse (one triangular GEMM), by blocked substitution in FP64, or by
    the substitution with its products on the INT8 tensor cores (path 4, ozaki.cuh)."""
    rng = np.random.default_rng(31)
    n, m, D = 700, 450, 5
    X = rng.uniform(-1, 1, (D, n))
    y = np.sum(np.cos(X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-1, 1, (D, m))
    ctx.set_option(gprc._lib.OPT_PREDICT_PATH, path)
    try:
        got = gprc.GPR(X, y, 0.05, gprc.cov_func(gprc.rationalquadratic, l=0.8, alpha=1.5), ctx=ctx).predict(Xs)
    finally:
        ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 0)
    ok = oracle.cov_func(oracle.rationalquadratic, l=0.8, alpha=1.5)
    ref = oracle.GPR(X, y, 0.05, ok).predict(Xs)
    assert_mean_var(got, ref, ok(Xs, Xs))


@pytest.mark.parametrize("digits,tile,tol", [(6, 64, 1e-10), (7, 64, 1e-12), (8, 64, 1e-12), (7, 128, 1e-12), (6, 128, 1e-10),
                                             (8, 128, 1e-12), (6, 2, 1e-10), (7, 2, 1e-12), (8, 2, 1e-12), (6, 1, 1e-10),
                                             (7, 1, 1e-12), (8, 1, 1e-12)])
def test_int8_substitution_matches_fp64_substitution(gprc, ctx, digits, tile, tol):
    """Path 4 against path 2 on the same factor: 1500 training points (12 block rows), polynomial kernel (k** varies
    per test point, so the per-point exponents differ), more than one 64-point tile per SM."""
    rng = np.random.default_rng(35)
    n, m, D = 1500, 148 * 64 + 333, 6
    X = rng.uniform(-1, 1, (D, n))
    y = np.sum(X ** 2, axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-1.5, 1.5, (D, m))
    k = gprc.cov_func(gprc.polynomial, sigma=1.0, p=3.0)
    g = gprc.GPR(X, y, 0.1, k, ctx=ctx)
    out = {}
    for path in (2, 4):
        ctx.set_option(gprc._lib.OPT_PREDICT_PATH, path)
        ctx.set_option(gprc._lib.OPT_OZAKI_DIGITS, digits)
        ctx.set_option(gprc._lib.OPT_INT8_TILE, tile)
        try:
            out[path] = g.predict(Xs)
        finally:
            ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 0)
            ctx.set_option(gprc._lib.OPT_OZAKI_DIGITS, 7)
            ctx.set_option(gprc._lib.OPT_INT8_TILE, gprc._lib.INT8_TILE_DEFAULT)
    kss = (np.sum(Xs * Xs, axis=0) + 1.0) ** 3
    np.testing.assert_allclose(out[2][:, 0], out[4][:, 0], rtol=1e-13, atol=0)  # the mean does not go through the variance pass
    assert np.max(np.abs(out[2][:, 1] - out[4][:, 1]) / kss) < tol


def test_persistent_and_multi_launch_substitution_are_bitwise_equal(gprc, ctx):
    rng = np.random.default_rng(34)
    n, m = 520, 148 * 128 + 300   # 151 tiles on 148 SMs: block rows of neighbouring tiles overlap in the persistent kernel
    X = rng.uniform(-3, 3, (2, n))
    y = np.sin(X[0]) * X[1] + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-3, 3, (2, m))
    out = []
    for path in (2, 3):
        ctx.set_option(gprc._lib.OPT_PREDICT_PATH, path)
        try:
            out.append(gprc.GPR(X, y, 0.05, gprc.cov_func(gprc.sqrexp, l=1.0), ctx=ctx).predict(Xs))
        finally:
            ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 0)
    np.testing.assert_array_equal(out[0], out[1])


def test_large_predict_takes_the_substitution_path(gprc, oracle, ctx):
    # >= 148 * 128 test points and no inverse yet: automatic choice = blocked substitution, whole-wave chunks
    rng = np.random.default_rng(32)
    n, m = 300, 148 * 128 + 77
    X = rng.uniform(-6, 6, (1, n))
    y = 0.1 * X[0] ** 3 + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-6, 6, (1, m))
    g = gprc.GPR(X, y, 0.01, gprc.cov_func(gprc.sqrexp, l=1.0), ctx=ctx)
    ctx.reset_timers()
    got = g.predict(Xs)
    timers, launches = ctx.timers()
    assert timers["trtri"] == 0.0 and timers["var"] > 0.0
    ref = oracle.GPR(X, y, 0.01, oracle.cov_func(oracle.sqrexp, l=1.0)).predict(Xs)
    assert_mean_var(got, ref, np.ones(m))


@pytest.mark.parametrize("n,m", [(1, 1), (2, 3), (127, 5), (128, 128), (129, 1), (513, 257)])
def test_edge_sizes(gprc, oracle, n, m):
    # padding boundaries of the 128-blocked kernels and degenerate sizes
    rng = np.random.default_rng(n * 7 + m)
    X = rng.uniform(-2, 2, (3, n))
    y = rng.standard_normal(n)
    Xs = rng.uniform(-2, 2, (3, m))
    g = gprc.GPR(X, y, 0.3, gprc.cov_func(gprc.gammaexp, l=1.2, gamma=1.3))
    ok = oracle.cov_func(oracle.gammaexp, l=1.2, gamma=1.3)
    o = oracle.GPR(X, y, 0.3, ok)
    assert abs(g.logp[0, 0] - o.logp) <= LOGP_RTOL * abs(o.logp)
    assert_mean_var(g.predict(Xs), o.predict(Xs), ok(Xs, Xs))
    assert g.L.shape == (n, n) and g.alpha.shape == (n,)


def test_linear_kernel_with_per_dimension_sigma(gprc, oracle):
    rng = np.random.default_rng(41)
    X = rng.standard_normal((3, 90))
    y = X[0] - 2 * X[2] + rng.normal(0, 0.1, 90)
    sig = np.array([0.5, 2.0, 1.25])
    g = gprc.GPR.linear.new(X, y, 0.2, sig)
    o = oracle.GPR(X, y, 0.2, oracle.cov_func(oracle.linear, sigma=sig))
    Xs = rng.standard_normal((3, 40))
    ok = oracle.cov_func(oracle.linear, sigma=sig)
    assert_mean_var(g.predict(Xs), o.predict(Xs), np.abs(ok(Xs, Xs)))
    with pytest.raises(ValueError):
        gprc.GPR.linear.new(X, y, 0.2, np.array([1.0, 2.0]))  # stopifnot(length(sigma) == nrow(X))


@pytest.mark.parametrize("name,params,noise", [("sqrexp", dict(l=1.0), 0.01), ("rationalquadratic", dict(l=1.0, alpha=1.0), 0.05),
                                               ("polynomial", dict(sigma=1.0, p=3.0), 0.1)])
def test_gpr_with_tensor_core_build_matches_oracle(gprc, oracle, ctx, name, params, noise):
    """End-to-end tolerance with K and K_star built by the DMMA Gram-tile kernel (norm expansion)."""
    rng = np.random.default_rng(51)
    X = rng.uniform(-1, 1, (8, 600))
    y = np.sum(np.sin(np.pi * X), axis=0) + rng.normal(0, 0.1, 600)
    Xs = rng.uniform(-1, 1, (8, 333))
    Xs[:, 0] = X[:, 0]
    ctx.set_option(gprc._lib.OPT_GRAM_DMMA, 2)
    try:
        g = gprc.GPR(X, y, noise, gprc.cov_func(getattr(gprc, name), **params), ctx=ctx)
        got = g.predict(Xs)
    finally:
        ctx.set_option(gprc._lib.OPT_GRAM_DMMA, 1)
    ok = oracle.cov_func(getattr(oracle, name), **params)
    o = oracle.GPR(X, y, noise, ok)
    assert abs(g.logp[0, 0] - o.logp) <= LOGP_RTOL * abs(o.logp)
    assert_mean_var(got, o.predict(Xs), ok(Xs, Xs))


GOLD = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "golden_v1.npz"))


@pytest.mark.parametrize("tag", ["c1", "c4s", "rq", "gx", "poly"])
def test_gpu_path_matches_committed_golden_vectors(gprc, tag):
    """The CUDA path against the committed snapshot (tests/golden/golden_v1.npz), independent of the live oracle."""
    import json
    name = str(GOLD[tag + "_kernel"])
    params = json.loads(str(GOLD[tag + "_params"]))
    g = gprc.GPR(GOLD[tag + "_X"], GOLD[tag + "_y"], float(GOLD[tag + "_noise"]), gprc.cov_func(getattr(gprc, name), **params))
    ref = GOLD[tag + "_pred"]
    got = g.predict(GOLD[tag + "_Xs"])
    kss = np.asarray(gprc.cov_func(getattr(gprc, name), **params)(GOLD[tag + "_Xs"], GOLD[tag + "_Xs"]))
    assert_mean_var(got, ref, np.abs(kss))
    assert abs(g.logp[0, 0] - float(GOLD[tag + "_logp"])) <= LOGP_RTOL * abs(float(GOLD[tag + "_logp"]))
    np.testing.assert_allclose(g.alpha, GOLD[tag + "_alpha"], rtol=0, atol=1e-8 * np.max(np.abs(GOLD[tag + "_alpha"])))


def test_long_predict_polls_the_interrupt_callback(gprc, ctx):
    """gprc_ctx_set_interrupt: polled between chunks of test points; a non-zero answer abandons the call
    with status -8 and leaves the context usable."""
    import ctypes as C
    rng = np.random.default_rng(39)
    n, m = 300, (1 << 20) + 5000          # the chunk capacity is capped at 2^20 test points: two chunks
    X = rng.uniform(-6, 6, (1, n))
    y = 0.1 * X[0] ** 3 + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-6, 6, (1, m))
    g = gprc.GPR(X, y, 0.01, gprc.cov_func(gprc.sqrexp, l=1.0), ctx=ctx)
    calls = []
    cb_type = C.CFUNCTYPE(C.c_int, C.c_void_p)
    cb = cb_type(lambda user: (calls.append(1), 1)[1])
    gprc._lib.check(ctx.lib.gprc_ctx_set_interrupt(ctx.handle, cb, None))
    try:
        with pytest.raises(gprc._lib.GprcError, match="interrupted"):
            g.predict(Xs)
    finally:
        gprc._lib.check(ctx.lib.gprc_ctx_set_interrupt(ctx.handle, C.cast(None, cb_type), None))
    assert len(calls) == 1
    out = g.predict(Xs[:, :1000])         # the context and the model still work
    assert np.all(np.isfinite(out))


def test_two_host_threads_share_the_default_context(gprc, oracle):
    """ctypes releases the GIL around every library call, so two Python threads can be inside libgprc with the SAME
    context at once; the context serialises them (per-context mutex held by every entry point).  Each thread fits and
    predicts its own model repeatedly; every result must equal the single-threaded one bit for bit."""
    import threading
    rng = np.random.default_rng(77)
    jobs = []
    for t in range(2):
        n = 300 + 150 * t
        X = rng.uniform(-2, 2, (2, n))
        y = np.sin(X[0]) * np.cos(X[1]) + rng.normal(0, 0.1, n)
        Xs = rng.uniform(-2, 2, (2, 400))
        k = gprc.cov_func(gprc.sqrexp, l=0.8 + 0.3 * t)
        g = gprc.GPR(X, y, 0.05, k)
        jobs.append((X, y, Xs, k, g.predict(Xs), g.logp[0, 0]))
    errors = []

    def work(job):
        X, y, Xs, k, want, lp = job
        try:
            for _ in range(20):
                g = gprc.GPR(X, y, 0.05, k)
                if not (np.array_equal(g.predict(Xs), want) and g.logp[0, 0] == lp):
                    errors.append("result changed under concurrency")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(j,)) for j in jobs]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:3]
