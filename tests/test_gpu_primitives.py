"""Parity of the raw device primitives (DMMA GEMM core, blocked Cholesky, level-wise inversion, kernel-matrix build)
against NumPy/SciPy and the oracle.  All through the C ABI."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg

pytestmark = pytest.mark.gpu


def _dev(ctx, a):
    return ctx.upload(np.asfortranarray(a))


@pytest.mark.parametrize("transb", [1, 0])
@pytest.mark.parametrize("shape", [(128, 128, 16), (256, 384, 528), (512, 128, 1024)])
def test_dgemm(ctx, transb, shape):
    M, N, K = shape
    rng = np.random.default_rng(M + N + K + transb)
    A = rng.standard_normal((M, K))
    B = rng.standard_normal((N, K)) if transb else rng.standard_normal((K, N))
    Cm = rng.standard_normal((M, N))
    dA, dB, dC = _dev(ctx, A), _dev(ctx, B), _dev(ctx, Cm)
    ldb = N if transb else K
    rc = ctx.lib.gprc_dev_dgemm(ctx.handle, transb, M, N, K, -1.5, dA, M, dB, ldb, 0.5, dC, M)
    assert rc == 0, ctx.lib.gprc_last_error()
    out = np.empty((M, N), order="F")
    ctx.d2h(out, dC)
    ref = 0.5 * Cm - 1.5 * (A @ (B.T if transb else B))
    # same products, different summation order: a few ulp of the accumulated magnitude
    assert np.max(np.abs(out - ref)) <= 1e-13 * np.sqrt(K) * 16
    for p in (dA, dB, dC):
        ctx.free(p)


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n))
    return G @ G.T / n + np.eye(n)


@pytest.mark.parametrize("n", [128, 256, 640, 1152])
def test_potrf_and_trtri(ctx, n):
    A = _spd(n, n)
    dA = _dev(ctx, A)
    dinv = ctx.malloc(n * 128 * 8)
    info = C.c_long(-1)
    rc = ctx.lib.gprc_dev_potrf(ctx.handle, dA, n, n, dinv, C.byref(info))
    assert rc == 0, ctx.lib.gprc_last_error()
    assert info.value == 0
    out = np.empty((n, n), order="F")
    ctx.d2h(out, dA)
    L = np.tril(out)
    Lref = scipy.linalg.cholesky(A, lower=True)
    assert np.max(np.abs(L - Lref)) <= 1e-12 * np.max(np.abs(Lref))
    assert np.max(np.abs(L @ L.T - A)) <= 1e-13 * n
    # inversion: W L = I
    dW = ctx.malloc(n * n * 8)
    rc = ctx.lib.gprc_dev_trtri(ctx.handle, dA, n, n, dinv, dW, dA)
    assert rc == 0, ctx.lib.gprc_last_error()
    W = np.empty((n, n), order="F")
    ctx.d2h(W, dW)
    W = np.tril(W)
    assert np.max(np.abs(W @ L - np.eye(n))) <= 1e-11
    for p in (dA, dinv, dW):
        ctx.free(p)


def test_potrf_reports_first_bad_pivot(ctx):
    n = 384
    A = _spd(n, 7)
    A[200, 200] = -5.0  # leading minor 201 is the first that fails
    dA = _dev(ctx, A)
    dinv = ctx.malloc(n * 128 * 8)
    info = C.c_long(-1)
    assert ctx.lib.gprc_dev_potrf(ctx.handle, dA, n, n, dinv, C.byref(info)) == 0
    assert info.value == 201
    ctx.free(dA)
    ctx.free(dinv)


KERNELS = [
    ("constant", dict(c=1.7)),
    ("linear", dict(sigma=0.8)),
    ("polynomial", dict(sigma=0.25, p=3.0)),
    ("polynomial", dict(sigma=1.0, p=2.0)),
    ("sqrexp", dict(l=0.7)),
    ("gammaexp", dict(l=1.3, gamma=1.5)),
    ("rationalquadratic", dict(l=0.9, alpha=2.5)),
]


@pytest.mark.parametrize("name,params", KERNELS)
@pytest.mark.parametrize("D,nA,nB", [(1, 5, 3), (2, 70, 129), (8, 257, 64), (11, 33, 200)])
def test_cov_matrix_matches_oracle(gprc, oracle, ctx, name, params, D, nA, nB):
    _cov_matrix_case(gprc, oracle, ctx, name, params, D, nA, nB, gram=1)


@pytest.mark.parametrize("name,params", [k for k in KERNELS if k[0] in ("linear", "polynomial", "sqrexp", "rationalquadratic")])
@pytest.mark.parametrize("D,nA,nB", [(4, 70, 129), (8, 257, 64), (16, 130, 200), (32, 64, 65)])
def test_cov_matrix_tensor_core_build(gprc, oracle, ctx, name, params, D, nA, nB):
    # GPRC_OPT_GRAM_DMMA = 2: the DMMA Gram-tile build wherever it is eligible
    _cov_matrix_case(gprc, oracle, ctx, name, params, D, nA, nB, gram=2)


def _cov_matrix_case(gprc, oracle, ctx, name, params, D, nA, nB, gram):
    ctx.set_option(gprc._lib.OPT_GRAM_DMMA, gram)
    try:
        _cov_matrix_check(gprc, oracle, ctx, name, params, D, nA, nB, gram)
    finally:
        ctx.set_option(gprc._lib.OPT_GRAM_DMMA, 1)


def _cov_matrix_check(gprc, oracle, ctx, name, params, D, nA, nB, gram):
    rng = np.random.default_rng(D * 1000 + nA + nB)
    A = rng.uniform(-1, 1, (D, nA))
    B = rng.uniform(-1, 1, (D, nB))
    B[:, 0] = A[:, 0]  # a coincident pair
    got = gprc.covariance_matrix(A, B, gprc.cov_func(getattr(gprc, name), **params), ctx=ctx)
    ref = oracle.covariance_matrix(A, B, oracle.cov_func(getattr(oracle, name), **params))
    assert got.shape == (nA, nB)
    # distance kernels: identical difference/sum order, exp/pow differ by <= 2 ulp between CUDA and glibc.
    # dot-product kernels: NumPy's reduction order differs, so a dot near zero only agrees to a few ulp of
    # sum_d |x_d y_d| (<= D here), which the polynomial epilogue scales by p (sigma + D)^(p-1)
    atol = 0.0 if name in ("sqrexp", "gammaexp", "rationalquadratic", "constant") else 4e-16 * D * 64
    rtol = 1e-14
    if gram == 2 and D % 4 == 0 and name in ("sqrexp", "rationalquadratic"):
        # tensor-core build: r2 = |a|^2 + |b|^2 - 2 a.b carries ~1e-16 (|a|^2 + |b|^2) of cancellation error,
        # i.e. ~1e-15..1e-14 relative on the entry (exactly 1 at coincident points is NOT guaranteed off the diagonal)
        rtol = 2e-13
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol)


@pytest.mark.parametrize("name,params", [k for k in KERNELS if k[0] != "constant"])
def test_cov_matrix_direct_build_is_faithful_to_1e14(gprc, oracle, ctx, name, params):
    # GPRC_OPT_GRAM_DMMA = 0: every kernel goes through direct differences in the reference's operation order
    rng = np.random.default_rng(77)
    A, B = rng.uniform(-1, 1, (8, 200)), rng.uniform(-1, 1, (8, 130))
    B[:, 3] = A[:, 5]
    ctx.set_option(gprc._lib.OPT_GRAM_DMMA, 0)
    try:
        got = gprc.covariance_matrix(A, B, gprc.cov_func(getattr(gprc, name), **params), ctx=ctx)
    finally:
        ctx.set_option(gprc._lib.OPT_GRAM_DMMA, 1)
    ref = oracle.covariance_matrix(A, B, oracle.cov_func(getattr(oracle, name), **params))
    atol = 0.0 if name in ("sqrexp", "gammaexp", "rationalquadratic") else 4e-16 * 8 * 64
    np.testing.assert_allclose(got, ref, rtol=1e-14, atol=atol)
    if name in ("sqrexp", "gammaexp", "rationalquadratic"):
        assert got[5, 3] == 1.0  # coincident points: exactly k(x, x)


def test_cov_matrix_linear_vector_sigma(gprc, oracle, ctx):
    rng = np.random.default_rng(5)
    A, B = rng.standard_normal((3, 40)), rng.standard_normal((3, 17))
    sig = np.array([0.5, 2.0, 1.25])
    got = gprc.covariance_matrix(A, B, gprc.cov_func(gprc.linear, sigma=sig), ctx=ctx)
    ref = oracle.covariance_matrix(A, B, oracle.cov_func(oracle.linear, sigma=sig))
    np.testing.assert_allclose(got, ref, rtol=1e-14)


def test_cov_matrix_closure_goes_through_host(gprc, oracle):
    # an opaque closure (test-gpc.R:7) is evaluated on the host exactly like the reference's outer()
    kappa = lambda x, y: np.exp(-3 * (x - y) ** 2)[0]
    X = np.arange(-1, 1.0001, 0.1).reshape(1, -1)
    np.testing.assert_array_equal(gprc.covariance_matrix(X, X, kappa), oracle.covariance_matrix(X, X, kappa))


def test_context_trim_releases_the_cached_blocks(gprc):
    """gprc_ctx_trim: blocks freed by the library stay in the context's pool until trimmed; afterwards the driver reports
    them free again and the context keeps working."""
    import torch
    ctx = gprc.Context(0)
    rng = np.random.default_rng(5)
    X = rng.uniform(-1, 1, (3, 3000))
    y = rng.normal(size=3000)
    g = gprc.GPR(X, y, 0.1, gprc.cov_func(gprc.sqrexp, l=1.0), ctx=ctx)
    g.predict(X[:, :100])
    del g
    free0 = torch.cuda.mem_get_info(0)[0]
    released = ctx.trim()
    assert released >= 8 * 3072 * 3072                    # at least the factor
    assert torch.cuda.mem_get_info(0)[0] >= free0 + released // 2
    assert ctx.trim() == 0
    g = gprc.GPR(X, y, 0.1, gprc.cov_func(gprc.sqrexp, l=1.0), ctx=ctx)
    assert np.isfinite(g.logp[0, 0])
    del g
    ctx.close()
