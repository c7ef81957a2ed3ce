"""The INT8 tensor-core variance pass (predict path 4, csrc/ozaki.cuh) against the ORACLE -- not against its FP64
sibling -- on inputs chosen to hurt a fixed-point product: several drain rounds, tiny noise, long length scales, a
polynomial kernel on |x| <= 3 (k** spans 6 decades across test points), rows of L whose entries span > 10 decades.
Replaces R/GPRclass.R:160-164 (v = solve(L, K_star); var = k** - colSums(v * v)).

Kernel variants (GPRC_OPT_INT8_TILE): 2 = stacked digit planes on cluster pairs with V multicast (default),
1 = stacked planes, 64 = one MMA per digit pair (round 1), 128 = wide tiles in two order passes (round 1).

A-priori bound of one update  R_i -= L[i, <i] V[<i]  with S digits: every entry is rounded to a
fixed-point grid 2^(e - 8S + 2) below its row (L) / test-point (V) bound 2^e, and digit pairs of order a + b >= S are
dropped, so per k-term  |error| <= S * 2^(-8S + 4) * rowmax_i * sqrt(k**_t)  and per entry of R at most K times that.
`update_bound` below evaluates it; the tests assert the observed variance error against the propagated bound
2 sqrt(k**) ||L^-1||_2 sqrt(n) * update_bound  as well as against the north_star tolerance 1e-9 max(|var|, k**)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

VARIANTS = [2, 1, 64, 128]


def predict_int8(gprc, ctx, model, Xs, variant, digits=7):
    ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 4)
    ctx.set_option(gprc._lib.OPT_OZAKI_DIGITS, digits)
    ctx.set_option(gprc._lib.OPT_INT8_TILE, variant)
    try:
        out = model.predict(Xs)
        assert ctx.last_predict_path() == 4
    finally:
        ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 0)
        ctx.set_option(gprc._lib.OPT_OZAKI_DIGITS, 7)
        ctx.set_option(gprc._lib.OPT_INT8_TILE, gprc._lib.INT8_TILE_DEFAULT)
    return out


def predict_fp64(gprc, ctx, model, Xs):
    ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 2)
    try:
        return model.predict(Xs)
    finally:
        ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 0)


def update_bound(S, K, rowmax, sqrt_kss):
    return K * S * 2.0 ** (-8 * S + 4) * rowmax * sqrt_kss


def assert_var_close(got, ref, kss, tol=1e-9):
    err = np.abs(got[:, 1] - ref[:, 1]) / np.maximum(np.abs(ref[:, 1]), kss)
    assert np.max(err) <= tol, "variance: max relative error %.3e" % np.max(err)
    scale = np.max(np.abs(ref[:, 0]))
    assert np.max(np.abs(got[:, 0] - ref[:, 0])) <= 1e-9 * scale


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n", [1408, 1536])  # 11 block rows (odd: the last one has no partner) and 12
def test_variants_against_the_oracle(gprc, oracle, ctx, variant, n):
    rng = np.random.default_rng(51)
    m, D = 148 * 64 + 77, 5
    X = rng.uniform(-1, 1, (D, n))
    y = np.sum(np.cos(2 * X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-1, 1, (D, m))
    g = gprc.GPR(X, y, 0.05, gprc.cov_func(gprc.rationalquadratic, l=0.8, alpha=1.5), ctx=ctx)
    ok = oracle.cov_func(oracle.rationalquadratic, l=0.8, alpha=1.5)
    ref = oracle.GPR(X, y, 0.05, ok).predict(Xs)
    got = predict_int8(gprc, ctx, g, Xs, variant)
    assert_var_close(got, ref, ok(Xs, Xs))


def test_stacked_kernel_is_bitwise_the_single_product_kernel(gprc, ctx):
    """Variant 1 computes the same integer sums as variant 64 (one MMA covers several digit pairs) and drains them the
    same way: identical bits.  Variant 2 splits block row 2j+1's product at the pair boundary (INT8 below, FP64 inside
    the pair): not bitwise, but within 1e-13 k**."""
    rng = np.random.default_rng(52)
    n, m, D = 1536, 148 * 64 + 333, 6
    X = rng.uniform(-1, 1, (D, n))
    y = np.sum(X ** 2, axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-1.5, 1.5, (D, m))
    g = gprc.GPR(X, y, 0.1, gprc.cov_func(gprc.polynomial, sigma=1.0, p=3.0), ctx=ctx)
    a = predict_int8(gprc, ctx, g, Xs, 64)
    b = predict_int8(gprc, ctx, g, Xs, 1)
    c = predict_int8(gprc, ctx, g, Xs, 2)
    np.testing.assert_array_equal(a, b)
    kss = (np.sum(Xs * Xs, axis=0) + 1.0) ** 3
    assert np.max(np.abs(a[:, 1] - c[:, 1]) / kss) < 1e-13
    np.testing.assert_array_equal(a[:, 0], c[:, 0])


@pytest.mark.parametrize("variant,digits", [(2, 7), (1, 7), (2, 8), (1, 8), (2, 6), (64, 8), (128, 8), (128, 7)])
def test_several_drain_rounds_against_the_oracle(gprc, oracle, ctx, variant, digits):
    """n = 20 480 (160 block rows, 80 pairs): the int32 accumulators are drained every 16 384 k-values (8192 with 8
    digits), so the last block rows / pairs run 2 (3) drain rounds (K up to 20 352).  Oracle = SciPy dpotrf / dtrtrs at
    the same n."""
    rng = np.random.default_rng(38)
    n, m, D = 20480, 200, 4
    X = rng.uniform(-1, 1, (D, n))
    y = np.sum(np.sin(2 * X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-1, 1, (D, m))
    g = gprc.GPR(X, y, 0.05, gprc.cov_func(gprc.rationalquadratic, l=0.7, alpha=2.0), ctx=ctx)
    ref = _oracle_big(oracle, X, y, Xs)
    got = predict_int8(gprc, ctx, g, Xs, variant, digits)
    assert_var_close(got, ref, np.ones(m), tol=1e-9 if digits >= 7 else 1e-8)
    assert abs(g.logp[0, 0] - _ORACLE_CACHE["logp"]) <= 1e-8 * abs(_ORACLE_CACHE["logp"])


_ORACLE_CACHE = {}


def _oracle_big(oracle, X, y, Xs):
    if "pred" not in _ORACLE_CACHE:
        ok = oracle.cov_func(oracle.rationalquadratic, l=0.7, alpha=2.0)
        o = oracle.GPR(X, y, 0.05, ok)
        _ORACLE_CACHE["pred"] = o.predict(Xs)
        _ORACLE_CACHE["logp"] = float(o.logp)
    return _ORACLE_CACHE["pred"]


ADVERSARIAL = [
    # name, kernel, params, noise, x-range train, x-range test, D
    ("tiny_noise", "sqrexp", dict(l=1.0), 1e-6, 1.0, 1.0, 3),          # cond(K + noise I) ~ n / 1e-6
    ("long_length_scale", "sqrexp", dict(l=3.0), 1e-4, 1.0, 1.0, 4),  # K nearly rank one
    ("polynomial_wide", "polynomial", dict(sigma=1.0, p=3.0), 0.1, 3.0, 3.0, 3),   # k** from 1 to 2e4 per test point
    ("gammaexp_rough", "gammaexp", dict(l=0.3, gamma=1.0), 0.01, 1.0, 1.0, 2),
]


@pytest.mark.parametrize("variant", [2, 1])
@pytest.mark.parametrize("case", ADVERSARIAL, ids=[c[0] for c in ADVERSARIAL])
def test_adversarial_inputs_against_the_oracle(gprc, oracle, ctx, variant, case):
    name, kname, params, noise, rx, rxs, D = case
    rng = np.random.default_rng(53)
    n, m = 1920, 148 * 64 + 11
    X = rng.uniform(-rx, rx, (D, n))
    y = np.sum(np.sin(X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-rxs, rxs, (D, m))
    g = gprc.GPR(X, y, noise, gprc.cov_func(getattr(gprc, kname), **params), ctx=ctx)
    ok = oracle.cov_func(getattr(oracle, kname), **params)
    o = oracle.GPR(X, y, noise, ok)
    ref = o.predict(Xs)
    kss = ok(Xs, Xs)
    got = predict_int8(gprc, ctx, g, Xs, variant)
    fp64 = predict_fp64(gprc, ctx, g, Xs)
    # what two correct FP64 implementations differ by on this input (cancellation in k** - v'v, conditioning of L):
    # the INT8 pass must not be worse than 4 x that, nor than the north_star tolerance where FP64 itself meets it
    e_int8 = np.max(np.abs(got[:, 1] - ref[:, 1]) / np.maximum(np.abs(ref[:, 1]), kss))
    e_fp64 = np.max(np.abs(fp64[:, 1] - ref[:, 1]) / np.maximum(np.abs(ref[:, 1]), kss))
    assert e_int8 <= max(1e-9, 4 * e_fp64), (name, e_int8, e_fp64)
    if e_fp64 <= 2.5e-10:
        assert e_int8 <= 1e-9, (name, e_int8, e_fp64)
    # a-priori bound of the digit arithmetic, propagated through the solve
    L = np.asarray(o.L)
    rowmax = np.max(np.abs(np.tril(L, -1)), axis=1)
    linv_norm = 1.0 / np.sqrt(max(np.linalg.eigvalsh(L @ L.T)[0], 1e-300))
    bound = 2 * np.sqrt(kss) * linv_norm * np.sqrt(n) * update_bound(7, n, np.max(rowmax), np.sqrt(kss))
    assert np.all(np.abs(got[:, 1] - fp64[:, 1]) <= bound + 1e-16 * kss), name


@pytest.mark.parametrize("variant", [2, 1])
def test_rows_of_L_with_ten_decades_of_dynamic_range(gprc, oracle, ctx, variant):
    """Precomputed K = D A D with D spanning 1e-6 .. 1e5: the rows of L inherit the scaling, entries below 2^-54 of
    their row maximum vanish from the digit planes.  Goes through the closure / precomputed entry points with the
    path forced to 4."""
    rng = np.random.default_rng(54)
    n, m = 1280, 148 * 64
    B = rng.normal(size=(n, 40))
    A = B @ B.T / 40 + 0.5 * np.eye(n)
    d = 10.0 ** rng.uniform(-6, 5, n)
    K = (A * d[:, None]) * d[None, :]
    y = rng.normal(size=n) * d
    W = rng.normal(size=(n, m)) * 0.1
    Ks = K @ W                                     # test covariances in the span of K: a valid GP
    kss = np.einsum("ij,ij->j", W, Ks) * 1.5 + 1e-12
    import scipy.linalg as sl
    L = sl.cholesky(K, lower=True)
    v = sl.solve_triangular(L, Ks, lower=True)
    ref_var = kss - np.sum(v * v, axis=0)

    def table(x, y):
        """closure kernel over point INDICES: training points 0 .. n-1, test points n .. n+m-1"""
        i, j = np.asarray(x[0], dtype=np.int64), np.asarray(y[0], dtype=np.int64)
        out = np.empty(i.shape)
        tt, ts, st, ss = (i < n) & (j < n), (i < n) & (j >= n), (i >= n) & (j < n), (i >= n) & (j >= n)
        out[tt] = K[i[tt], j[tt]]
        out[ts] = Ks[i[ts], j[ts] - n]
        out[st] = Ks[j[st], i[st] - n]
        out[ss] = kss[i[ss] - n]      # only the pointwise prior variances k(x*, x*) are ever asked for
        return out

    g = gprc.GPR(np.arange(n, dtype=float)[None, :], y, 0.0, table, ctx=ctx)
    Xs = (n + np.arange(m, dtype=float))[None, :]
    out = predict_int8(gprc, ctx, g, Xs, variant)
    out2 = predict_fp64(gprc, ctx, g, Xs)
    assert ctx.last_predict_path() == 2
    e_int8 = np.max(np.abs(out[:, 1] - ref_var) / kss)
    e_fp64 = np.max(np.abs(out2[:, 1] - ref_var) / kss)
    assert e_int8 <= max(1e-9, 4 * e_fp64), (e_int8, e_fp64)
