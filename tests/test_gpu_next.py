"""GPU tests that have NOT yet run on a B200 (written after the round's GPU budget was spent): the non-gating `gpu_next`
tier.  `python -m pytest tests -m gpu_next` on a GPU box; move each green test into a `gpu`-marked file."""
import numpy as np
import pytest


@pytest.mark.gpu_next
@pytest.mark.parametrize("digits,tile", [(8, 64), (8, 128), (7, 128)])
def test_int8_substitution_with_several_drain_rounds(gprc, ctx, digits, tile):
    """n = 16 640 (130 block rows): with 8 digits the accumulators are drained every 8192 k-values, so the last block rows
    run 3 drain intervals (6 (chunk, pass) rounds of the wide kernel); with 7 digits the interval is 16 384 (2 intervals).
    tools/oz_test checked single launches of this shape against the digit emulation; this is the whole pass."""
    rng = np.random.default_rng(38)
    n, m, D = 16640, 200, 4
    X = rng.uniform(-1, 1, (D, n))
    y = np.sum(np.sin(2 * X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-1, 1, (D, m))
    g = gprc.GPR(X, y, 0.05, gprc.cov_func(gprc.rationalquadratic, l=0.7, alpha=2.0), ctx=ctx)
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
            ctx.set_option(gprc._lib.OPT_INT8_TILE, 64)
    np.testing.assert_allclose(out[4][:, 0], out[2][:, 0], rtol=1e-13, atol=0)
    assert np.max(np.abs(out[4][:, 1] - out[2][:, 1])) < 1e-11


@pytest.mark.gpu_next
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
