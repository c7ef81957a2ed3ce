"""Variance-pass paths on a shard-sized predict (default n = 50 000, m = 125 000 = 6.6 waves of 148 tiles)."""
import ctypes as C, math, sys
import numpy as np
sys.path.insert(0, ".")
import gprc_b200 as g
ctx = g.default_context()
n, m, d = int(sys.argv[1]) if len(sys.argv) > 1 else 50000, int(sys.argv[2]) if len(sys.argv) > 2 else 125000, 8
rng = np.random.default_rng(4)
X = rng.uniform(-1, 1, (d, n)); y = np.sum(np.sin(math.pi * X), axis=0) + rng.normal(0, 0.1, n)
Xs = rng.uniform(-1, 1, (d, m))
dX, dy, dXs = ctx.upload(np.ascontiguousarray(X.T)), ctx.upload(y), ctx.upload(np.ascontiguousarray(Xs.T))
dm, dv = ctx.malloc(8 * m), ctx.malloc(8 * m)
kc, _ = g.KernelSpec("sqrexp", l=1.0).to_c()
h = C.c_void_p(); lp, info = C.c_double(0), C.c_long(0)
g._lib.check(ctx.lib.gprc_gpr_fit_dev(ctx.handle, kc, dX, d, n, dy, 0.01, C.byref(h), C.byref(lp), C.byref(info)))
ref = None
for path in [2, 3, 0, 0]:
    ctx.set_option(g._lib.OPT_PREDICT_PATH, path)
    ctx.reset_timers()
    g._lib.check(ctx.lib.gprc_gpr_predict_dev(h, dXs, m, dm, dv))
    tm, launches = ctx.timers()
    v = np.empty(m); ctx.d2h(v, dv)
    if ref is None: ref = v.copy()
    print("path %d: predict span %.1f ms (var %.1f) = %.2f TF/s  launches %d  identical %s" % (
        path, tm["predict"], tm["var"], n * n * m / tm["predict"] / 1e9, launches, np.array_equal(v, ref)), flush=True)
