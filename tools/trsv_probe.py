"""alpha + logp (two triangular sweeps, R/GPRclass.R:152-153) with the cooperative sweeps (GPRC_OPT_TRSV = 0) and the
dataflow kernel (1): the library's own `solve` phase timer (CUDA events), best of 3 fits per size."""
import sys

import numpy as np

sys.path.insert(0, ".")
import gprc_b200 as g

ctx = g.default_context()
sizes = [int(a) for a in sys.argv[1:]] or [2048, 5000, 16384, 50000]
for n in sizes:
    rng = np.random.default_rng(n)
    X = rng.uniform(-1, 1, (8, n))
    y = np.sum(np.sin(3 * X), axis=0) + rng.normal(0, 0.1, n)
    res = {}
    for mode in (0, 1):
        ctx.set_option(g._lib.OPT_TRSV, mode)
        best, alpha = 1e9, None
        for rep in range(3):
            ctx.reset_timers()
            m = g.GPR(X, y, 0.01, g.cov_func(g.sqrexp, l=1.0), ctx=ctx)
            t, _ = ctx.timers()
            best = min(best, t["solve"])
            alpha = m.alpha.copy()
            chol = t["chol"]
            del m
        res[mode] = (best, alpha, chol)
    ctx.set_option(g._lib.OPT_TRSV, g._lib.TRSV_DEFAULT)
    d = float(np.max(np.abs(res[0][1] - res[1][1])) / np.max(np.abs(res[0][1])))
    floor_ms = 8.0 * n * n / 6553e9 * 1e3
    print("n=%6d  solve phase: cooperative sweeps %.3f ms, dataflow kernel %.3f ms (HBM floor %.3f ms: L read twice); "
          "alpha rel diff %.1e; Cholesky %.2f ms" % (n, res[0][0], res[1][0], floor_ms, d, res[1][2]), flush=True)
