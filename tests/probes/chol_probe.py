"""Cholesky phase (library timer, CUDA events) with the launch-sequence factorisation (GPRC_OPT_CHOL_TILES = 0) and the
persistent tile kernel (128): best of 3 fits per size; and the GPC Newton fit of config 2."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gprc_b200 as g
from oracle import gprc_oracle as o

ctx = g.default_context()
sizes = [int(a) for a in sys.argv[1:]] or [2048, 5000, 8192, 16384]
for n in sizes:
    rng = np.random.default_rng(n)
    X = rng.uniform(-1, 1, (8, n))
    y = np.sum(np.sin(3 * X), axis=0) + rng.normal(0, 0.1, n)
    res = {}
    for tiles in (0, 128):
        ctx.set_option(g._lib.OPT_CHOL_TILES, tiles)
        best, lp = 1e9, None
        for rep in range(3):
            ctx.reset_timers()
            m = g.GPR(X, y, 0.01, g.cov_func(g.sqrexp, l=1.0), ctx=ctx)
            t, _ = ctx.timers()
            best = min(best, t["chol"])
            lp = m.logp[0, 0]
            del m
        res[tiles] = (best, lp)
    ctx.set_option(g._lib.OPT_CHOL_TILES, g._lib.CHOL_TILES_DEFAULT)
    tf = lambda ms: n ** 3 / 3 / (ms * 1e-3) / 1e12
    print("n=%6d  Cholesky: launch sequence %.3f ms (%.1f TFLOP/s), persistent tile kernel %.3f ms (%.1f TFLOP/s); "
          "logp rel diff %.1e" % (n, res[0][0], tf(res[0][0]), res[128][0], tf(res[128][0]),
                                  abs(res[0][1] - res[128][1]) / abs(res[0][1])), flush=True)
c = o.make_config("C2")
for tiles in (0, 128):
    ctx.set_option(g._lib.OPT_CHOL_TILES, tiles)
    best = 1e9
    for rep in range(3):
        ctx.sync()
        t0 = time.perf_counter()
        gc = g.GPC(c["X"], c["y"], g.cov_func(g.sqrexp, l=0.2), verbose=False)
        ctx.sync()
        best = min(best, time.perf_counter() - t0)
    print("C2 GPC n=2000 fit (%d Newton iterations), chol tiles %d: %.2f ms" % (gc.iterations, tiles, best * 1e3), flush=True)
ctx.set_option(g._lib.OPT_CHOL_TILES, g._lib.CHOL_TILES_DEFAULT)
