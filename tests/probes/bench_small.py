"""Timings of the secondary BASELINE configs (C1 GPR vignette scale, C2 GPC n=2000, C3 fit objective n=5000) next to
the oracle on the host cores.  Informational (parity cases, not bench lines)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gprc_b200 as g
from oracle import gprc_oracle as o

ctx = g.default_context()


def timed(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        ctx.sync()
        t = time.perf_counter()
        r = fn()
        ctx.sync()
        best = min(best, time.perf_counter() - t)
    return best, r


what = sys.argv[1:] or ["C1", "C2", "C3"]
if "C1" in what:
    c = o.make_config("C1")
    tg, pg = timed(lambda: g.GPR(c["X"], c["y"], c["noise"], g.cov_func(g.sqrexp, l=1.0)).predict(c["Xs"]))
    t0 = time.perf_counter()
    po = o.GPR(c["X"], c["y"], c["noise"], o.cov_func(o.sqrexp, l=1.0)).predict(c["Xs"])
    to = time.perf_counter() - t0
    print("C1 GPR n=200 m=1000: gpu %.2f ms  oracle %.2f ms  max|dmean| %.2e max|dvar| %.2e" % (
        tg * 1e3, to * 1e3, np.max(np.abs(pg[:, 0] - po[:, 0])), np.max(np.abs(pg[:, 1] - po[:, 1]))), flush=True)
if "C2" in what:
    c = o.make_config("C2")
    ctx.reset_timers()
    t0 = time.perf_counter()
    gc = g.GPC(c["X"], c["y"], g.cov_func(g.sqrexp, l=0.2), verbose=False)
    t1 = time.perf_counter()
    fs, V = gc.predict_latent(c["Xs"])
    t2 = time.perf_counter()
    pr = gc.predict_class(c["Xs"])
    t3 = time.perf_counter()
    tm, launches = ctx.timers()
    print("C2 GPC n=2000 m=%d: gpu fit %.1f ms (%d it) latent %.1f ms quadrature(host) %.1f ms; launches %d timers %s" % (
        c["Xs"].shape[1], (t1 - t0) * 1e3, gc.iterations, (t2 - t1) * 1e3, (t3 - t2) * 1e3, launches,
        {k: round(v, 2) for k, v in tm.items() if v}), flush=True)
    t0 = time.perf_counter()
    oc = o.GPC(c["X"], c["y"], o.cov_func(o.sqrexp, l=0.2))
    t1 = time.perf_counter()
    ofs, oV = oc.predict_latent(c["Xs"])
    t2 = time.perf_counter()
    print("   oracle fit %.1f ms (%d it) latent %.1f ms; iterations equal %s, max|dfs| %.2e max|dV| %.2e labels equal %s" % (
        (t1 - t0) * 1e3, oc.iterations, (t2 - t1) * 1e3, gc.iterations == oc.iterations, np.max(np.abs(fs - ofs)),
        np.max(np.abs(V - oV)), np.array_equal(pr >= 0.5, oc.predict_class(c["Xs"]) >= 0.5)), flush=True)
if "C3" in what:
    c = o.make_config("C3")
    obj = g.Objective(c["X"], c["y"], c["noise"], minors="cholesky")
    tg, lp = timed(lambda: obj.dens_batch("rationalquadratic", c["starts"]))
    t0 = time.perf_counter()
    ref = o.dens(c["X"], c["y"], c["noise"], "rationalquadratic", list(c["starts"][0]), minors="cholesky")
    to = time.perf_counter() - t0
    print("C3 dens rationalquadratic n=5000 d=4: gpu %.1f ms per evaluation (16 starts in %.1f ms); oracle %.1f ms per "
          "evaluation; rel diff %.2e" % (tg / 16 * 1e3, tg * 1e3, to * 1e3, abs(lp[0][0] - ref) / abs(ref)), flush=True)
    tg, gr = timed(lambda: obj.dens_deriv("rationalquadratic", c["starts"][0], formula=1), reps=2)
    print("   textbook gradient n=5000: gpu %.1f ms %s" % (tg * 1e3, gr), flush=True)
