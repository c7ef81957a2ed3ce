"""Generates tests/golden/*.npz from the oracle (oracle/gprc_oracle.py).

The reference is an R package and R is not installed in the build image, so the reference itself cannot produce
vectors here.  What is pinned by the reference are the closed forms of tests/testthat/test-gpr.R (stored under
"known_*"); everything else in these files is a regression snapshot of the restatement (PARITY UNPINNED, see the
oracle's header) so that the GPU path and the oracle are both checked against committed numbers.

    python tests/golden/make_golden.py
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gprc_oracle as o  # noqa: E402


def main():
    e = math.exp
    known = dict(
        known_poly=np.array([2.0, 1 / 8]),                                   # test-gpr.R:6-9
        known_const1=np.array([4 / 3, 1 / 3]),                               # :12-15
        known_const2=np.array([5 / 3, 1 / 3]),                               # :16-19
        known_sqrexp=np.array([(2 * e(-2) - e(-1)) / (4 - e(-1)),            # :23-27
                               1 - (2 * e(-1) - 2 * e(-3) + 2 * e(-4)) / (4 - e(-1))]),
    )
    # GPR: BASELINE config 1 at reduced size, three kernels
    out = dict(known)
    for tag, name, params, D, n, m, noise in [
        ("c1", "sqrexp", dict(l=1.0), 1, 120, 64, 0.01),
        ("c4s", "sqrexp", dict(l=1.0), 8, 300, 50, 0.01),
        ("rq", "rationalquadratic", dict(l=1.0, alpha=1.0), 4, 200, 40, 0.05),
        ("gx", "gammaexp", dict(l=1.0, gamma=1.5), 8, 150, 30, 0.1),
        ("poly", "polynomial", dict(sigma=1.0, p=3.0), 8, 100, 30, 0.1),
    ]:
        rng = np.random.default_rng(1000 + n + m)
        lim = 6.0 if D == 1 else 1.0
        X = rng.uniform(-lim, lim, (D, n))
        y = np.sum(np.sin(X), axis=0) + rng.normal(0, 0.1, n)
        Xs = rng.uniform(-lim, lim, (D, m))
        g = o.GPR(X, y, noise, o.cov_func(getattr(o, name), **params))
        out.update({tag + "_X": X, tag + "_y": y, tag + "_Xs": Xs, tag + "_noise": np.array(noise),
                    tag + "_pred": g.predict(Xs), tag + "_logp": np.array(g.logp), tag + "_alpha": g.alpha,
                    tag + "_kernel": np.array(name), tag + "_params": np.array(json.dumps(params))})
    # GPC: BASELINE config 2 at reduced size
    cfg = o.make_config("C2", n=250, m=400)
    c = o.GPC(cfg["X"], cfg["y"], o.cov_func(o.sqrexp, l=0.5))
    fs, V = c.predict_latent(cfg["Xs"])
    out.update(gpc_X=cfg["X"], gpc_y=cfg["y"], gpc_Xs=cfg["Xs"], gpc_l=np.array(0.5), gpc_f_hat=c.f_hat,
               gpc_logq=np.array(c.logq), gpc_iter=np.array(c.iterations), gpc_trace=np.array(c.objective_trace),
               gpc_fs_bar=fs, gpc_Vfs=V, gpc_prob=c.predict_class(cfg["Xs"]))
    # dens on a theta grid (config 3 shape at n = 120)
    c3 = o.make_config("C3", n=120)
    thetas = c3["starts"][:6]
    out.update(dens_X=c3["X"], dens_y=c3["y"], dens_thetas=thetas,
               dens_rq=np.array([o.dens(c3["X"], c3["y"], 0.05, "rationalquadratic", list(t), minors="cholesky")
                                 for t in thetas]))
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
