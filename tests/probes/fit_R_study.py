"""tests/testthat/test-fit.R:12-17 of the reference, examined without R.

The reference asserts that fit(X, Y_i, 0.05, <all six families>)$cov names the family that generated Y_i.  The oracle
(and the CUDA path, which follows it) reproduces outcomes 1-3 and selects `polynomial` for Y4-Y6.  This script shows, with
an independent 40-digit evaluation (mpmath: its own Cholesky, nothing shared with NumPy / LAPACK / the oracle), that

  (a) the oracle's dens() (R/fit.R:117-124) is evaluated correctly at every family's optimum (agreement ~1e-13), and
  (b) for Y4-Y6 even the GLOBAL maximum of the generating family's log marginal likelihood over a dense parameter
      grid lies far below the polynomial family's score, so no optimiser -- R's included -- can make fit() return
      sqrexp / gammaexp / rationalquadratic there: the kernels have unit prior variance (no signal-variance parameter,
      R/GPRclass.R:394-402) and the targets have amplitude 5, a > 4-sigma event under every such prior, whereas
      (sigma + x y)^p scales freely.

    python tests/probes/fit_R_study.py            # prints the table of profiles/r2_test_fit_R_study.md
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gprc_oracle as o  # noqa: E402

NAMES = ["linear", "constant", "polynomial", "sqrexp", "gammaexp", "rationalquadratic"]
EXPECTED = ["linear", "constant", "polynomial", "sqrexp", "gammaexp", "rationalquadratic"]   # test-fit.R:12-17


def targets():
    X = np.arange(0, 1.1001, 0.1).reshape(1, -1)              # seq(0, 1.1, by = 0.1)       test-fit.R:2
    x = X[0]
    Ys = [3 * x, np.full(12, 5.0), 3 * x ** 2 - 2 * x, 5 * np.exp(-x ** 2), 5 * np.exp(-x ** 5), 5 / (1 + x ** 2)]  # :4-9
    return X, Ys


def mp_dens(x, y, noise, name, par, dps=40):
    """log p(y | X, theta) in 40-digit arithmetic, written from the formulas of R/GPRclass.R:381-403 and R/fit.R:121-123."""
    import mpmath as mp
    mp.mp.dps = dps
    n = len(x)
    xs = [mp.mpf(float(v)) for v in x]
    p = [mp.mpf(float(v)) for v in par]

    def k(a, b):
        r2 = (a - b) ** 2
        if name == "constant":
            return p[0]
        if name == "linear":
            return p[0] * a * b
        if name == "polynomial":
            return (a * b + p[0]) ** p[1]
        if name == "sqrexp":
            return mp.exp(-r2 / (2 * p[0] ** 2))
        if name == "gammaexp":                      # (l, gamma) positional, R/fit.R:118
            return mp.exp(-(mp.sqrt(r2) / p[0]) ** p[1]) if r2 != 0 else mp.mpf(1)
        if name == "rationalquadratic":             # (l, alpha)
            return (1 + r2 / (2 * p[1] * p[0] ** 2)) ** (-p[1])
        raise KeyError(name)

    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = k(xs[i], xs[j]) + (mp.mpf(noise) if i == j else 0)
    L = mp.cholesky(K)
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    z = mp.lu_solve(K, yv)
    return float(-(yv.T * z)[0] / 2 - sum(mp.log(L[i, i]) for i in range(n)) - mp.mpf(n) / 2 * mp.log(2 * mp.pi))


def grid_max(X, y, noise, name):
    """dense-grid maximum of dens over the family's parameter domain (the optimiser's box for the 1-parameter families,
    a wide box for the 2-parameter ones)"""
    best = (-math.inf, None)

    def ev(par):
        nonlocal best
        try:
            v = o.dens(X, y, noise, name, list(par), minors="cholesky")
        except o.OptimError:
            return
        if v > best[0]:
            best = (v, tuple(par))

    if name in ("constant", "linear", "sqrexp"):
        for a in np.concatenate([np.linspace(0.01, 10, 2000)]):          # Brent's interval [0, 10], R/fit.R:143
            ev([a])
    elif name == "polynomial":
        for p in range(1, 11):                                            # R/fit.R:148
            for s in np.linspace(0.0, 5.0, 501):
                ev([s, float(p)])
    else:
        for a in np.exp(np.linspace(math.log(0.02), math.log(200), 160)):      # l
            for b in np.exp(np.linspace(math.log(0.02), math.log(200), 160)):  # gamma / alpha
                if name == "gammaexp" and b > 2.0:
                    continue                                               # gamma > 2 is not a covariance function
                ev([a, b])
    return best


def study():
    X, Ys = targets()
    rows = []
    for t, y in enumerate(Ys):
        res = o.fit(X, y, 0.05, NAMES)
        scores = dict(zip(NAMES, res["score"]))
        gen = EXPECTED[t]
        gmax, gpar = grid_max(X, y, 0.05, gen)
        pmax, ppar = grid_max(X, y, 0.05, "polynomial")
        rows.append(dict(target="Y%d" % (t + 1), expected=gen, selected=res["cov"], scores=scores,
                         generating_family_global_max=gmax, generating_family_argmax=gpar,
                         polynomial_global_max=pmax, polynomial_argmax=ppar, par=res["par"]))
    return X, Ys, rows


def main():
    X, Ys, rows = study()
    print("| target | test-fit.R expects | oracle / CUDA path select | " + " | ".join(NAMES) +
          " | global max of the expected family (grid) | global max of polynomial (grid) |")
    print("|---|---|---|" + "---|" * (len(NAMES) + 2))
    for r in rows:
        print("| %s | %s | %s | " % (r["target"], r["expected"], r["selected"]) +
              " | ".join("%.4f" % r["scores"][k] for k in NAMES) +
              " | %.4f at %s | %.4f at %s |" % (r["generating_family_global_max"],
                                                 tuple(round(float(v), 3) for v in r["generating_family_argmax"]),
                                                 r["polynomial_global_max"],
                                                 tuple(round(float(v), 3) for v in r["polynomial_argmax"])))
    print()
    print("independent 40-digit check of dens() at the selected optimum and at the expected family's grid maximum:")
    for r, y in zip(rows, Ys):
        a = mp_dens(X[0], y, 0.05, r["selected"], r["par"])
        b = o.dens(X, y, 0.05, r["selected"], list(r["par"]), minors="cholesky")
        c = mp_dens(X[0], y, 0.05, r["expected"], r["generating_family_argmax"])
        print("  %s: %s%s  mpmath %.12f  oracle %.12f  (diff %.1e);  %s at its grid maximum: mpmath %.12f"
              % (r["target"], r["selected"], tuple(round(float(v), 6) for v in r["par"]), a, b, abs(a - b), r["expected"], c))


if __name__ == "__main__":
    main()
