"""The per-test-point quadrature of GPC$predict_class (R/GPCclass.R:116-117).  It stays on the host in this round
(SURVEY.md section 8f item 1): R's integrate() is QUADPACK dqagi, and so is scipy.integrate.quad on an infinite range;
tolerances are R's defaults (rel.tol = abs.tol = .Machine$double.eps^0.25, 100 subdivisions)."""
from __future__ import annotations

import math
import warnings

import numpy as np
import scipy.integrate

_TOL = float(np.finfo(float).eps ** 0.25)
_SQRT2PI = math.sqrt(2.0 * math.pi)


def logistic_gaussian_integral(mean, sd):
    if not (sd > 0):
        # dnorm(sd < 0) is NaN -> integrate() stops with "non-finite function value"
        raise FloatingPointError("non-finite function value")

    def integrand(z):
        try:
            s = 1.0 / (1.0 + math.exp(-z))
        except OverflowError:
            s = 0.0
        return s * math.exp(-0.5 * ((z - mean) / sd) ** 2) / (sd * _SQRT2PI)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        val = scipy.integrate.quad(integrand, -np.inf, np.inf, epsabs=_TOL, epsrel=_TOL, limit=100, full_output=1)[0]
    return val
