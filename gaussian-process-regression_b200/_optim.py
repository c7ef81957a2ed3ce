"""Host-side optimisers that drive fit(): restatements of R's ``optimize``/``optim(method = "Brent")`` engine
(src/appl/fmin.c, Brent_fmin) and of ``optim(method = "BFGS")`` (src/appl/optim.c, vmmin), which R/fit.R:143-160
calls.  They stay on the host: every objective / gradient evaluation they request is
one call into libgprc.

Provenance / licence: see the header of csrc/optim.hpp (Brent 1973 / Nash 1990 algorithms; R's GPL-2+ C sources were read
to reproduce constants, acceptance rules and counters exactly; independent restatement, no code copied)."""
from __future__ import annotations

import math

import numpy as np

EPS = np.finfo(float).eps


class OptimError(RuntimeError):
    pass


def brent_fmin(f, ax, bx, tol):
    """R src/appl/fmin.c Brent_fmin (the engine of optimize() and optim(method = "Brent"))."""
    c = (3.0 - math.sqrt(5.0)) * 0.5
    eps = math.sqrt(EPS)
    a, b = ax, bx
    v = a + c * (b - a)
    w = x = v
    d = e = 0.0
    fx = f(x)
    fv = fw = fx
    tol3 = tol / 3.0
    while True:
        xm = (a + b) * 0.5
        tol1 = eps * abs(x) + tol3
        t2 = tol1 * 2.0
        if abs(x - xm) <= t2 - (b - a) * 0.5:
            break
        p = q = r = 0.0
        if abs(e) > tol1:
            r = (x - w) * (fx - fv)
            q = (x - v) * (fx - fw)
            p = (x - v) * q - (x - w) * r
            q = (q - r) * 2.0
            if q > 0.0:
                p = -p
            else:
                q = -q
            r = e
            e = d
        if abs(p) >= abs(q * 0.5 * r) or p <= q * (a - x) or p >= q * (b - x):
            e = (b - x) if x < xm else (a - x)
            d = c * e
        else:
            d = p / q
            u = x + d
            if u - a < t2 or b - u < t2:
                d = tol1
                if x >= xm:
                    d = -d
        if abs(d) >= tol1:
            u = x + d
        elif d > 0.0:
            u = x + tol1
        else:
            u = x - tol1
        fu = f(u)
        if fu <= fx:
            if u < x:
                b = x
            else:
                a = x
            v, w, x = w, x, u
            fv, fw, fx = fw, fx, fu
        else:
            if u < x:
                a = u
            else:
                b = u
            if fu <= fw or w == x:
                v, fv, w, fw = w, fw, u, fu
            elif fu <= fv or v == x or v == w:
                v, fv = u, fu
    return x


def vmmin(b0, fminfn, fmingr, maxit=100, abstol=-math.inf, reltol=math.sqrt(EPS)):
    """R src/appl/optim.c vmmin (optim(method = "BFGS")).  Returns (par, value, fail)."""
    stepredn, acctol, reltest = 0.2, 0.0001, 10.0
    b = np.array(b0, dtype=float)
    n = len(b)
    B = np.zeros((n, n))
    f = fminfn(b)
    if not math.isfinite(f):
        raise OptimError("initial value in 'vmmin' is not finite")
    Fmin = f
    funcount = gradcount = 1
    g = np.array(fmingr(b), dtype=float)
    it = 1
    ilast = gradcount
    t = np.zeros(n)
    X = np.zeros(n)
    c = np.zeros(n)
    while True:
        if ilast == gradcount:
            B[:] = 0.0
            for i in range(n):
                B[i, i] = 1.0
        X[:] = b
        c[:] = g
        gradproj = 0.0
        for i in range(n):
            s = 0.0
            for j in range(i + 1):
                s -= B[i, j] * g[j]
            for j in range(i + 1, n):
                s -= B[j, i] * g[j]
            t[i] = s
            gradproj += s * g[i]
        if gradproj < 0.0:
            steplength = 1.0
            accpoint = False
            while True:
                count = 0
                for i in range(n):
                    b[i] = X[i] + steplength * t[i]
                    if reltest + X[i] == reltest + b[i]:
                        count += 1
                if count < n:
                    f = fminfn(b)
                    funcount += 1
                    accpoint = math.isfinite(f) and (f <= Fmin + gradproj * steplength * acctol)
                    if not accpoint:
                        steplength *= stepredn
                if count == n or accpoint:
                    break
            enough = (f > abstol) and abs(f - Fmin) > reltol * (abs(Fmin) + reltol)
            if not enough:
                count = n
                Fmin = f
            if count < n:
                Fmin = f
                g = np.array(fmingr(b), dtype=float)
                gradcount += 1
                it += 1
                D1 = 0.0
                for i in range(n):
                    t[i] = steplength * t[i]
                    c[i] = g[i] - c[i]
                    D1 += t[i] * c[i]
                if D1 > 0:
                    D2 = 0.0
                    for i in range(n):
                        s = 0.0
                        for j in range(i + 1):
                            s += B[i, j] * c[j]
                        for j in range(i + 1, n):
                            s += B[j, i] * c[j]
                        X[i] = s
                        D2 += s * c[i]
                    D2 = 1.0 + D2 / D1
                    for i in range(n):
                        for j in range(i + 1):
                            B[i, j] += (D2 * t[i] * t[j] - X[i] * t[j] - t[i] * X[j]) / D1
                else:
                    ilast = gradcount
            else:
                if ilast < gradcount:
                    count = 0
                    ilast = gradcount
        else:
            count = 0
            if ilast == gradcount:
                count = n
            else:
                ilast = gradcount
        if it >= maxit:
            break
        if gradcount - ilast > 2 * n:
            ilast = gradcount
        if not (count != n or ilast != gradcount):
            break
    return b, Fmin, (0 if it < maxit else 1)


def r_optim(start, fn, gr=None, method="BFGS", lower=None, upper=None, fnscale=-1.0):
    """stats::optim for the two methods fit() uses, with control = list(fnscale = -1) (R/fit.R:149-150,158)."""
    if method == "Brent":
        x = brent_fmin(lambda p: fn(np.array([p])) / fnscale, lower, upper, math.sqrt(EPS))
        return dict(par=np.array([x]), value=fn(np.array([x])))
    assert gr is not None
    par, val, _fail = vmmin(np.asarray(start, float), lambda p: fn(p) / fnscale,
                            lambda p: np.asarray(gr(p), float) / fnscale)
    return dict(par=par, value=val * fnscale)
