"""The optimisers inside libgprc (csrc/optim.hpp: Brent_fmin, vmmin, optim_until_error -- SURVEY.md 8f-3) against the
Python mirror (_optim.py / fit.optim_until_error) that drives fit() from the host today, through the C ABI with
ctypes callbacks.  Host code only: needs no GPU.  Both are restatements of the same R sources, statement by statement,
so every iterate -- and therefore every result -- is identical, not merely close."""
import ctypes as C
import math

import numpy as np
import pytest

import gprc_b200 as gprc
from gprc_b200 import _lib
from gprc_b200._optim import OptimError, brent_fmin, vmmin
from gprc_b200.fit import optim_until_error

EPS = np.finfo(float).eps


def _wrap(fn, counter=None):
    """Python callable (may raise) -> gprc_objective_fn"""
    def cb(par, npar, user, out):
        if counter is not None:
            counter.append(tuple(par[i] for i in range(npar)))
        try:
            v = fn(np.array([par[i] for i in range(npar)]))
        except (OptimError, ValueError, ZeroDivisionError, FloatingPointError):
            return 1
        v = np.atleast_1d(np.asarray(v, dtype=float))
        for i, x in enumerate(v):
            out[i] = float(x)
        return 0
    return _lib.OBJECTIVE_FN(cb)


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


@pytest.mark.parametrize("f,lo,hi", [
    (lambda x: (x - 2.0) ** 2 + 1.0, 0.0, 10.0),
    (lambda x: math.sin(3.0 * x) + 0.1 * x * x, 0.0, 5.0),
    (lambda x: abs(x - 1.234567) ** 1.5, 0.0, 10.0),
    (lambda x: -math.exp(-(x - 7.0) ** 2) + 1e-3 * x, 0.0, 10.0),
    (lambda x: 1.0 / (x + 1e-3) + x, 0.0, 10.0),
])
def test_brent_fmin_is_bitwise_the_python_mirror(lib, f, lo, hi):
    seen_py, seen_c = [], []
    want = brent_fmin(lambda x: (seen_py.append(x), f(x))[1], lo, hi, math.sqrt(EPS))
    x = C.c_double(0.0)
    cb = _wrap(lambda p: f(float(p[0])), seen_c)
    assert lib.gprc_optim_brent(cb, None, lo, hi, math.sqrt(EPS), C.byref(x)) == 0
    assert x.value == want
    assert [s[0] for s in seen_c] == seen_py          # the same evaluation points, in the same order


def _rosenbrock(p):
    return 100.0 * (p[1] - p[0] * p[0]) ** 2 + (1.0 - p[0]) ** 2


def _rosenbrock_gr(p):
    return np.array([-400.0 * p[0] * (p[1] - p[0] * p[0]) - 2.0 * (1.0 - p[0]), 200.0 * (p[1] - p[0] * p[0])])


def test_vmmin_reproduces_example_optim_of_R(lib):
    # ?optim: optim(c(-1.2, 1), fr, grr, method = "BFGS") -> value 9.594956e-18, counts 110 / 43
    par = np.array([-1.2, 1.0])
    value, counts, fail = C.c_double(0.0), (C.c_int * 2)(), C.c_int(0)
    rc = lib.gprc_optim_vmmin(_wrap(_rosenbrock), _wrap(_rosenbrock_gr), None, _lib.dptr(par), 2, 100, -math.inf,
                              math.sqrt(EPS), C.byref(value), counts, C.byref(fail))
    assert rc == 0 and fail.value == 0
    assert (counts[0], counts[1]) == (110, 43)
    assert value.value == pytest.approx(9.594956e-18, rel=1e-6)
    np.testing.assert_allclose(par, [1.0, 1.0], atol=1e-8)
    want_par, want_val, want_fail = vmmin([-1.2, 1.0], _rosenbrock, _rosenbrock_gr)
    assert value.value == want_val and np.array_equal(par, want_par)


@pytest.mark.parametrize("seed", range(6))
def test_vmmin_is_bitwise_the_python_mirror(lib, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 5))
    A = rng.normal(size=(n, n))
    A = A @ A.T + 0.5 * np.eye(n)
    b = rng.normal(size=n)
    fn = lambda p: float(0.5 * p @ A @ p - b @ p + 0.05 * np.sum(p ** 4))
    gr = lambda p: A @ p - b + 0.2 * p ** 3
    start = rng.normal(size=n)
    want_par, want_val, want_fail = vmmin(start.copy(), fn, gr)
    par = start.copy()
    value, counts, fail = C.c_double(0.0), (C.c_int * 2)(), C.c_int(0)
    assert lib.gprc_optim_vmmin(_wrap(fn), _wrap(gr), None, _lib.dptr(par), n, 100, -math.inf, math.sqrt(EPS),
                                C.byref(value), counts, C.byref(fail)) == 0
    assert np.array_equal(par, want_par) and value.value == want_val and fail.value == want_fail


def test_vmmin_rejects_a_non_finite_start(lib):
    par = np.array([0.0, 0.0])
    value = C.c_double(0.0)
    rc = lib.gprc_optim_vmmin(_wrap(lambda p: math.inf), _wrap(lambda p: p), None, _lib.dptr(par), 2, 100, -math.inf,
                              math.sqrt(EPS), C.byref(value), None, None)
    assert rc == 1                                     # R: "initial value in 'vmmin' is not finite"
    with pytest.raises(OptimError):
        vmmin([0.0, 0.0], lambda p: math.inf, lambda p: p)


def _until_error_c(lib, fn, gr, start, method, lower=0.0, upper=0.0):
    start = np.asarray(start, dtype=float)
    par, value = np.zeros(len(start)), C.c_double(0.0)
    grcb = _wrap(gr) if gr is not None else C.cast(None, _lib.OBJECTIVE_FN)
    rc = lib.gprc_optim_until_error(_wrap(fn), grcb, None, _lib.dptr(start), len(start), method, lower, upper,
                                    _lib.dptr(par), C.byref(value))
    assert rc == 0
    return par, value.value


def _raise():
    raise OptimError("objective failed")


@pytest.mark.parametrize("case", ["smooth", "objective fails on a region", "objective always fails",
                                  "brent with a failing region"])
def test_optim_until_error_matches_fit_R_semantics(lib, case):
    # maximisation (fnscale = -1); failing evaluations score -10000 (R/fit.R:50)
    if case == "smooth":
        fn = lambda p: -((p[0] - 1.5) ** 2 + 2.0 * (p[1] + 0.5) ** 2 + 0.3 * p[0] * p[1])
        gr = lambda p: -np.array([2.0 * (p[0] - 1.5) + 0.3 * p[1], 4.0 * (p[1] + 0.5) + 0.3 * p[0]])
        args = dict(start=[1.0, 1.0], method="BFGS")
    elif case == "objective fails on a region":
        fn = lambda p: _raise() if p[0] > 2.0 else -((p[0] - 3.0) ** 2 + (p[1] - 1.0) ** 2)
        gr = lambda p: -np.array([2.0 * (p[0] - 3.0), 2.0 * (p[1] - 1.0)])
        args = dict(start=[1.0, 1.0], method="BFGS")
    elif case == "objective always fails":
        fn = lambda p: _raise()
        gr = lambda p: np.zeros(2)
        args = dict(start=[1.0, 1.0], method="BFGS")
    else:
        fn = lambda p: _raise() if p[0] < 1.0 else -(p[0] - 4.0) ** 2
        gr = None
        args = dict(start=[1.0], method="Brent", lower=0.0, upper=10.0)
    kw = dict(method=args["method"])
    if gr is not None:
        kw["gr"] = gr
    if args["method"] == "Brent":
        kw.update(lower=args["lower"], upper=args["upper"])
    want = optim_until_error(args["start"], fn, **kw)
    par, value = _until_error_c(lib, fn, gr, args["start"], 0 if args["method"] == "Brent" else 1,
                                args.get("lower", 0.0), args.get("upper", 0.0))
    assert np.array_equal(par, np.atleast_1d(want["par"]))
    assert value == float(want["value"])


def test_gradient_failure_returns_the_best_recorded_evaluation(lib):
    calls = {"n": 0}

    def gr(p):
        calls["n"] += 1
        if calls["n"] >= 3:
            raise OptimError("solve(): computationally singular")
        return -np.array([2.0 * (p[0] - 2.0), 2.0 * (p[1] + 1.0)])

    fn = lambda p: -((p[0] - 2.0) ** 2 + (p[1] + 1.0) ** 2)
    want = optim_until_error([1.0, 1.0], fn, gr=gr, method="BFGS")
    calls["n"] = 0
    par, value = _until_error_c(lib, fn, gr, [1.0, 1.0], 1)
    assert np.array_equal(par, want["par"]) and value == float(want["value"])
