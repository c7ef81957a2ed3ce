"""Host side of the kernel API: the six covariance functions of R/GPRclass.R:381-403 and ``cov_func``
(R/GPRclass.R:424-427), kept as user-callable functions with the reference's ``.matrix`` / ``.numeric`` contract.

In R a kernel is an opaque closure ``function(x, y)``.  Closures made by ``cov_func`` (and by ``fit()``) additionally
carry a ``gprc_kernel`` attribute (kernel id + NAMED parameters), which lets GPR/GPC hand the kernel
matrix construction to the CUDA library.  Any other callable is treated exactly like the reference treats it: it is
evaluated on the host through ``covariance_matrix`` (the ``outer`` gather of R/GPRclass.R:355-357) and the resulting
K / K_star are uploaded; factorisation, solves and reductions still run on the device."""
from __future__ import annotations

import numpy as np

from . import _lib


def _is_matrix(x):
    return isinstance(x, np.ndarray) and x.ndim == 2


def _pow(x, p):  # R's `^`: x*x for an exponent of exactly 2
    if np.isscalar(p) and p == 2.0:
        return x * x
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        return np.power(x, p)


def constant(x, y, c):
    return np.full(x.shape[1], float(c)) if _is_matrix(x) else float(c)


def linear(x, y, sigma):
    sigma = np.asarray(sigma, dtype=float)
    if _is_matrix(x):
        s = np.resize(sigma, x.shape[0])[:, None] if sigma.ndim else sigma
        return np.sum(s * x * y, axis=0)
    return float(np.sum(sigma * np.asarray(x, float) * np.asarray(y, float)))


def polynomial(x, y, sigma, p):
    if _is_matrix(x):
        return _pow(np.sum(x * y, axis=0) + sigma, p)
    return float(_pow(np.dot(np.asarray(x, float), np.asarray(y, float)) + sigma, p))


def sqrexp(x, y, l):
    d = (x - y) if _is_matrix(x) else np.asarray(x, float) - np.asarray(y, float)
    r2 = np.sum(d * d, axis=0)
    return np.exp(-r2 / (2 * l * l)) if _is_matrix(x) else float(np.exp(-r2 / (2 * l * l)))


def gammaexp(x, y, l, gamma):
    d = (x - y) if _is_matrix(x) else np.asarray(x, float) - np.asarray(y, float)
    v = np.exp(-_pow(np.sqrt(np.sum(d * d, axis=0)) / l, gamma))
    return v if _is_matrix(x) else float(v)


def rationalquadratic(x, y, l, alpha):
    d = (x - y) if _is_matrix(x) else np.asarray(x, float) - np.asarray(y, float)
    v = _pow(1 + np.sum(d * d, axis=0) / (2 * alpha * (l * l)), -alpha)
    return v if _is_matrix(x) else float(v)


BUILTIN = dict(constant=constant, linear=linear, polynomial=polynomial, sqrexp=sqrexp, gammaexp=gammaexp,
               rationalquadratic=rationalquadratic)
# positional parameter names of each kernel after (x, y): the order fit() passes its vector v (R/fit.R:118)
PARAM_NAMES = dict(constant=("c",), linear=("sigma",), polynomial=("sigma", "p"), sqrexp=("l",),
                   gammaexp=("l", "gamma"), rationalquadratic=("l", "alpha"))
_NAME_OF = {f: n for n, f in BUILTIN.items()}


class KernelSpec:
    """Kernel id + named parameters: what crosses the C ABI as ``gprc_kernel``."""

    def __init__(self, name, **params):
        if name not in BUILTIN:
            raise ValueError("unknown kernel %r" % (name,))
        missing = [p for p in PARAM_NAMES[name] if p not in params]
        if missing:
            raise TypeError("kernel %s needs parameters %s" % (name, missing))
        self.name = name
        self.params = {p: params[p] for p in PARAM_NAMES[name]}

    def to_c(self):
        k = _lib.GprcKernel()
        k.id = _lib.KERNEL_IDS[self.name]
        keep = None
        for p, v in self.params.items():
            if p == "sigma" and self.name == "linear" and np.ndim(v) > 0 and np.size(v) > 1:
                keep = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
                k.sigma_vec = _lib.dptr(keep)
                k.sigma_len = keep.size
            else:
                setattr(k, p, float(np.asarray(v).ravel()[0]))
        return k, keep  # keep the sigma buffer alive while the struct is in use

    def __repr__(self):
        return "KernelSpec(%s, %s)" % (self.name, ", ".join("%s=%r" % kv for kv in self.params.items()))


class CovFunc:
    """A covariance closure ``function(x, y)``; ``gprc_kernel`` is set when it wraps a built-in kernel."""

    def __init__(self, func, args=(), kwargs=None):
        self.func, self.args, self.kwargs = func, tuple(args), dict(kwargs or {})
        self.gprc_kernel = None
        name = _NAME_OF.get(func)
        if name is not None:
            names = PARAM_NAMES[name]
            params = dict(zip(names, self.args))
            params.update(self.kwargs)
            if set(params) == set(names):
                self.gprc_kernel = KernelSpec(name, **params)

    def __call__(self, x, y):
        return self.func(x, y, *self.args, **self.kwargs)


def cov_func(func, *args, **kwargs):
    """``cov_func(func, ...)``, R/GPRclass.R:424-427: fix the parameters of a covariance function."""
    if not callable(func):
        raise TypeError("func must be a function")
    return CovFunc(func, args, kwargs)


def kernel_spec_of(k):
    return getattr(k, "gprc_kernel", None)


def as_matrix(X):
    """``if (!is.matrix(X)) dim(X) <- c(1, length(X))`` (R/GPRclass.R:132, R/GPCclass.R:70)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim < 2:
        X = X.reshape(1, -1)
    return X


def host_covariance_matrix(A, B, k, chunk=1 << 22):
    """R/GPRclass.R:355-357 for opaque closures: one vectorised call of ``k`` on gathered columns (chunked)."""
    A, B = as_matrix(A), as_matrix(B)
    nA, nB = A.shape[1], B.shape[1]
    out = np.empty((nA, nB))
    cols = max(1, chunk // max(nA * A.shape[0], 1))
    for j0 in range(0, nB, cols):
        j1 = min(nB, j0 + cols)
        ii = np.tile(np.arange(nA), j1 - j0)
        jj = np.repeat(np.arange(j0, j1), nA)
        out[:, j0:j1] = np.asarray(k(A[:, ii], B[:, jj]), dtype=float).reshape(j1 - j0, nA).T
    return out


def covariance_matrix(A, B, k, ctx=None):
    """covariance_matrix(A, B, k), R/GPRclass.R:355-357: built on the GPU for built-in kernels."""
    A, B = as_matrix(A), as_matrix(B)
    spec = kernel_spec_of(k)
    if spec is None:
        return host_covariance_matrix(A, B, k)
    ctx = ctx or _lib.default_context()
    if A.shape[0] != B.shape[0]:
        raise ValueError("A and B must have the same number of rows")
    kc, keep = spec.to_c()
    pa, pb = _lib.points(A), _lib.points(B)
    out = np.empty((A.shape[1], B.shape[1]), order="F")
    _lib.check(ctx.lib.gprc_cov_matrix(ctx.handle, kc, _lib.dptr(pa), A.shape[0], A.shape[1], _lib.dptr(pb),
                                       B.shape[1], _lib.dptr(out)))
    return out
