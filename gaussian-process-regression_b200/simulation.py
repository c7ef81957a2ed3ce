"""Callers of the hot path: the data generators of R/simulation.R that define the synthetic workloads
(SURVEY.md section 8d).  Plotting is out of scope; the numeric parts of the drivers are kept so that existing
simulate_* call sites keep working."""
from __future__ import annotations

import sys

import numpy as np

from .gpc import GPC
from .gpr import GPR
from .kernels import as_matrix, covariance_matrix


def iid_noise(distribution, *args, **kwargs):
    """R/simulation.R:375-378: noise function X -> distribution(ncol(X), ...)."""
    if not callable(distribution):
        raise TypeError("distribution must be a function")
    return lambda X: distribution(as_matrix(X).shape[1], *args, **kwargs)


def combine_all(lst):
    """R/simulation.R:338-349."""
    l = len(lst)
    lengths = [len(v) for v in lst]
    prods = np.concatenate([[1], np.cumprod(lengths)])
    rev_prods = np.concatenate([[1], np.cumprod(lengths[::-1])])
    out = np.zeros((l, int(prods[l])))
    for k in range(1, l + 1):
        out[k - 1, :] = np.tile(np.repeat(np.asarray(lst[k - 1], float), int(rev_prods[l - k])), int(prods[k - 1]))
    return out


def _grid(limits, test_size):
    D = limits.shape[0]
    per_dim = int(np.ceil(round(test_size ** (1.0 / D), 9)))  # seq(length.out = x) rounds x up
    return combine_all([np.linspace(limits[i, 0], limits[i, 1], per_dim) for i in range(D)])


def multivariate_normal(n, mean, covariance, tol=1e-6, rng=None):
    """R/GPRclass.R:360-370 (host helper, not on the north-star path): Cholesky with eigen fallback."""
    rng = rng or np.random.default_rng()
    mean = np.asarray(mean, float).ravel()
    covariance = np.asarray(covariance, float)
    if len(mean) != covariance.shape[0]:
        raise ValueError("length(mean) == nrow(covariance) is not TRUE")
    import ctypes as C
    from . import _lib
    m = len(mean)
    Z = np.asfortranarray(rng.standard_normal((m, n)))  # the normals are the caller's (R: rnorm)
    ctx = _lib.default_context()
    out = np.empty((m, n), order="F")
    info = C.c_long(0)
    _lib.check(ctx.lib.gprc_mvn_sample(ctx.handle, _lib.dptr(np.ascontiguousarray(mean)),
                                       _lib.dptr(np.asfortranarray(covariance)), m, _lib.dptr(Z), n, _lib.dptr(out),
                                       C.byref(info)))
    if info.value == 0:
        return out
    # chol() failed: the reference's eigen fallback (R/GPRclass.R:363-368) -- host code in the reference as well
    w, V = np.linalg.eigh(covariance)
    w, V = w[::-1], V[:, ::-1]
    if not np.all(w > -tol * abs(w[0])):
        raise ValueError("all(eigval > -tol * abs(eigval[1])) is not TRUE")
    L = V @ np.diag(np.sqrt(np.maximum(w, 0)))
    return mean[:, None] + L @ Z


def simulate_regression(func, limits, training_points=None, training_size=10, observation_noise=lambda X: 0.0,
                        test_size=10000, rng=None, **gpr_args):
    """R/simulation.R:80-139 without the plots: returns dict(model, test_points, predictions, residual)."""
    rng = rng or np.random.default_rng()
    limits = np.asarray(limits, float)
    if limits.ndim < 2:
        limits = limits.reshape(-1, 2)
    D = limits.shape[0]
    if training_points is None:
        training_points = np.vstack([rng.uniform(limits[i, 0], limits[i, 1], training_size) for i in range(D)])
    training_points = as_matrix(training_points)
    y = np.array([func(training_points[:, i]) for i in range(training_points.shape[1])], float) \
        + observation_noise(training_points)
    model = GPR(training_points, np.asarray(y, float).ravel(), **gpr_args)
    test_points = _grid(limits, test_size)
    predictions = model.predict(test_points, pointwise_var=True)
    truth = np.array([func(test_points[:, i]) for i in range(test_points.shape[1])], float)
    residual = predictions[:, 0] - truth
    print("The mean absolute difference of predictions and ground truth in the considered limits is ",
          np.mean(np.abs(residual)), file=sys.stderr)
    return dict(model=model, test_points=test_points, predictions=predictions, residual=residual)


def simulate_regression_gp(actual_cov, limits, observation_noise=lambda X: 0.0, test_size=300, training_size=10,
                           random_training=True, rng=None, **gpr_args):
    """R/simulation.R:212-255 without the plots."""
    rng = rng or np.random.default_rng()
    limits = np.asarray(limits, float)
    if limits.ndim < 2:
        limits = limits.reshape(-1, 2)
    testpoints = _grid(limits, test_size)
    K = covariance_matrix(testpoints, testpoints, actual_cov)
    f = multivariate_normal(1, np.zeros(K.shape[0]), K, rng=rng)[:, 0]
    if random_training:
        training_set = rng.choice(testpoints.shape[1], training_size, replace=False)
    else:
        training_set = (np.arange(1, training_size + 1) * (testpoints.shape[1] // training_size)) - 1
    X = testpoints[:, training_set]
    y = f[training_set] + observation_noise(X)
    model = GPR(X, np.asarray(y, float).ravel(), **gpr_args)
    predictions = model.predict(testpoints)
    residual = predictions[:, 0] - f
    return dict(model=model, test_points=testpoints, f=f, predictions=predictions, residual=residual)


def simulate_classification(func, limits, training_points=None, training_size=10, test_size=10000, rng=None,
                            **gpc_args):
    """R/simulation.R:311-336 without the plot."""
    rng = rng or np.random.default_rng()
    limits = np.asarray(limits, float)
    if limits.ndim < 2:
        limits = limits.reshape(-1, 2)
    D = limits.shape[0]
    if training_points is None:
        training_points = np.vstack([rng.uniform(limits[i, 0], limits[i, 1], training_size) for i in range(D)])
    training_points = as_matrix(training_points)
    y = np.array([func(training_points[:, i]) for i in range(training_points.shape[1])], float)
    model = GPC(training_points, y, **gpc_args)
    test_points = _grid(limits, test_size)
    predictions = model.predict_class(test_points)
    truth = np.array([func(test_points[:, i]) for i in range(test_points.shape[1])], float)
    residual = 2 * (predictions >= 0.5).astype(int) - 1 - truth
    print("Proportion of misclassified test points.", np.mean(np.abs(residual)) / 2, file=sys.stderr)
    return dict(model=model, test_points=test_points, predictions=predictions, residual=residual)
