"""GPC: host mirror of the R6 class of R/GPCclass.R:55-210.  The Laplace Newton loop runs inside libgprc
(gprc_gpc_fit); the stopping rule and the reference's (mis-signed) divergence guard are applied there literally and
reported back; messages, errors and the logistic-Gaussian quadrature stay on the host as in the reference."""
from __future__ import annotations

import ctypes as C
import math
import sys
import warnings

import numpy as np

from . import _lib
from ._quad import logistic_gaussian_integral
from .gpr import _ReadOnly
from .kernels import as_matrix, host_covariance_matrix, kernel_spec_of


class GPC:
    X = _ReadOnly("X")
    k = _ReadOnly("k")
    y = _ReadOnly("y")
    logq = _ReadOnly("logq")

    def __init__(self, X, y, k, epsilon=1e-5, ctx=None, guard=True, max_iter=0, verbose=True):
        X = np.asarray(X)
        y = np.asarray(y)
        # stopifnot(...) of R/GPCclass.R:67-68,71
        if not (np.issubdtype(X.dtype, np.number) and np.issubdtype(y.dtype, np.number) and y.ndim == 1):
            raise TypeError("is.numeric(X), is.vector(y), is.numeric(y) are not all TRUE")
        if not (isinstance(epsilon, (int, float)) and epsilon > 0 and callable(k)):
            raise TypeError("is.numeric(epsilon), epsilon > 0, is.function(k) are not all TRUE")
        X = as_matrix(X)
        y = y.astype(np.float64)
        if len(y) != X.shape[1]:
            raise ValueError("length(y) == ncol(X) is not TRUE")
        self._ctx = ctx or _lib.default_context()
        lib = self._ctx.lib
        n = len(y)
        spec = kernel_spec_of(k)
        handle = _lib._P()
        iters = C.c_int(0)
        status = C.c_int(0)
        cap = 256
        trace = np.zeros(cap)
        sum_diag = C.c_double(0.0)
        sum_log = C.c_double(0.0)
        if spec is not None:
            kc, keep = spec.to_c()
            xp = _lib.points(X)
            _lib.check(lib.gprc_gpc_fit(self._ctx.handle, kc, _lib.dptr(xp), X.shape[0], n, _lib.dptr(y),
                                        float(epsilon), int(bool(guard)), int(max_iter), C.byref(handle),
                                        C.byref(iters), _lib.dptr(trace), cap, C.byref(sum_diag), C.byref(sum_log),
                                        C.byref(status)))
        else:
            K = np.asfortranarray(host_covariance_matrix(X, X, k))
            _lib.check(lib.gprc_gpc_fit_precomputed(self._ctx.handle, _lib.dptr(K), n, _lib.dptr(y), float(epsilon),
                                                    int(bool(guard)), int(max_iter), C.byref(handle), C.byref(iters),
                                                    _lib.dptr(trace), cap, C.byref(sum_diag), C.byref(sum_log),
                                                    C.byref(status)))
        self.objective_trace = trace[:min(iters.value, cap)].copy()
        self.iterations = iters.value
        if status.value == 1:
            raise RuntimeError("Apparently does not converge.")  # R/GPCclass.R:91
        if status.value != 0:
            raise np.linalg.LinAlgError("the leading minor of B = I + W^1/2 K W^1/2 is not positive definite")
        if verbose:
            print("Convergence after %s iterations" % iters.value, file=sys.stderr)  # message(), R/GPCclass.R:98
        self._handle = handle
        self._X, self._y, self._k = X, y, k
        # logq <- objective - sum(diag(L))  (sum of the diagonal, not of its log: SURVEY.md A.3)
        self._logq = float(self.objective_trace[-1] - sum_diag.value)
        self.sum_log_diagL = sum_log.value
        self._f_hat = None
        self._L = None

    @classmethod
    def new(cls, *args, **kwargs):
        return cls(*args, **kwargs)

    def __del__(self):  # pragma: no cover
        try:
            if getattr(self, "_handle", None) and getattr(self._ctx, "handle", None):
                self._ctx.lib.gprc_gpc_free(self._handle)
                self._handle = None
        except Exception:
            pass

    @property
    def f_hat(self):
        if self._f_hat is None:
            f = np.empty(len(self._y))
            _lib.check(self._ctx.lib.gprc_gpc_get(self._handle, _lib.GET_FHAT, _lib.dptr(f)))
            self._f_hat = f
        return self._f_hat

    @f_hat.setter
    def f_hat(self, value):
        raise AttributeError("`$f_hat` is read only")

    @property
    def L(self):
        if self._L is None:
            n = len(self._y)
            out = np.empty((n, n), order="F")
            _lib.check(self._ctx.lib.gprc_gpc_get(self._handle, _lib.GET_L, _lib.dptr(out)))
            self._L = out
        return self._L

    @L.setter
    def L(self, value):
        raise AttributeError("`$L` is read only")

    def predict_latent(self, X_star):
        """fs_bar and Vfs of R/GPCclass.R:109-115."""
        X_star = np.asarray(X_star, dtype=np.float64)
        if X_star.ndim < 2:
            X_star = X_star.reshape(1, -1)  # always ONE row, unlike GPR$predict (R/GPCclass.R:109)
        if X_star.shape[0] != self._X.shape[0]:
            raise ValueError("non-conformable arguments")
        m = X_star.shape[1]
        fs_bar, Vfs = np.empty(m), np.empty(m)
        lib = self._ctx.lib
        if kernel_spec_of(self._k) is not None:
            xs = _lib.points(X_star)
            _lib.check(lib.gprc_gpc_predict_latent(self._handle, _lib.dptr(xs), m, _lib.dptr(fs_bar), _lib.dptr(Vfs)))
        else:
            Ks = np.asfortranarray(host_covariance_matrix(self._X, X_star, self._k))
            kss = np.ascontiguousarray(np.asarray(self._k(X_star, X_star), dtype=np.float64))
            _lib.check(lib.gprc_gpc_predict_latent_precomputed(self._handle, _lib.dptr(Ks), _lib.dptr(kss), m,
                                                               _lib.dptr(fs_bar), _lib.dptr(Vfs)))
        return fs_bar, Vfs

    _IER_MESSAGES = {1: "maximum number of subdivisions reached", 2: "roundoff error was detected",
                     3: "extremely bad integrand behaviour", 4: "roundoff error is detected in the extrapolation table",
                     5: "the integral is probably divergent", 6: "the input is invalid",
                     -1: "non-finite function value"}

    def predict_class(self, X_star, quadrature="device"):
        """R/GPCclass.R:108-118; the integral passes the latent variance as ``sd`` like the reference (A.1).

        quadrature="device": latent prediction and the QUADPACK dqagi port in one library call
        (gprc_gpc_predict_class); "host": scipy's QUADPACK per point, as the reference loops over integrate()."""
        X_star = np.asarray(X_star, dtype=np.float64)
        if X_star.ndim < 2:
            X_star = X_star.reshape(1, -1)
        if quadrature == "host":
            fs_bar, Vfs = self.predict_latent(X_star)
            return np.array([logistic_gaussian_integral(m, s) for m, s in zip(fs_bar, Vfs)])
        m = X_star.shape[1]
        prob = np.empty(m)
        ier = np.zeros(m, dtype=np.int32)
        lib = self._ctx.lib
        if kernel_spec_of(self._k) is not None:
            if X_star.shape[0] != self._X.shape[0]:
                raise ValueError("non-conformable arguments")
            xs = _lib.points(X_star)
            _lib.check(lib.gprc_gpc_predict_class(self._handle, _lib.dptr(xs), m, _lib.dptr(prob),
                                                  ier.ctypes.data_as(_lib.c_int_p)))
        else:
            fs_bar, Vfs = self.predict_latent(X_star)
            _lib.check(lib.gprc_logistic_gaussian(self._ctx.handle, _lib.dptr(np.ascontiguousarray(fs_bar)),
                                                  _lib.dptr(np.ascontiguousarray(Vfs)), m, _lib.dptr(prob),
                                                  ier.ctypes.data_as(_lib.c_int_p)))
        bad = np.flatnonzero(ier != 0)
        if bad.size:  # integrate(stop.on.error = TRUE)
            raise FloatingPointError(self._IER_MESSAGES.get(int(ier[bad[0]]), "integrate() failed"))
        return prob
