"""gprc-b200: the Gaussian-process hot path of the R package ``gprc`` (MoHawastaken/Gaussian-Process-Regression) on
B200.  ``libgprc.so`` (csrc/, C ABI in include/gprc.h) is the product; this package is the host-side mirror of the
reference's R6 API used by the tests and the benchmark (R is not available in the build image; the R host code and
the .Call shim that bind the same ABI are in R/ and src/).

Import it with ``importlib.import_module("gaussian-process-regression_b200")`` or through the ``gprc_b200`` alias
module at the repository root."""
from . import _lib
from ._lib import Context, GprcError, default_context
from .kernels import (constant, linear, polynomial, sqrexp, gammaexp, rationalquadratic, cov_func, covariance_matrix,
                      KernelSpec)
from .gpr import GPR, GPR_constant, GPR_linear, GPR_polynomial, GPR_sqrexp, GPR_gammaexp, GPR_rationalquadratic
from .gpc import GPC
from .fit import fit, cov_dict, optim_until_error, Objective, multistart
from .simulation import (iid_noise, combine_all, multivariate_normal, simulate_regression, simulate_regression_gp,
                         simulate_classification)

__all__ = ["GPR", "GPC", "fit", "cov_func", "covariance_matrix", "iid_noise", "simulate_regression",
           "simulate_regression_gp", "simulate_classification", "multivariate_normal", "constant", "linear",
           "polynomial", "sqrexp", "gammaexp", "rationalquadratic", "Context", "GprcError"]
