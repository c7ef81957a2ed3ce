"""GPU tests written after this round's GPU budget was spent: not yet run on a B200, therefore NOT in the gating `gpu`
tier.  Run them with `python -m pytest tests -m gpu_next` on a GPU box and move each green test to the `gpu` tier."""
import numpy as np
import pytest


@pytest.mark.gpu_next
@pytest.mark.parametrize("cov", ["sqrexp", "gammaexp", "constant", "linear", "polynomial", "rationalquadratic"])
def test_fit_family_inside_the_library_follows_the_host_optimiser(gprc, cov):
    """gprc_fit_family: the optimiser trajectory of R/fit.R:113-162 run inside libgprc with X, y
    resident on the device.  csrc/optim.hpp is bitwise the Python mirror on the CPU tier (tests/test_optim_library.py)
    and both engines evaluate dens / dens_deriv with the same kernels, so par and value must be IDENTICAL."""
    from gprc_b200.fit import Objective, _fit_one
    rng = np.random.default_rng(77)
    X = rng.uniform(-2, 2, (2, 60))
    y = np.sin(X[0]) + 0.5 * X[1] + rng.normal(0, 0.1, 60)
    obj = Objective(X, y, 0.1, minors="cholesky")
    host = _fit_one(obj, cov, engine="host")
    lib = _fit_one(obj, cov, engine="library")
    np.testing.assert_array_equal(np.atleast_1d(lib["par"]), np.atleast_1d(host["par"]))
    assert lib["value"] == float(host["value"])
