"""Library-side optimiser, device test grids (8f-4), the INT8 path's overflow -> FP64 redo and GPC on
the INT8 path.  Written late in round 1 in the non-gating `gpu_next` tier, run green on a B200
(profiles/r1_pytest_gpu_next_promoted.log: 13 passed) and promoted to the gating `gpu` tier."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("limits,per_dim", [([[-6.0, 6.0]], 1000), ([[-4.0, 4.0], [-1.0, 3.0]], 100),
                                            ([[0.0, 1.0], [-2.0, 2.0], [5.0, 5.5]], 7), ([[1.0, 2.0]], 1)])
def test_device_grid_is_bitwise_the_host_grid(gprc, limits, per_dim):
    """gprc_grid_points against combine_all(seq(...)) of R/simulation.R:101-102, 338-349 as mirrored by
    simulation._grid / combine_all on the host (numpy.linspace arithmetic)."""
    from gprc_b200.simulation import combine_all, grid_points
    lim = np.asarray(limits, float)
    want = combine_all([np.linspace(lim[i, 0], lim[i, 1], per_dim) for i in range(lim.shape[0])])
    got = grid_points(lim, per_dim)
    np.testing.assert_array_equal(got, want)


def test_predict_grid_equals_predict_on_the_host_grid(gprc):
    from gprc_b200.simulation import combine_all
    rng = np.random.default_rng(5)
    X = rng.uniform(-4, 4, (2, 300))
    y = np.sin(X[0]) * np.cos(X[1]) + rng.normal(0, 0.1, 300)
    g = gprc.GPR(X, y, 0.05, gprc.cov_func(gprc.sqrexp, l=1.0))
    lim = np.array([[-4.0, 4.0], [-4.0, 4.0]])
    grid = combine_all([np.linspace(lim[i, 0], lim[i, 1], 60) for i in range(2)])
    np.testing.assert_array_equal(g.predict_grid(lim, 60), g.predict(grid))


def test_int8_overflow_flag_redoes_the_chunk_in_fp64(gprc, ctx):
    """Path 4 bounds |v| by sqrt(k**).  With the exponents lowered by 30 bits (test option) every block row of V overflows
    its fixed-point range: the device flag must fire and the chunk be recomputed by the FP64 substitution, so the result
    is the FP64 result bit for bit."""
    rng = np.random.default_rng(36)
    n, m = 700, 1000
    X = rng.uniform(-2, 2, (3, n))
    y = np.sum(np.sin(X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-2, 2, (3, m))
    g = gprc.GPR(X, y, 0.05, gprc.cov_func(gprc.sqrexp, l=1.0), ctx=ctx)
    out = {}
    for path, shrink in ((2, 0), (4, 30)):
        ctx.set_option(gprc._lib.OPT_PREDICT_PATH, path)
        ctx.set_option(gprc._lib.OPT_INT8_TEST_SHRINK, shrink)
        try:
            out[path] = g.predict(Xs)
        finally:
            ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 0)
            ctx.set_option(gprc._lib.OPT_INT8_TEST_SHRINK, 0)
    np.testing.assert_array_equal(out[4], out[2])


def test_gpc_latent_prediction_on_the_int8_path(gprc, ctx):
    """GPC$predict_class's latent mean / variance (R/GPCclass.R:110-115) go through the same variance pass with the rows of
    K_star scaled by sqrt(W); |v| <= sqrt(k**) holds there too, so path 4 applies unchanged."""
    rng = np.random.default_rng(37)
    n, m = 600, 900
    X = rng.uniform(-4, 4, (2, n))
    y = np.where(np.abs(X[0]) + np.abs(X[1]) > 2.5, 1.0, -1.0)
    Xs = rng.uniform(-4, 4, (2, m))
    g = gprc.GPC(X, y, gprc.cov_func(gprc.sqrexp, l=0.5), 1e-5, verbose=False, ctx=ctx)
    out = {}
    for path in (2, 4):
        ctx.set_option(gprc._lib.OPT_PREDICT_PATH, path)
        try:
            out[path] = np.column_stack(g.predict_latent(Xs))
        finally:
            ctx.set_option(gprc._lib.OPT_PREDICT_PATH, 0)
    np.testing.assert_allclose(out[4][:, 0], out[2][:, 0], rtol=1e-13, atol=0)
    assert np.max(np.abs(out[4][:, 1] - out[2][:, 1])) < 1e-12
