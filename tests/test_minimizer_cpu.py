"""Batched Nelder-Mead against scipy.optimize.minimize(method="Nelder-Mead"),
the routine Fitter.run_minimizer calls once per start (cup1d/likelihood/fitter.py:356-368)."""
import numpy as np
import pytest
from scipy.optimize import minimize

from cup1d_b200.minimizer import latin_hypercube, nelder_mead_batch, run_minimizer


def _objective(n, seed):
    rng = np.random.default_rng(seed)
    c = 0.2 + 0.6 * rng.random(n)
    A = rng.normal(0, 1, (n, n))
    H = A @ A.T / n + np.eye(n)

    def f(x):
        d = np.atleast_2d(x) - c
        q = np.einsum("mi,ij,mj->m", d, H, d)
        return 40 * q + 3 * np.sin(7 * d).sum(axis=1) ** 2  # bumpy bowl, minimum inside the cube
    return f


@pytest.mark.parametrize("n,maxit", [(3, 400), (7, 1000), (12, 300), (12, 60)])
def test_each_start_reproduces_scipy(n, maxit):
    f = _objective(n, n)
    rng = np.random.default_rng(100 + n)
    x0 = rng.random((9, n))
    x0[0] = 0.999  # initial simplex is reflected off the upper bound
    x0[1, 0] = 0.0  # zero coordinate: zdelt step
    x0[2] = np.clip(x0[2] * 1.4 - 0.2, -0.1, 1.1)  # outside the bounds: clipped like scipy
    res = nelder_mead_batch(f, x0, fatol=0.1, xatol=1e-6, maxiter=maxit, maxfev=maxit)
    for s in range(len(x0)):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = minimize(lambda v: float(f(v)[0]), x0[s], method="Nelder-Mead", bounds=((0.0, 1.0),) * n,
                           options={"fatol": 0.1, "xatol": 1e-6, "maxiter": maxit, "maxfev": maxit})
        np.testing.assert_array_equal(res["x"][s], ref.x)
        assert res["fun"][s] == ref.fun and res["nit"][s] == ref.nit and res["nfev"][s] == ref.nfev
        assert res["status"][s] == ref.status and bool(res["success"][s]) == ref.success
    # lockstep: at most three batched calls per simplex iteration, plus the initial simplex
    assert res["n_batches"] <= 3 * int(res["nit"].max()) + 1
    assert res["n_points"] == int(res["nfev"].sum())


def test_default_limits_and_tiny_budget():
    f = _objective(4, 1)
    x0 = np.full((2, 4), 0.5)
    res = nelder_mead_batch(f, x0)  # maxiter = maxfev = 200 n
    ref = minimize(lambda v: float(f(v)[0]), x0[0], method="Nelder-Mead", bounds=((0.0, 1.0),) * 4)
    np.testing.assert_array_equal(res["x"][0], ref.x)
    assert res["nfev"][0] == ref.nfev
    res = nelder_mead_batch(f, x0, maxfev=3, maxiter=50)  # fewer calls than simplex vertices
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = minimize(lambda v: float(f(v)[0]), x0[0], method="Nelder-Mead", bounds=((0.0, 1.0),) * 4, options={"maxfev": 3, "maxiter": 50})
    np.testing.assert_array_equal(res["x"][0], ref.x)
    assert res["nfev"][0] == ref.nfev == 3 and res["status"][0] == ref.status == 1


def test_latin_hypercube_strata():
    pts = latin_hypercube(5, 25, np.random.default_rng(0))
    assert pts.shape == (25, 5)
    for j in range(5):
        assert sorted(np.floor(pts[:, j] * 25).astype(int).tolist()) == list(range(25))


class _ToyLike(object):
    """quadratic -lnP with the B200Likelihood methods run_minimizer uses"""

    def __init__(self, n):
        self.ndim = n
        self.c = np.linspace(0.3, 0.7, n)
        self.calls = 0

    def log_prob_batch(self, x, zmask=None):
        self.calls += 1
        return -0.5 * self.get_chi2_rows(x)

    def get_chi2_rows(self, x):
        return 400 * np.sum((np.atleast_2d(x) - self.c) ** 2, axis=1)

    def get_chi2(self, x, zmask=None):
        return float(self.get_chi2_rows(x)[0])

    def log_prob(self, x, zmask=None):
        return -0.5 * self.get_chi2(x)


def test_run_minimizer_flow():
    like = _ToyLike(5)
    out = run_minimizer(like, burn_in=True, neval=400, chi2_tol=1e-3, seed=3)
    assert out["chi2"] < 1e-2 and np.max(np.abs(out["mle_cube"] - like.c)) < 1e-2
    assert len(out["burn_in"]) == 25
    # the 25 starts of the burn-in share their likelihood calls
    assert out["n_batches"] == like.calls and out["burn_in_evals"] > 5 * out["burn_in_batches"]
    fixed = run_minimizer(like, p0=np.full(5, 0.4), ind_fix=np.array([1, 3]), neval=400, chi2_tol=1e-3)
    # fixed coordinates are overridden in the evaluation only (likelihood.py:1066-1072): the free ones
    # reach their optimum, and the reported chi2 is get_chi2 at the returned point, like the reference's
    assert np.max(np.abs(np.delete(fixed["mle_cube"], [1, 3]) - np.delete(like.c, [1, 3]))) < 1e-2
    assert fixed["chi2"] == like.get_chi2(fixed["mle_cube"])
