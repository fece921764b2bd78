"""Mock generation and the batched minimiser on the device."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _like(spec):
    from cup1d_b200.likelihood import B200Likelihood

    return B200Likelihood(spec, device=0, emulator_precision="fp64")


def test_batched_starts_equal_scipy_on_the_device_likelihood(golden):
    """every start of the lockstep search follows the path scipy takes with the scalar API
    (Fitter.run_minimizer calls minimize(like.minus_log_prob, ...), fitter.py:356-368)"""
    import copy

    from scipy.optimize import minimize

    from cup1d_b200.minimizer import nelder_mead_batch

    spec, ref = golden("c1_chab_nn")
    like = _like(copy.deepcopy(spec))
    n = like.ndim
    rng = np.random.default_rng(5)
    x0 = np.clip(like.fid_cube[None, :] + 0.1 * (rng.random((4, n)) - 0.5), 0.05, 0.95)
    opts = dict(fatol=0.1, xatol=1e-6, maxiter=120, maxfev=120)
    res = nelder_mead_batch(lambda x: -np.asarray(like.log_prob_batch(x)), x0, **opts)
    for s in range(len(x0)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = minimize(like.minus_log_prob, x0[s], method="Nelder-Mead", bounds=((0.0, 1.0),) * n, options=opts)
        np.testing.assert_array_equal(res["x"][s], r.x)
        assert res["fun"][s] == r.fun and res["nfev"][s] == r.nfev and res["nit"][s] == r.nit
    assert res["n_batches"] < res["n_points"] / 2
    like.close()


def test_mock_and_minimizer_round_trip(golden):
    """noise-free mock at a known point, then run_minimizer recovers chi2 ~ 0 (Mock_P1D
    mock_data.py:96-116 + Fitter.run_minimizer fitter.py:288-464)"""
    import copy

    from cup1d_b200.minimizer import run_minimizer
    from cup1d_b200.mock import make_mock

    spec, ref = golden("c1_chab_nn")
    like = _like(copy.deepcopy(spec))
    truth_x = np.clip(like.fid_cube + 0.03, 0.05, 0.95)
    mock = make_mock(like, values=truth_x, add_noise=False)
    assert like.get_chi2(truth_x) < 1e-20 * like.nd
    np.testing.assert_array_equal(np.concatenate(mock["Pk_kms"]), mock["truth"])
    chi2_start = like.get_chi2(like.fid_cube.copy())
    out = run_minimizer(like, p0=like.fid_cube.copy(), burn_in=True, neval=600, chi2_tol=1e-3, nsamples=12, seed=1)
    assert out["chi2"] < 1e-2 * max(chi2_start, 1.0) and out["chi2"] < chi2_start
    assert out["burn_in_evals"] > 4 * out["burn_in_batches"]  # the 12 starts share their likelihood calls
    # noisy mock: per-z draws from the legacy generator, chi2 at the truth ~ N_d
    noisy = make_mock(like, values=truth_x, add_noise=True, seed=3)
    chi2 = like.get_chi2(truth_x)
    assert 0.5 * like.nd < chi2 < 1.6 * like.nd
    np.testing.assert_array_equal(noisy["full_Pk_kms"], mock["truth"])  # the full vector stays noise free
    like.close()
