"""Mock noise against the reference's BaseMockP1D.get_Pk_iz_perturbed
(cup1d/p1ds/base_p1d_mock.py:57-75), run in the build container by tools/make_golden_mock.py."""
import os

import numpy as np

from cup1d_b200.mock import perturb_Pk

GOLD = os.path.join(os.path.dirname(__file__), "golden", "mock_noise.npz")


def _load():
    g = np.load(GOLD)
    off = np.concatenate([[0], np.cumsum(g["nk"])])
    Pk = [g["Pk"][off[i] : off[i + 1]] for i in range(len(g["nk"]))]
    cov = [g["cov_%d" % i] for i in range(len(g["nk"]))]
    return g, Pk, cov


def test_same_draws_as_reference():
    g, Pk, cov = _load()
    for seed in (0, 7):
        got = np.concatenate(perturb_Pk(Pk, cov, seed=seed))
        np.testing.assert_array_equal(got, g["perturbed_seed%d" % seed])
    got3 = np.concatenate(perturb_Pk(Pk, cov, nsamples=3, seed=5), axis=1)
    np.testing.assert_array_equal(got3, g["perturbed3_seed5"])


def test_global_numpy_state_untouched_and_noise_scale():
    g, Pk, cov = _load()
    np.random.seed(123)
    a = np.random.random()
    np.random.seed(123)
    perturb_Pk(Pk, cov, seed=0)
    assert np.random.random() == a
    draws = np.concatenate(perturb_Pk(Pk, cov, nsamples=4000, seed=1), axis=1)
    sig = np.sqrt(np.concatenate([np.diag(c) for c in cov]))
    assert np.allclose(draws.std(axis=0) / sig, 1.0, atol=0.06)
    assert np.allclose((draws.mean(axis=0) - g["Pk"]) / sig, 0.0, atol=0.08)
