"""The vectorised stretch-move driver (cup1d_b200/sampler.py) on a known
target: statistics of a correlated Gaussian, emcee-style accessors, blobs."""

import numpy as np

from cup1d_b200.sampler import EnsembleSampler, get_initial_walkers


def test_stretch_move_samples_a_gaussian():
    ndim, nw = 4, 64
    rng = np.random.default_rng(3)
    A = rng.normal(0, 1, (ndim, ndim))
    cov = A @ A.T / ndim + 0.3 * np.eye(ndim)
    icov = np.linalg.inv(cov)
    mu = np.array([0.5, -1.0, 2.0, 0.0])
    calls = []

    def logp(x):  # vectorised: [N, ndim] -> rows (lnP, blob0, blob1)
        calls.append(len(x))
        d = x - mu
        lnp = -0.5 * np.einsum("ni,ij,nj->n", d, icov, d)
        return np.column_stack([lnp, x[:, 0] * 2, np.full(len(x), 7.0)])

    s = EnsembleSampler(nw, ndim, logp, blobs_dtype=[("twice_x0", float), ("seven", float)], seed=11)
    p0 = mu + 0.1 * rng.normal(0, 1, (nw, ndim))
    s.run_mcmc(p0, 1500)
    assert set(calls[1:]) == {nw // 2}  # one vectorised call per half-ensemble
    chain = s.get_chain(flat=True, discard=300, thin=5)
    assert chain.shape == ((1500 - 300) // 5 * nw, ndim)
    np.testing.assert_allclose(chain.mean(axis=0), mu, atol=0.06)
    np.testing.assert_allclose(np.cov(chain.T), cov, atol=0.12)
    assert 0.2 < s.acceptance_fraction.mean() < 0.9
    lnp = s.get_log_prob(discard=300, thin=5)
    assert lnp.shape == (240, nw)
    blobs = s.get_blobs(discard=300, thin=5)
    assert blobs.dtype.names == ("twice_x0", "seven") and blobs.shape == (240, nw)
    np.testing.assert_allclose(blobs["twice_x0"], 2 * s.get_chain(discard=300, thin=5)[..., 0])
    assert np.all(blobs["seven"] == 7.0)


def test_initial_walkers_follow_the_reference_rule():
    p0 = get_initial_walkers(5, 20, pini=np.array([0.5, 0.995, 0.0, 0.2, 0.9]), rng=np.random.default_rng(0))
    assert p0.shape == (20, 5)
    assert np.all((p0 > 0) & (p0 < 1))
    assert np.all(p0[:, 0] >= 0.5) and np.all(p0[:, 0] <= 0.51)
    assert np.all(np.isin(p0[:, 1], [0.95]) | (p0[:, 1] < 1.0))
