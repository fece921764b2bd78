"""Ensemble sampler driven by the batched likelihood.

The reference runs ``emcee.EnsembleSampler(nwalkers, ndim, like.log_prob_and_blobs,
blobs_dtype=...)`` and calls the likelihood once per walker
(cup1d/likelihood/fitter.py:187-235).  emcee is not installed here, so this
module provides the part of emcee's interface that ``Fitter.run_sampler`` uses
— ``sample`` / ``run_mcmc``, ``get_chain``, ``get_log_prob``, ``get_blobs``
with ``flat`` / ``discard`` / ``thin`` — around the affine-invariant stretch
move of Goodman & Weare (2010) in its parallel two-half form (Foreman-Mackey
et al. 2013, the algorithm emcee's ``StretchMove`` implements), with ONE
vectorised log-probability call per half-ensemble.

``log_prob_fn`` receives ``[N, ndim]`` and returns ``[N, 1 + nblobs]`` rows
``(lnP, *blobs)``: ``B200Likelihood.log_prob_and_blobs_vectorized`` is exactly
that.  With emcee installed the same callable plugs into
``emcee.EnsembleSampler(..., vectorize=True)``.
"""

from __future__ import annotations

import numpy as np


class EnsembleSampler(object):
    def __init__(self, nwalkers, ndim, log_prob_fn, a=2.0, blobs_dtype=None, seed=None, randomize_split=True):
        if nwalkers < 2 * ndim or nwalkers % 2:
            # emcee raises below 2*ndim; an even count keeps the two halves equal
            raise ValueError("need an even number of walkers >= 2 * ndim")
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)
        self.log_prob_fn = log_prob_fn
        self.a = float(a)
        self.blobs_dtype = blobs_dtype
        self.rng = np.random.default_rng(seed)
        self.randomize_split = randomize_split
        self.reset()

    def reset(self):
        self.iteration = 0
        self._chain, self._lnp, self._blobs = [], [], []
        self.naccepted = np.zeros(self.nwalkers)
        self.n_evals = 0

    # ------------------------------------------------------------------
    def _evaluate(self, x):
        rows = np.asarray(self.log_prob_fn(x), dtype=float)
        if rows.ndim == 1:
            rows = rows[:, None]
        self.n_evals += len(x)
        lnp = rows[:, 0].copy()
        if np.any(np.isnan(lnp)):
            raise ValueError("Probability function returned NaN")
        return lnp, rows[:, 1:].copy()

    def _propose(self, s, c):
        """stretch move: q = c_j - (c_j - s) z,  z ~ g(z) on [1/a, a]"""
        ns, nc = len(s), len(c)
        zz = ((self.a - 1.0) * self.rng.random(ns) + 1.0) ** 2.0 / self.a
        factors = (self.ndim - 1.0) * np.log(zz)
        rint = self.rng.integers(nc, size=ns)
        return c[rint] - (c[rint] - s) * zz[:, None], factors

    def sample(self, initial_state, iterations=1, thin_by=1, store=True, skip_initial_state_check=False):
        x = np.array(initial_state, dtype=float)
        if x.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions")
        lnp, blobs = self._evaluate(x)
        nw = self.nwalkers
        for _ in range(iterations):
            for _ in range(thin_by):
                inds = np.arange(nw) % 2
                if self.randomize_split:
                    self.rng.shuffle(inds)
                for split in range(2):
                    S = inds == split
                    q, factors = self._propose(x[S], x[~S])
                    new_lnp, new_blobs = self._evaluate(q)
                    lnpdiff = factors + new_lnp - lnp[S]
                    accepted = np.log(self.rng.random(len(q))) < lnpdiff
                    idx = np.nonzero(S)[0][accepted]
                    x[idx] = q[accepted]
                    lnp[idx] = new_lnp[accepted]
                    if blobs.shape[1]:
                        blobs[idx] = new_blobs[accepted]
                    self.naccepted[idx] += 1
            self.iteration += 1
            if store:
                self._chain.append(x.copy())
                self._lnp.append(lnp.copy())
                self._blobs.append(blobs.copy())
            yield x, lnp, blobs

    def run_mcmc(self, initial_state, nsteps, **kw):
        res = None
        for res in self.sample(initial_state, iterations=nsteps, **kw):
            pass
        return res

    # ------------------------------------------------------------------
    @property
    def acceptance_fraction(self):
        return self.naccepted / max(1, self.iteration) / 1.0

    def _get(self, store, flat, discard, thin):
        arr = np.array(store)[discard::thin]
        if flat:
            arr = arr.reshape((-1,) + arr.shape[2:])
        return arr

    def get_chain(self, flat=False, discard=0, thin=1):
        return self._get(self._chain, flat, discard, thin)

    def get_log_prob(self, flat=False, discard=0, thin=1):
        return self._get(self._lnp, flat, discard, thin)

    def get_blobs(self, flat=False, discard=0, thin=1):
        """structured array when ``blobs_dtype`` is given (emcee's behaviour,
        Theory.get_blobs_dtype lya_theory.py:500-512)"""
        arr = self._get(self._blobs, flat, discard, thin)
        if self.blobs_dtype is None:
            return arr
        out = np.zeros(arr.shape[:-1], dtype=self.blobs_dtype)
        for i, (name, _) in enumerate(self.blobs_dtype):
            out[name] = arr[..., i]
        return out


def get_initial_walkers(ndim, nwalkers, pini=None, rms=0.01, rng=None):
    """Fitter.get_initial_walkers (fitter.py:746-767)"""
    rng = rng if rng is not None else np.random.default_rng()
    p0 = rng.random((nwalkers, ndim))
    centre = np.full(ndim, 0.5) if pini is None else np.asarray(pini, dtype=float)
    p0 = centre[None, :] + p0 * rms
    p0[p0 >= 1.0] = 0.95
    p0[p0 <= 0.0] = 0.05
    return p0


def run_sampler(like, pini=None, nwalkers=None, nburn=0, nsteps=100, thin=1, seed=0, zmask=None):
    """The sampling half of Fitter.run_sampler (fitter.py:158-246) on a
    ``B200Likelihood``: returns dict(lnprob[steps, walkers], chain[steps,
    walkers, ndim], blobs[steps, walkers] structured) after discarding
    ``nburn`` steps and thinning by ``thin``."""
    ndim = like.ndim
    if nwalkers is None:
        nwalkers = 2 * ndim + 2
    nwalkers = max(nwalkers, 2 * ndim + 2)
    nwalkers += nwalkers % 2
    if zmask is None:
        fn = like.log_prob_and_blobs_vectorized
    else:
        def fn(x):
            lnp, blobs = like.log_prob_and_blobs_batch(x, zmask=zmask)
            return np.concatenate([lnp[:, None], blobs], axis=1)
    rng = np.random.default_rng(seed)
    p0 = get_initial_walkers(ndim, nwalkers, pini=pini, rng=rng)
    sampler = EnsembleSampler(nwalkers, ndim, fn, blobs_dtype=like.get_blobs_dtype(), seed=rng)
    sampler.run_mcmc(p0, nburn + nsteps)
    return {
        "lnprob": sampler.get_log_prob(flat=False, discard=nburn, thin=thin),
        "chain": sampler.get_chain(flat=False, discard=nburn, thin=thin),
        "blobs": sampler.get_blobs(flat=False, discard=nburn, thin=thin),
        "acceptance_fraction": sampler.acceptance_fraction,
        "n_evals": sampler.n_evals,
    }
