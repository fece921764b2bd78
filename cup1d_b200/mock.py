"""Mock P1D generation.

``Mock_P1D`` (cup1d/p1ds/mock_data.py:21-116) replaces the signal of an observed data set with the
theory prediction at the fiducial model and optionally perturbs it with Gaussian noise drawn from
the per-redshift covariance (``BaseMockP1D.get_Pk_iz_perturbed``, cup1d/p1ds/base_p1d_mock.py:57-75:
``np.random.seed(seed)`` then one ``np.random.multivariate_normal`` draw per redshift, no
correlation between redshifts).  Here the prediction is one row of the batched device evaluation
(``B200Likelihood.get_p1d_kms_batch``) and the noise uses the same legacy NumPy generator, so a
mock made from the same model vector and covariance is the reference's mock.

Like the reference, noise only perturbs the per-redshift ``Pk_kms``: the concatenated
``full_Pk_kms`` handed to ``BaseDataP1D`` stays the noise-free prediction (mock_data.py:96-116).
"""

from __future__ import annotations

import numpy as np


def perturb_Pk(Pk_kms, cov_Pk_kms, nsamples=1, seed=0):
    """``BaseMockP1D.get_Pk_iz_perturbed`` (base_p1d_mock.py:57-75) with a private legacy generator
    seeded like ``np.random.seed(seed)`` (same stream, no global state touched)"""
    rs = np.random.RandomState(seed)
    out = []
    for pk, cov in zip(Pk_kms, cov_Pk_kms):
        draw = rs.multivariate_normal(pk, cov, nsamples)
        out.append(draw[0] if nsamples == 1 else draw)
    return out


def make_mock(like, cov_Pk_kms=None, values=None, add_noise=False, seed=0, install=True):
    """Mock data vector from a ``B200Likelihood``: model P1D at ``values`` (unit cube; default the
    fiducial point) evaluated on the device, plus optional noise from ``cov_Pk_kms`` (list of
    per-z covariances; default: inverses of the context's per-z inverse covariances).

    Returns dict(Pk_kms [list per z], full_Pk_kms [noise-free concatenation], truth [N_d]).
    With ``install`` the per-z mock becomes the context's data vector (``set_data``).
    """
    x = like.fid_cube if values is None else np.asarray(values, dtype=float)
    truth = np.asarray(like.get_p1d_kms_batch(x[None, :], apply_hull=False)[0], dtype=float)
    if not np.all(np.isfinite(truth)):
        raise ValueError("the model is not defined at the requested point (outside the priors)")
    off = like.zoff_d
    Pk = [truth[off[i] : off[i + 1]].copy() for i in range(like.nz)]
    if add_noise:
        if cov_Pk_kms is None:
            cov_Pk_kms = [np.linalg.inv(c) for c in like.spec["data"]["icov"]]
        Pk = perturb_Pk(Pk, cov_Pk_kms, seed=seed)
    if install:
        if add_noise and like.has_full:
            import warnings

            warnings.warn("make_mock(add_noise=True, install=True) on a full-covariance context: the device keeps ONE data "
                          "vector, so the dense chi2 sees the perturbed mock, whereas the reference's Mock_P1D leaves "
                          "full_Pk_kms noise free (mock_data.py:96-116) and its dense chi2 uses the unperturbed vector",
                          stacklevel=2)
        like.set_data(np.concatenate(Pk))
    return {"Pk_kms": Pk, "full_Pk_kms": truth.copy(), "truth": truth}
