"""Rebind the hot-path methods of a live reference ``Likelihood`` to the device path.

    like = pip.fitter.like                 # cup1d.likelihood.likelihood.Likelihood, unchanged
    glike = cup1d_b200.attach.attach(like) # compiles it once
    pip.run_minimizer(...)                 # Fitter.run_minimizer -> like.minus_log_prob -> B200
    pip.run_sampler()                      # emcee -> like.log_prob_and_blobs -> B200

Everything else the ``Fitter`` reads from the likelihood (``free_params``, ``data``,
``theory.model_igm``, plotting helpers, fitter.py:158-1097) keeps working because the object is
still the reference's; only the seven methods of the hot path (likelihood.py:722-1072) and
``theory.get_p1d_kms`` (lya_theory.py:610-814) are replaced on the instance.  Calls that use a
switch outside the device path (``return_covar=``, ``hires=``, ``return_contaminants=``:
plot-only) go to the reference's own method.  ``detach`` restores the
instance.
"""

from __future__ import annotations

from .likelihood import B200Likelihood

HOT_METHODS = ("get_p1d_kms", "get_chi2", "get_log_like", "compute_log_prob", "log_prob", "log_prob_and_blobs", "minus_log_prob")
BATCH_METHODS = ("log_prob_batch", "log_prob_and_blobs_batch", "log_prob_and_blobs_vectorized", "get_log_like_batch",
                 "get_chi2_batch", "get_p1d_kms_batch")


def _route(device_method, reference_method):
    def call(*args, **kwargs):
        try:
            return device_method(*args, **kwargs)
        except NotImplementedError:
            if reference_method is None:
                raise
            return reference_method(*args, **kwargs)

    call.__doc__ = device_method.__doc__
    return call


def attach(like, device=None, emulator_precision="auto", spec=None):
    """Compile ``like`` (or the given spec of it) and rebind its hot-path methods; returns the
    ``B200Likelihood``, also kept as ``like._b200``.  The batched entry points are added to the instance."""
    if getattr(like, "_b200", None) is not None:
        raise RuntimeError("this likelihood is already attached")
    glike = B200Likelihood(spec if spec is not None else like, device=device, emulator_precision=emulator_precision)
    saved = {}
    for name in HOT_METHODS:
        saved[name] = like.__dict__.get(name, None)
        setattr(like, name, _route(getattr(glike, name), getattr(like, name, None)))
    for name in BATCH_METHODS:
        saved[name] = like.__dict__.get(name, None)
        setattr(like, name, getattr(glike, name))
    theory = getattr(like, "theory", None)
    if theory is not None:
        saved["theory.get_p1d_kms"] = theory.__dict__.get("get_p1d_kms", None)
        theory.get_p1d_kms = _route(glike.theory.get_p1d_kms, getattr(theory, "get_p1d_kms", None))
    like._b200 = glike
    like._b200_saved = saved
    return glike


def detach(like):
    """Undo ``attach``: the instance gets its class's methods back and the device context is released."""
    glike = getattr(like, "_b200", None)
    if glike is None:
        return
    saved = like._b200_saved
    for name, old in saved.items():
        obj, attr = (like.theory, name.split(".", 1)[1]) if name.startswith("theory.") else (like, name)
        if old is None:
            if attr in obj.__dict__:
                delattr(obj, attr)
        else:
            setattr(obj, attr, old)
    glike.close()
    del like._b200, like._b200_saved
