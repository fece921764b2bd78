"""cup1d_b200.attach: the hot-path methods of a live reference-style likelihood object rebound to the
device path, plot-only switches routed to the object's own methods, detach restoring it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _FakeTheory(object):
    def get_p1d_kms(self, zs, k_kms, like_params=[], **kw):
        return "reference theory"


class _FakeLike(object):
    """stands in for cup1d's Likelihood on the GPU box (the reference cannot travel): same method names"""

    def __init__(self):
        self.theory = _FakeTheory()

    def get_p1d_kms(self, zs=None, _k_kms=None, values=None, return_covar=False, return_blob=False,
                    return_emu_params=False, apply_hull=True, remove=None):
        return "reference get_p1d_kms"

    def log_prob_and_blobs(self, values, ignore_log_det_cov=True, zmask=None):
        return "reference log_prob_and_blobs"

    def minus_log_prob(self, values, zmask=None, ind_fix=None, pfix=None):
        return "reference minus_log_prob"


def test_attach_rebinds_the_hot_path(golden):
    from cup1d_b200.attach import attach, detach
    from oracle.model import OracleLikelihood

    spec, ref = golden("c1_chab_nn")
    like = _FakeLike()
    glike = attach(like, device=0, spec=spec, emulator_precision="fp64")
    x = ref["x"]
    ora = OracleLikelihood(spec)
    i = int(np.argmax(ref["lnprob_blobs"][:, 0]))
    got = like.log_prob_and_blobs(x[i].copy())
    assert abs(got[0] - ref["lnprob_blobs"][i, 0]) <= 1e-4
    np.testing.assert_allclose(got[1:], ref["lnprob_blobs"][i, 1:], rtol=1e-12)
    assert abs(like.minus_log_prob(x[i].copy()) + ref["lnprob_blobs"][i, 0]) <= 1e-4
    assert abs(like.get_chi2(x[i].copy()) - ora.get_chi2(x[i].copy())) <= 2e-4
    rows = like.log_prob_and_blobs_vectorized(x)
    ok = ref["lnprob_blobs"][:, 0] > -1e99
    assert np.all(np.abs(rows[ok, 0] - ref["lnprob_blobs"][ok, 0]) <= 1e-4 + 1e-11 * np.abs(ref["lnprob_blobs"][ok, 0]))
    # the device model through both doors
    p = like.get_p1d_kms(values=x[i].copy())[0]
    assert np.max(np.abs(np.concatenate(p) / ref["p1d"][i] - 1)) <= 1e-5
    zs, ks = np.asarray(spec["zs"]), spec["data"]["k_kms"]
    pt = like.theory.get_p1d_kms(zs, ks, like_params=glike.parameters_from_sampling_point(x[i]), return_blob=False)
    assert np.max(np.abs(np.concatenate(pt) / ref["p1d"][i] - 1)) <= 1e-5
    # plot-only switches stay with the object's own methods
    assert like.get_p1d_kms(values=x[i].copy(), return_covar=True) == "reference get_p1d_kms"
    # ... while remove= is served by a derived device context
    pr = like.get_p1d_kms(values=x[i].copy(), remove={"SiIII_Lya": 0})[0]
    want = np.concatenate(ora.get_p1d_kms(values=x[i].copy(), remove={"SiIII_Lya": 0})[0])
    assert np.max(np.abs(np.concatenate(pr) / want - 1)) <= 1e-5 and np.max(np.abs(want / ref["p1d"][i] - 1)) > 1e-6
    assert like.theory.get_p1d_kms(zs, ks, like_params=[], return_contaminants=True) == "reference theory"
    with pytest.raises(RuntimeError):
        attach(like, device=0, spec=spec, emulator_precision="fp64")
    detach(like)
    assert like.log_prob_and_blobs(x[i]) == "reference log_prob_and_blobs"
    assert like.theory.get_p1d_kms(zs, ks) == "reference theory"
    assert not hasattr(like, "log_prob_batch") and not hasattr(like, "_b200")
