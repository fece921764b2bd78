"""attach() on REAL reference objects: the unmodified reference (baseline/_ref, vendored by tools/vendor_reference.py, or
/root/reference) builds its own ``Likelihood`` through tools/ref_harness.py; ``cup1d_b200.attach.attach`` rebinds its hot
path to the device; the reference's own ``Fitter.run_minimizer`` (fitter.py:288-464, scipy Nelder-Mead) then runs
unchanged on the device path.  Skipped where no copy of the reference is present."""

import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _reference_likelihood(emulator, nz=3, fit_type="at_a_time_global"):
    from tools import ref_harness as rh
    from tools.make_golden import grids, synth_cov

    zs = np.round(2.2 + 0.2 * np.arange(nz), 10)
    ks, kmins, kmaxs = grids("chab", nz)
    nd = sum(len(k) for k in ks)

    def data_ns(Pk_full, cov_full):
        Pk, cov, off = [], [], 0
        for k in ks:
            Pk.append(Pk_full[off : off + len(k)])
            cov.append(cov_full[off : off + len(k), off : off + len(k)])
            off += len(k)
        return rh.make_data_namespace(zs, ks, kmins, kmaxs, Pk, cov, [0.8 * c for c in cov])

    with contextlib.redirect_stdout(io.StringIO()):
        like0, _ = rh.build_reference_likelihood(emulator, data_ns(np.ones(nd), np.eye(nd) * 0.01), rebin_k=1, fit_type=fit_type, use_ic=False)
        x_fid = np.clip(like0.sampling_point_from_parameters(), 0.03, 0.97)
        p_truth = np.concatenate(like0.get_p1d_kms(values=x_fid, apply_hull=False)[0])
        cov = synth_cov(p_truth, 5)
        dvec = p_truth + 0.3 * np.linalg.cholesky(cov) @ np.random.default_rng(6).normal(0, 1, nd)
        like, _ = rh.build_reference_likelihood(emulator, data_ns(dvec, cov), rebin_k=1, fit_type=fit_type, use_ic=False)
    return like, x_fid


def _need_reference():
    from tools import ref_harness as rh

    if not rh.available():
        pytest.skip("no copy of the reference (baseline/_ref or /root/reference)")
    rh.setup()


def test_reference_fitter_minimizer_on_the_device_path():
    _need_reference()
    from cup1d.likelihood.fitter import Fitter

    from cup1d_b200 import synth
    from cup1d_b200.attach import attach, detach
    from oracle.emulators import RefMLPEmulator

    emu = RefMLPEmulator(synth.synth_nn_emulator(seed=21, nhidden=3, max_neurons=24, head=12), "synthetic_mpg_nn", "mpg")
    like, x_fid = _reference_likelihood(emu)
    rng = np.random.default_rng(2)
    pts = np.clip(x_fid[None, :] + rng.normal(0, 0.05, (12, len(x_fid))), 0.0, 1.0)
    pts[3, 1] = 1.2  # outside the cube
    ref_rows = np.array([like.log_prob_and_blobs(v.copy()) for v in pts])
    p0 = np.clip(x_fid + 0.02, 0.03, 0.97)
    tmp = tempfile.mkdtemp()
    with contextlib.redirect_stdout(io.StringIO()):
        f_ref = Fitter(like=like, nwalkers=4, nsteps=2, nburn=0, rootdir=tmp, subfolder="ref", parallel=False)
        f_ref.run_minimizer(log_func_minimize=like.minus_log_prob, p0=p0.copy(), neval=300)
    chi2_ref, chi2_start = like.get_chi2(f_ref.mle_cube), like.get_chi2(p0)

    glike = attach(like, device=0, emulator_precision="fp64")
    try:
        dev_rows = np.array([like.log_prob_and_blobs(v.copy()) for v in pts])
        bad = ref_rows[:, 0] <= -1e99
        np.testing.assert_array_equal(dev_rows[bad, 0], ref_rows[bad, 0])
        assert np.all(np.abs(dev_rows[~bad, 0] - ref_rows[~bad, 0]) <= 1e-4 + 1e-11 * np.abs(ref_rows[~bad, 0]))
        np.testing.assert_allclose(dev_rows[~bad, 1:], ref_rows[~bad, 1:], rtol=1e-12)
        n0 = glike.ctx.launch_count()
        with contextlib.redirect_stdout(io.StringIO()):
            f_dev = Fitter(like=like, nwalkers=4, nsteps=2, nburn=0, rootdir=tmp, subfolder="dev", parallel=False)
            f_dev.run_minimizer(log_func_minimize=like.minus_log_prob, p0=p0.copy(), neval=300)
        assert glike.ctx.launch_count() > n0 + 100  # the search ran on the device
        chi2_dev = like.get_chi2(f_dev.mle_cube)
    finally:
        detach(like)
    # Nelder-Mead follows the same decisions until rounding differences (1e-13 relative) flip one: same optimum
    assert chi2_dev < 0.2 * chi2_start and chi2_ref < 0.2 * chi2_start
    assert abs(chi2_dev - chi2_ref) <= 0.05 * max(1.0, chi2_ref)
    # after detach the reference's own code answers again
    assert like.log_prob_and_blobs(pts[0].copy())[0] == ref_rows[0, 0]


def test_attach_accepts_a_lace_style_gp_emulator():
    """the reference's default emulator kind (CH24_mpgcen_gpr, input_pipeline.py:21): a scikit-learn GP behind the
    attributes the reference reads (tests/fake_lace.py).  attach() flattens it and the device model agrees with the
    reference Likelihood that calls sklearn.predict through the object's own emulate_p1d_Mpc"""
    _need_reference()
    from fake_lace import FakeLaceGP

    from cup1d_b200.attach import attach, detach

    emu = FakeLaceGP(seed=4)
    like, x_fid = _reference_likelihood(emu, nz=4, fit_type="global_opt")
    rng = np.random.default_rng(3)
    pts = np.clip(x_fid[None, :] + rng.normal(0, 0.04, (10, len(x_fid))), 0.0, 1.0)
    ref_rows = np.array([like.log_prob_and_blobs(v.copy()) for v in pts])
    ref_p = [np.concatenate(like.get_p1d_kms(values=v.copy())[0]) for v in pts[:4]]
    glike = attach(like, device=0)
    try:
        assert glike.spec["emulator"]["kind"] == "gp" and glike.spec["emulator"]["norm_table"] is not None
        dev_rows = np.array([like.log_prob_and_blobs(v.copy()) for v in pts])
        # sklearn's float64 kernel sums and the device's differ in summation order: 1e-9 relative in the model
        assert np.all(np.abs(dev_rows[:, 0] - ref_rows[:, 0]) <= 1e-4 + 1e-9 * np.abs(ref_rows[:, 0]))
        np.testing.assert_allclose(dev_rows[:, 1:], ref_rows[:, 1:], rtol=1e-12)
        for v, pr in zip(pts[:4], ref_p):
            pd = np.concatenate(like.get_p1d_kms(values=v.copy())[0])
            assert np.max(np.abs(pd / pr - 1)) <= 1e-8
    finally:
        detach(like)
