"""Dense chi2 on tcgen05 (cup1d_b200/csrc/c1d_chi2_i8.cu: five int8 digit planes per operand, exact int32
accumulation, float64 recombination) against an extended-precision evaluation of d^T C^-1 d
(likelihood.py:956-960) from the device's own model P1D, against the float64 (DMMA) kernel, on ragged batch
sizes, with rejected walkers in the batch, after set_icov, and for independence of a walker's value from
its position in the batch."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL_CHI2 = 2e-10  # design: 5e-11 relative (tools/chi2_ozaki_sim.py)

DENSE_GOLDENS = ["dr1_full_nn", "hcd_const", "nyx_gp_priors"]


def _reference_chi2(spec, like, x):
    """longdouble d^T C^-1 d from the device's model P1D (rows of rejected walkers: NaN)"""
    p = like.get_p1d_kms_batch(x)
    p = p.cpu().numpy() if hasattr(p, "cpu") else np.asarray(p)
    d = (np.asarray(spec["data"]["full_Pk_kms"])[None, :] - p).astype(np.longdouble)
    A = np.asarray(spec["data"]["full_icov"]).astype(np.longdouble)
    return np.sum((d @ A) * d, axis=1).astype(np.float64)


def _check(spec, like, x):
    from cup1d_b200 import _abi

    assert like.ctx.emulator_paths() & _abi.FLAG_CHI2_I8
    like.set_chi2_precision("fp64")
    c64 = np.asarray(like.get_chi2_batch(x))
    like.set_chi2_precision("i8")
    assert like._base_flags & _abi.FLAG_CHI2_I8
    n0 = like.ctx.launch_count()
    c8 = np.asarray(like.get_chi2_batch(x))
    assert like.ctx.launch_count() > n0
    ref = _reference_chi2(spec, like, x)
    ok = np.isfinite(c64)
    assert ok.sum() > 0
    # rejected walkers (out of the cube, outside the hulls): the same values as the float64 path
    assert np.array_equal(c8[~ok], c64[~ok], equal_nan=True)
    # the float64 kernel is exact to ~1e-15 on the same residuals; the extended-precision value from the model P1D
    # (another launch of stage 3) anchors both
    err8 = np.max(np.abs(c8[ok] - c64[ok]) / np.abs(c64[ok]))
    err64 = np.max(np.abs(c64[ok] - ref[ok]) / np.abs(ref[ok]))
    assert err64 <= 1e-9, err64
    assert err8 <= RTOL_CHI2, err8
    return err8


@pytest.mark.parametrize("name", DENSE_GOLDENS)
def test_goldens_dense(golden, name):
    from cup1d_b200.likelihood import B200Likelihood

    spec, extra = golden(name)
    like = B200Likelihood(spec, device=0, emulator_precision="fp64")
    x = np.asarray(extra["x"])
    _check(spec, like, x)
    # the reference's own chi2 (tools/make_golden.py: Likelihood.get_chi2) through the digit-slice kernel
    like.set_chi2_precision("i8")
    c8 = np.asarray(like.get_chi2_batch(x))
    exp = np.asarray(extra["chi2"])
    ok = np.isfinite(exp)
    assert np.max(np.abs(c8[ok] - exp[ok]) / np.abs(exp[ok])) <= 1e-9
    assert np.array_equal(c8[~ok], exp[~ok], equal_nan=True)
    like.close()


@pytest.mark.parametrize("W", [1, 47, 48, 49, 1000, 20011])
def test_c4_ragged_batches(W):
    """C4 (681 points -> K = 704: the largest supported size): one walker .. several groups per CTA"""
    from cup1d_b200 import synth
    from cup1d_b200.workloads import make_likelihood

    spec, like = make_likelihood("c4", device=0, emulator_precision="fp64")
    x = synth.walkers_sweep(spec, W, seed=21 + W, frac_out=0.02 if W > 100 else 0.0)
    _check(spec, like, x)
    like.close()


def test_position_independent_and_default():
    """a walker's chi2 does not depend on where it sits in the batch; the default arithmetic ("auto" -> int8 digit
    slices for the emulator) selects the digit-slice chi2 and holds the north-star log-likelihood tolerance"""
    from cup1d_b200 import _abi, synth
    from cup1d_b200.workloads import make_likelihood

    spec, like = make_likelihood("c4", device=0)
    assert like.emulator_precision == "i8" and like.chi2_precision == "i8"
    assert like._base_flags & _abi.FLAG_CHI2_I8
    x = synth.walkers_sweep(spec, 9000, seed=5)
    a = np.asarray(like.log_prob_batch(x))
    perm = np.random.default_rng(3).permutation(len(x))
    b = np.asarray(like.log_prob_batch(x[perm]))
    assert np.array_equal(a[perm], b)
    like.set_emulator_precision("fp64")
    assert like.chi2_precision == "fp64"
    c = np.asarray(like.log_prob_batch(x))
    ok = c > -1e99
    assert np.max(np.abs(a[ok] - c[ok]) / (1.0 + 1e-5 * np.abs(c[ok]))) <= 1e-4
    like.close()


def test_set_icov_refreshes_digit_planes():
    from cup1d_b200 import synth
    from cup1d_b200.workloads import make_likelihood

    spec, like = make_likelihood("c3", device=0, emulator_precision="fp64")
    x = synth.walkers_sweep(spec, 300, seed=8)
    _check(spec, like, x)
    A = np.asarray(spec["data"]["full_icov"])
    rng = np.random.default_rng(2)
    s = 1.0 + 0.3 * rng.uniform(size=A.shape[0])
    A2 = A * s[:, None] * s[None, :]
    like.set_icov(spec["data"]["icov"], A2)
    spec2 = dict(spec)
    spec2["data"] = dict(spec["data"])
    spec2["data"]["full_icov"] = A2
    _check(spec2, like, x)
    like.close()
