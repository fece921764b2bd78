"""Parity of the CUDA path (through the C ABI) with the reference's own
outputs (goldens) and with the oracle on synthetic full-size configurations.

Tolerances are those of BASELINE.json:north_star: model P1D within 1e-5
relative, log-like within 1e-4 absolute.  The property tests of stages 1, 3, 4, 5 run with the float64
emulator kernel (bit-level statements); the reference goldens and the full-size workloads are also checked
with the default arithmetic (int8 digit slices on tcgen05 where the network shape allows it)."""

import numpy as np
import pytest

from conftest import GOLDEN_NAMES

pytestmark = pytest.mark.gpu

RTOL_P1D = 1e-5
ATOL_LOGLIKE = 1e-4


def _like(spec, emulator_precision="fp64", **kw):
    from cup1d_b200.likelihood import B200Likelihood

    return B200Likelihood(spec, device=0, emulator_precision=emulator_precision, **kw)


REL = {"fp64": 1e-11, "i8": 4e-9}


def _tol(v, prec="fp64"):
    # 1e-4 absolute; far-from-fit walkers have |lnL| ~ 1e5..1e14 where one ulp exceeds it (float64 kernels: 1e-11
    # relative).  The int8 digit-slice emulator carries a 3e-9 relative model error, which the pull of the residuals
    # turns into d lnL <= eps sqrt(2 |lnL|) sqrt(P^T C^-1 P): below 1e-4 near the fit, up to ~2e-9 |lnL| on far-from-fit walkers
    # (|lnL| ~ 1e6..1e7; measured 1.9e-9 in the randomised configurations) - 4e-9 relative.
    return ATOL_LOGLIKE + REL[prec] * np.abs(v)


@pytest.mark.parametrize("prec", ["fp64", "auto"])
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden_reference_outputs(golden, name, prec):
    spec, ref = golden(name)
    like = _like(spec, emulator_precision=prec)
    if prec == "auto":
        if like.emulator_precision == "fp64":
            like.close()
            pytest.skip("GP emulator: float64 is the only arithmetic")
        assert like.emulator_precision == "i8"
    prec = like.emulator_precision
    _tol = lambda v: globals()["_tol"](v, prec)  # noqa: E731
    x = ref["x"]
    exp = ref["lnprob_blobs"]
    bad = exp[:, 0] <= -1e99

    got = like.log_prob_and_blobs_vectorized(x)  # host-buffer C-ABI call
    np.testing.assert_array_equal(got[bad, 0], exp[bad, 0])
    np.testing.assert_array_equal(np.isnan(got[bad, 1:]), np.isnan(exp[bad, 1:]))
    np.testing.assert_array_equal(np.nan_to_num(got[bad, 1:]), np.nan_to_num(exp[bad, 1:]))
    assert np.all(np.abs(got[~bad, 0] - exp[~bad, 0]) <= _tol(exp[~bad, 0]))
    np.testing.assert_allclose(got[~bad, 1:], exp[~bad, 1:], rtol=1e-12)

    # device-tensor entry points
    import torch

    xd = torch.as_tensor(x, device="cuda:0")
    lnp, blobs = like.log_prob_and_blobs_batch(xd)
    np.testing.assert_array_equal(lnp.cpu().numpy(), got[:, 0])
    np.testing.assert_array_equal(np.nan_to_num(blobs.cpu().numpy()), np.nan_to_num(got[:, 1:]))
    np.testing.assert_array_equal(like.log_prob_batch(x), got[:, 0])

    # model P1D, emulator inputs
    incube = ~((x > 1).any(axis=1) | (x < 0).any(axis=1))
    p, emu = like.get_p1d_kms_batch(x[incube], return_emu_params=True)
    refp = ref["p1d"][incube]
    none = np.isnan(refp).all(axis=1)
    assert np.isnan(p[none]).all()
    assert np.max(np.abs(p[~none] / refp[~none] - 1)) <= RTOL_P1D
    np.testing.assert_allclose(emu[~none], ref["emu_call"][incube][~none], rtol=1e-12)

    # log-like, per-z log-like, chi2
    lnl, lnl_z = like.get_log_like_batch(x)
    rl = ref["loglike"]
    inf = np.isinf(rl)
    np.testing.assert_array_equal(lnl[inf], rl[inf])
    assert np.all(np.abs(lnl[~inf] - rl[~inf]) <= _tol(rl[~inf]))
    assert np.all(np.abs(lnl_z[~inf] - ref["loglike_all"][~inf]) <= _tol(ref["loglike_all"][~inf]))
    chi2 = like.get_chi2_batch(x)
    assert np.all(np.abs(chi2[~inf] - ref["chi2"][~inf]) <= 2 * _tol(ref["chi2"][~inf]))
    assert np.all(np.isinf(chi2[inf]))

    # zmask (the reference itself fails for zmask + IC correction: NaN in the fixture)
    if not np.isnan(ref["lnprob_zmask"]).any():
        gz = like.log_prob_batch(x, zmask=ref["zmask"])
        ez = ref["lnprob_zmask"]
        badz = ez <= -1e99
        np.testing.assert_array_equal(gz[badz], ez[badz])
        assert np.all(np.abs(gz[~badz] - ez[~badz]) <= _tol(ez[~badz]))
    like.close()


def test_scalar_api_conventions(golden):
    """same names / return conventions as likelihood.py:722-1072"""
    spec, ref = golden("nyx_gp_priors")
    like = _like(spec)
    x = ref["x"]
    exp = ref["lnprob_blobs"]
    for i in [0, 5, len(x) - 1, int(np.argmin(exp[:, 0]))]:
        v = x[i].copy()
        out = like.log_prob_and_blobs(v)
        assert len(out) == 7
        if exp[i, 0] <= -1e99:
            assert out[0] == exp[i, 0]
        else:
            assert abs(out[0] - exp[i, 0]) <= _tol(exp[i, 0])
        assert like.log_prob(v) == out[0]
        assert like.minus_log_prob(v) == -out[0]
        ll = like.get_log_like(v, return_blob=True)
        if np.isinf(ref["loglike"][i]):
            assert ll[0] == -np.inf and ll[1] == -np.inf and ll[2] == (0, 0, 0, 0, 0, 0)
            assert like.get_chi2(v) == np.inf
            incube = not ((v > 1).any() or (v < 0).any())
            if incube:
                assert like.get_p1d_kms(values=v) is None
        else:
            assert ll[1].shape == (1, len(spec["zs"]))
            assert abs(ll[0] - ref["loglike"][i]) <= _tol(ref["loglike"][i])
            chi2, chi2_all = like.get_chi2(v, return_all=True)
            assert abs(chi2 - ref["chi2"][i]) <= 2 * _tol(ref["chi2"][i])
            res = like.get_p1d_kms(values=v, return_blob=True, return_emu_params=True)
            assert len(res) == 3 and len(res[0]) == len(spec["zs"])
            assert np.max(np.abs(np.concatenate(res[0]) / ref["p1d"][i] - 1)) <= RTOL_P1D
    # ind_fix / pfix of minus_log_prob (likelihood.py:1069-1070) acts in place
    v = x[0].copy()
    like.minus_log_prob(v, ind_fix=3, pfix=0.5)
    assert v[3] == 0.5
    assert like.get_blobs_dtype()[0] == ("Delta2_star", float)
    like.close()


def _synthetic(kind):
    from cup1d_b200 import synth

    if kind == "c1":  # sam_sim-style mock: NN emulator, Chabanier grid, block-diagonal cov
        emu = synth.synth_nn_emulator(seed=11)
        spec = synth.build_spec(emu, grid="chab", rebin_k=1, full_cov=False, seed=14)
    elif kind == "c2":  # GP emulator, DR1 grid, full covariance
        emu = synth.synth_gp_emulator(seed=12, n_train=330)
        spec = synth.build_spec(emu, grid="dr1", rebin_k=8, full_cov=True, seed=14)
    elif kind == "c4":  # DR1-like: NN emulator, rebin 8, dense 681^2 inverse covariance
        emu = synth.synth_nn_emulator(seed=11)
        spec = synth.build_spec(emu, grid="dr1", rebin_k=8, full_cov=True, seed=14)
    else:  # nyx-like full nuisance set: 7 inputs, nrun free, HCD_const free
        emu = synth.synth_nn_emulator(seed=13, suite="nyx", ndeg=6)
        spec = synth.build_spec(emu, grid="dr1", rebin_k=8, full_cov=True, vary_nrun=True, nz_igm=6, hcd_const=True, gauss_priors=True, seed=14)
    like = _like(spec)
    p_truth = like.get_p1d_kms_batch(synth.truth_point(spec)[None, :])[0]
    synth.attach_mock_data(spec, p_truth, add_noise=(kind != "c1"))
    like.close()
    return spec


@pytest.mark.parametrize("prec", ["fp64", "auto"])
@pytest.mark.parametrize("kind", ["c1", "c2", "c4", "c3"])
def test_synthetic_full_size_vs_oracle(kind, prec):
    from cup1d_b200 import synth
    from oracle.model import OracleLikelihood

    spec = _synthetic(kind)
    like = _like(spec, emulator_precision=prec)
    if prec == "auto" and like.emulator_precision == "fp64":
        like.close()
        pytest.skip("GP emulator: float64 is the only arithmetic")
    prec = like.emulator_precision
    _tol = lambda v: globals()["_tol"](v, prec)  # noqa: E731
    ora = OracleLikelihood(spec)
    x = np.concatenate([synth.walkers_ball(spec, 40, seed=0), synth.walkers_sweep(spec, 88, seed=16, frac_out=0.05)])
    got = like.log_prob_and_blobs_vectorized(x)
    exp = ora.log_prob_and_blobs_many(x)
    bad = exp[:, 0] <= -1e99
    np.testing.assert_array_equal(got[bad, 0], exp[bad, 0])
    assert np.all(np.abs(got[~bad, 0] - exp[~bad, 0]) <= _tol(exp[~bad, 0])), np.max(np.abs(got[~bad, 0] - exp[~bad, 0]))
    np.testing.assert_allclose(got[~bad, 1:], exp[~bad, 1:], rtol=1e-12)
    p = like.get_p1d_kms_batch(x[~bad][:32])
    for i, v in enumerate(x[~bad][:32]):
        po = np.concatenate(ora.get_p1d_kms(values=v)[0])
        assert np.max(np.abs(p[i] / po - 1)) <= RTOL_P1D
    like.close()


def test_full_size_properties():
    """size-independent properties at the benchmark size (C4, 65 536 walkers)"""
    import torch

    from cup1d_b200 import synth

    spec = _synthetic("c4")
    # (a) block-diagonal full covariance: the dense chi2 kernel must equal the sum of the
    #     per-z chi2 computed inside the fused kernel
    from scipy.linalg import block_diag

    spec_bd = dict(spec)
    spec_bd["data"] = dict(spec["data"])
    spec_bd["data"]["full_icov"] = block_diag(*spec["data"]["icov"])
    like = _like(spec_bd)
    W = 65536
    x = synth.walkers_sweep(spec, W, seed=16)
    xd = torch.as_tensor(x, device="cuda:0")
    lnl, lnl_z = like.get_log_like_batch(xd)
    ok = torch.isfinite(lnl)
    assert int((~ok).sum()) == int(round(0.01 * W))  # the rows put outside the cube
    rel = (lnl[ok] - lnl_z[ok].sum(dim=1)).abs() / lnl[ok].abs().clamp(min=1.0)
    assert float(rel.max()) < 1e-11
    assert bool((lnl[ok] <= 0).all())
    # (b) row independence: any batch split gives bit-identical results within one form of stage 3
    #     (production picks the form by batch size: the two agree to rounding, not bit for bit)
    import os

    for layout in ("lanes", "split", "warp"):
        os.environ["C1D_P1D_LAYOUT"] = layout
        try:
            lnp_all = like.log_prob_batch(xd)
            lnp_a = like.log_prob_batch(xd[:1000])
            lnp_b = like.log_prob_batch(xd[1000:1777])
            assert torch.equal(lnp_all[:1000], lnp_a) and torch.equal(lnp_all[1000:1777], lnp_b)
            # (c) permutation equivariance
            perm = torch.randperm(4096, device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(1))
            assert torch.equal(like.log_prob_batch(xd[:4096][perm]), lnp_all[:4096][perm])
        finally:
            del os.environ["C1D_P1D_LAYOUT"]
    big = like.log_prob_batch(xd)          # walker-per-lane form
    small = like.log_prob_batch(xd[:200])  # warp-per-walker form
    fin = torch.isfinite(big[:200]) & (big[:200] > -1e99)
    assert float(((big[:200] - small)[fin].abs() / big[:200][fin].abs().clamp(min=1.0)).max()) < 1e-10
    like.close()
    # (d) noise-free mock: chi2 at the truth is zero (mock generation round trip)
    like = _like(spec)
    xt = synth.truth_point(spec)
    p_truth = like.get_p1d_kms_batch(xt[None, :])[0]
    like.set_data(p_truth)
    assert abs(like.get_chi2(xt)) < 1e-18 * 681
    like.close()


def test_set_icov_replaces_the_metric(golden):
    """B200Likelihood.set_icov: new per-z and full inverse covariances take effect
    on the device exactly as in an oracle rebuilt with them"""
    import copy

    from oracle.model import OracleLikelihood

    spec, ref = golden("dr1_full_nn")
    spec2 = copy.deepcopy(spec)
    rng = np.random.default_rng(3)
    new_blocks = []
    for c in spec["data"]["icov"]:
        s = 1.0 + 0.3 * rng.random(len(c))
        new_blocks.append(c * s[:, None] * s[None, :])
    s = 1.0 + 0.3 * rng.random(len(spec["data"]["full_icov"]))
    new_full = spec["data"]["full_icov"] * s[:, None] * s[None, :]
    spec2["data"]["icov"], spec2["data"]["full_icov"] = new_blocks, new_full
    like = _like(copy.deepcopy(spec))
    x = ref["x"]
    before = like.log_prob_batch(x)
    like.set_icov(new_blocks, new_full)
    after = like.log_prob_batch(x)
    ora = OracleLikelihood(spec2)
    exp = ora.log_prob_and_blobs_many(x)[:, 0]
    ok = exp > -1e99
    assert np.all(np.abs(after[ok] - exp[ok]) <= _tol(exp[ok]))
    assert np.max(np.abs(after[ok] - before[ok])) > 1.0  # it did change
    zm = [float(spec["zs"][1])]
    ll_z = like.get_log_like_batch(x, zmask=zm)[0]
    exp_z = np.array([ora.get_log_like(values=v, zmask=zm)[0] if o else np.nan for v, o in zip(x, ok)])
    assert np.all(np.abs(ll_z[ok] - exp_z[ok]) <= _tol(exp_z[ok]))
    with pytest.raises(ValueError):
        like.set_icov(new_blocks[:-1], new_full)
    with pytest.raises(ValueError):
        like.set_icov(new_blocks, None)
    like.close()


def test_edge_batches(golden):
    """empty, single-row and ragged batches; a batch that crosses the internal chunk size (131 072 walkers);
    every row rejected; non-contiguous / float32 inputs"""
    import torch

    from oracle.model import OracleLikelihood

    spec, ref = golden("c1_chab_nn")
    like = _like(spec)
    ora = OracleLikelihood(spec)
    n = like.ndim
    # empty
    lnp, blobs = like.log_prob_and_blobs_batch(np.zeros((0, n)))
    assert lnp.shape == (0,) and blobs.shape == (0, 6)
    assert like.get_p1d_kms_batch(np.zeros((0, n))).shape == (0, like.nd)
    # one row, and ragged sizes around the tile sizes of the kernels (24/64/72/128 rows, 32-walker items)
    x = ref["x"]
    exp = ref["lnprob_blobs"][:, 0]
    for W in (1, 7, 23, 25, 33, 47):
        got = like.log_prob_batch(x[:W])
        ok = exp[:W] > -1e99
        assert np.all(np.abs(got[ok] - exp[:W][ok]) <= _tol(exp[:W][ok]))
        np.testing.assert_array_equal(got[~ok], exp[:W][~ok])
    # every row outside the cube: exact min_log_like, NaN blobs (likelihood.py:989-994), no model; in both
    # forms of stage 3 (a whole 32-walker item is rejected in the walker-per-lane form)
    import os

    bad = np.full((37, n), 1.5)
    for layout in ("lanes", "split", "warp"):
        os.environ["C1D_P1D_LAYOUT"] = layout
        try:
            lnp, blobs = like.log_prob_and_blobs_batch(bad)
            assert np.all(lnp == like.min_log_like) and np.all(np.isnan(blobs))
            assert np.isnan(like.get_p1d_kms_batch(bad)).all()
            mixed = np.concatenate([bad, x[:3], bad[:2]])
            pm = like.get_p1d_kms_batch(mixed)
            assert np.isnan(pm[:37]).all() and np.isnan(pm[40:]).all()
            np.testing.assert_array_equal(pm[37:40], like.get_p1d_kms_batch(x[:3]))
        finally:
            del os.environ["C1D_P1D_LAYOUT"]
    # crossing the chunk boundary: rows are independent, so a tiled batch repeats its block bit for bit
    # (within one form of stage 3: the short last chunk would otherwise take the warp-per-walker kernel)
    reps = 131072 // len(x) + 2
    big = torch.as_tensor(np.tile(x, (reps, 1)), device="cuda:0")
    assert big.shape[0] > 131072
    for layout in ("lanes", "split", "warp"):
        os.environ["C1D_P1D_LAYOUT"] = layout
        try:
            out = like.log_prob_batch(big)
            first = out[: len(x)]
            assert torch.equal(out.view(reps, len(x)), first[None, :].expand(reps, len(x)))
            np.testing.assert_array_equal(first.cpu().numpy(), like.log_prob_batch(x))
        finally:
            del os.environ["C1D_P1D_LAYOUT"]
    # strided and float32 inputs are converted, not misread
    xs = np.asfortranarray(x[:16])
    np.testing.assert_array_equal(like.log_prob_batch(xs), like.log_prob_batch(x[:16]))
    x32 = x[:16].astype(np.float32)
    np.testing.assert_array_equal(like.log_prob_batch(x32), like.log_prob_batch(x32.astype(np.float64)))
    # scalar API on a rejected point keeps the reference's conventions
    assert ora.log_prob(bad[0].copy()) == like.log_prob(bad[0].copy()) == like.min_log_like
    like.close()


def test_host_entry_point_pipelines_large_batches():
    """c1d_loglike_batch_host uploads batches of 32 768 walkers or more in two pieces on its copy stream
    (and splits beyond 131 072 walkers into chunks): the rows must equal the device-tensor path bit for bit,
    with pinned and with pageable host buffers"""
    import torch

    from cup1d_b200 import synth

    spec = _synthetic("c4")
    like = _like(spec)
    for W in (40000, 140000):
        x = synth.walkers_sweep(spec, W, seed=31, frac_out=0.01)
        xd = torch.as_tensor(x, device="cuda:0")
        lnp, blobs = like.log_prob_and_blobs_batch(xd)
        want = torch.cat([lnp[:, None], blobs], dim=1).cpu().numpy()
        got = like.log_prob_and_blobs_vectorized(x)  # pageable
        np.testing.assert_array_equal(np.nan_to_num(got, nan=-7.0), np.nan_to_num(want, nan=-7.0))
        xp = torch.from_numpy(x).pin_memory()
        out = torch.empty((W, 7), dtype=torch.float64).pin_memory()
        like.log_prob_and_blobs_vectorized(xp.numpy(), out=out.numpy())
        np.testing.assert_array_equal(np.nan_to_num(out.numpy(), nan=-7.0), np.nan_to_num(want, nan=-7.0))
        # a second call reuses the staging buffers: still the same rows
        like.log_prob_and_blobs_vectorized(xp.numpy(), out=out.numpy())
        np.testing.assert_array_equal(np.nan_to_num(out.numpy(), nan=-7.0), np.nan_to_num(want, nan=-7.0))
    like.close()


@pytest.mark.parametrize("kernel", ["lanes", "pairs", "global"])
@pytest.mark.parametrize("name", ["nyx_gp_priors", "hcd_const"])
def test_every_gp_kernel_on_goldens(golden, monkeypatch, name, kernel):
    """the three forms of the GP emulator kernel (lane = row with the training set split over four warps,
    two rows per warp, row per warp from global memory; chosen by batch size in production) are all held
    to the reference goldens, and a row's result does not depend on the batch it travels in"""
    import torch

    monkeypatch.setenv("C1D_GP_KERNEL", kernel)
    spec, ref = golden(name)
    like = _like(spec)
    x, exp = ref["x"], ref["lnprob_blobs"]
    got = like.log_prob_and_blobs_vectorized(x)
    bad = exp[:, 0] <= -1e99
    np.testing.assert_array_equal(got[bad, 0], exp[bad, 0])
    assert np.all(np.abs(got[~bad, 0] - exp[~bad, 0]) <= _tol(exp[~bad, 0]))
    incube = ~((x > 1).any(axis=1) | (x < 0).any(axis=1))
    p = like.get_p1d_kms_batch(x[incube])
    refp = ref["p1d"][incube]
    none = np.isnan(refp).all(axis=1)
    assert np.max(np.abs(p[~none] / refp[~none] - 1)) <= RTOL_P1D
    big = torch.as_tensor(np.tile(x, (40, 1)), device="cuda:0")
    out = like.log_prob_batch(big).view(40, len(x))
    assert torch.equal(out, out[:1].expand(40, len(x)))
    np.testing.assert_array_equal(out[0].cpu().numpy(), got[:, 0])
    like.close()


@pytest.mark.parametrize("layout", ["lanes", "split"])
@pytest.mark.parametrize("name", ["c1", "c3", "c4"])
def test_stage3_segments_are_bit_identical(monkeypatch, name, layout):
    """medium batches cut every redshift bin of the walker-per-lane stage 3 into 2, 4 or 8 k-segments
    (C1D_P1D_PARTS forces the cut): log-posteriors, per-z log-likelihoods and the model P1D must not
    change by a single bit, rejected and ragged rows included"""
    import torch

    from cup1d_b200 import synth

    spec = _synthetic(name)
    like = _like(spec)
    x = np.concatenate([synth.walkers_sweep(spec, 1500, seed=21, frac_out=0.02), synth.walkers_ball(spec, 77, seed=3)])
    x[40:72] = 1.5  # one whole 32-walker item outside the cube
    xd = torch.as_tensor(x, device="cuda:0")
    monkeypatch.setenv("C1D_P1D_LAYOUT", layout)
    ref = None
    for parts in (1, 2, 4, 8):
        monkeypatch.setenv("C1D_P1D_PARTS", str(parts))
        lnp, blobs = like.log_prob_and_blobs_batch(xd)
        lnl, lnl_z = like.get_log_like_batch(xd)
        p = like.get_p1d_kms_batch(xd)
        out = [t.clone() for t in (lnp, blobs, lnl, lnl_z, p)]
        if ref is None:
            ref = out
            assert int(torch.isfinite(lnl).sum()) > 1000
            continue
        for a, b in zip(ref, out):
            assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)) and torch.equal(torch.isnan(a), torch.isnan(b))
    like.close()


@pytest.mark.parametrize("layout", ["lanes", "split", "warp"])
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_both_stage3_forms_on_goldens(golden, monkeypatch, name, layout):
    """the walker-per-lane and the warp-per-walker stage 3 (chosen by batch size in production) are both
    held to the reference goldens: C1D_P1D_LAYOUT forces one form"""
    monkeypatch.setenv("C1D_P1D_LAYOUT", layout)
    spec, ref = golden(name)
    like = _like(spec)
    x = ref["x"]
    exp = ref["lnprob_blobs"]
    got = like.log_prob_and_blobs_vectorized(x)
    bad = exp[:, 0] <= -1e99
    np.testing.assert_array_equal(got[bad, 0], exp[bad, 0])
    assert np.all(np.abs(got[~bad, 0] - exp[~bad, 0]) <= _tol(exp[~bad, 0]))
    incube = ~((x > 1).any(axis=1) | (x < 0).any(axis=1))
    p = like.get_p1d_kms_batch(x[incube])
    refp = ref["p1d"][incube]
    none = np.isnan(refp).all(axis=1)
    assert np.isnan(p[none]).all()
    assert np.max(np.abs(p[~none] / refp[~none] - 1)) <= RTOL_P1D
    # per-z log-likelihoods (k_chi2z_dmma behind the walker-per-lane form) and a z-masked call
    lnl, lnl_z = like.get_log_like_batch(x)
    rl = ref["loglike"]
    inf = np.isinf(rl)
    assert np.all(np.isinf(lnl[inf]))
    assert np.all(np.abs(lnl[~inf] - rl[~inf]) <= _tol(rl[~inf]))
    assert np.all(np.abs(lnl_z[~inf] - ref["loglike_all"][~inf]) <= _tol(ref["loglike_all"][~inf]))
    if not np.isnan(ref["lnprob_zmask"]).any():
        gz = like.log_prob_batch(x, zmask=ref["zmask"])
        ez = ref["lnprob_zmask"]
        okz = ez > -1e99
        np.testing.assert_array_equal(gz[~okz], ez[~okz])
        assert np.all(np.abs(gz[okz] - ez[okz]) <= _tol(ez[okz]))
    like.close()


def test_gp_large_batch_kernel():
    """the shared-memory GP kernel (two rows per warp, table-based exp) takes over above ~215 walkers:
    same tolerances against the oracle as the row-per-warp kernel"""
    from cup1d_b200 import synth
    from oracle.model import OracleLikelihood

    spec = _synthetic("c2")
    like = _like(spec)
    ora = OracleLikelihood(spec)
    x = np.concatenate([synth.walkers_ball(spec, 600, seed=3), synth.walkers_sweep(spec, 425, seed=4, frac_out=0.02)])
    got = like.log_prob_and_blobs_vectorized(x)       # 1025 walkers: large-batch kernels, odd row count
    small = like.log_prob_and_blobs_vectorized(x[:64])  # row-per-warp GP kernel
    idx = np.r_[0:24, 600:640, len(x) - 1]
    exp = ora.log_prob_and_blobs_many(x[idx])
    bad = exp[:, 0] <= -1e99
    np.testing.assert_array_equal(got[idx][bad, 0], exp[bad, 0])
    assert np.all(np.abs(got[idx][~bad, 0] - exp[~bad, 0]) <= _tol(exp[~bad, 0]))
    fin = small[:, 0] > -1e99
    assert np.all(np.abs(got[:64][fin, 0] - small[fin, 0]) <= _tol(small[fin, 0]))
    p = like.get_p1d_kms_batch(x[idx][~bad][:16])
    for i, v in enumerate(x[idx][~bad][:16]):
        po = np.concatenate(ora.get_p1d_kms(values=v)[0])
        assert np.max(np.abs(p[i] / po - 1)) <= RTOL_P1D
    like.close()


def test_non_blocking_stream_first_call(golden, monkeypatch):
    """the first evaluation of a fresh context (workspace allocation and its memsets, the finalisation copies) and a
    set_icov re-upload, issued on a non-blocking side stream: same rows as on the default stream"""
    import torch

    monkeypatch.setenv("C1D_P1D_LAYOUT", "warp")  # one form of stage 3 whatever the batch size: rows comparable bit for bit

    spec, ref = golden("dr1_full_nn")
    x = np.tile(ref["x"], (40, 1))
    want_like = _like(spec)
    want = want_like.log_prob_and_blobs_vectorized(torch.as_tensor(x, device="cuda:0")).cpu().numpy()
    want_like.close()
    side = torch.cuda.Stream(device="cuda:0")
    like = _like(spec)
    with torch.cuda.stream(side):
        xd = torch.as_tensor(x, device="cuda:0")
        got = like.log_prob_and_blobs_vectorized(xd)
        # a larger batch: the workspace grows (new allocation and memsets) under the same stream
        big = like.log_prob_and_blobs_vectorized(torch.cat([xd] * 3))
        like.set_icov(spec["data"]["icov"], spec["data"]["full_icov"])
        again = like.log_prob_and_blobs_vectorized(xd)
    side.synchronize()
    np.testing.assert_array_equal(np.nan_to_num(got.cpu().numpy(), nan=-7.0), np.nan_to_num(want, nan=-7.0))
    np.testing.assert_array_equal(np.nan_to_num(big[: len(x)].cpu().numpy(), nan=-7.0), np.nan_to_num(want, nan=-7.0))
    np.testing.assert_array_equal(np.nan_to_num(again.cpu().numpy(), nan=-7.0), np.nan_to_num(want, nan=-7.0))
    like.close()


@pytest.mark.parametrize("name", ["c1_chab_nn", "hcd_const"])
def test_log_prob_with_log_det_cov(golden, name):
    """log_prob(..., ignore_log_det_cov=False) (likelihood.py:1026-1034, the term of :950 / :962): scalar and
    batched calls against the oracle, with and without a z mask; rejected rows stay at min_log_like"""
    from oracle.model import OracleLikelihood

    spec, ref = golden(name)
    like = _like(spec)
    ora = OracleLikelihood(spec)
    x = ref["x"]
    zm = [spec["zs"][1], spec["zs"][min(3, len(spec["zs"]) - 1)]]
    for zmask in (None, zm):
        exp = np.array([ora.log_prob(v.copy(), ignore_log_det_cov=False, zmask=zmask) for v in x])
        if not np.all(np.isfinite(exp)):
            continue  # the reference's det overflowed (dense matrix); the device path stays finite by design
        got = like.log_prob_batch(x, zmask=zmask, ignore_log_det_cov=False)
        bad = exp <= -1e99
        np.testing.assert_array_equal(got[bad], exp[bad])
        assert np.all(np.abs(got[~bad] - exp[~bad]) <= _tol(exp[~bad]))
        plain = like.log_prob_batch(x, zmask=zmask)
        assert np.any(np.abs(plain[~bad] - got[~bad]) > 1e-3)  # the term is there
        for i in (0, 1, len(x) - 1):
            s = like.log_prob(x[i].copy(), ignore_log_det_cov=False, zmask=zmask)
            assert s == pytest.approx(got[i], rel=1e-14)
            r = like.log_prob_and_blobs(x[i].copy(), ignore_log_det_cov=False, zmask=zmask)
            assert r[0] == pytest.approx(got[i], rel=1e-14)
    like.close()


@pytest.mark.parametrize("n_hcd", [1, 2])
def test_hcd_mcdonald2005_on_device(n_hcd):
    """HCD_Model_McDonald2005 (hcd_model_McDonald2005.py:71-92; formula pinned to the reference class in
    tests/golden/mcdonald_kat.npz): the CUDA path against the oracle in every form of stage 3, walkers on both sides
    of the null value of the constant coefficient"""
    import os

    from cup1d_b200 import synth
    from oracle.model import OracleLikelihood

    emu = synth.synth_nn_emulator(seed=5, nhidden=2, max_neurons=16, head=8)
    spec = synth.build_spec(emu, grid="dr1", rebin_k=4, full_cov=True, nz=5, nz_igm=2, seed=9)
    synth.use_mcdonald_hcd(spec, n_hcd=n_hcd)
    ora = OracleLikelihood(spec)
    synth.attach_mock_data(spec, np.concatenate(ora.get_p1d_kms(values=synth.truth_point(spec))[0]))
    ora = OracleLikelihood(spec)
    x = np.concatenate([synth.walkers_ball(spec, 40, seed=3), synth.walkers_sweep(spec, 24, seed=4, frac_out=0.1)])
    i0 = spec["params"]["names"].index("HCD_damp1_0")
    x[:10, i0] = 0.05  # ln_A_damp_0 = -6.5 <= null: HCD term = 1
    exp = ora.log_prob_and_blobs_many(x)
    ok = exp[:, 0] > -1e99
    pref = np.stack([np.concatenate(ora.get_p1d_kms(values=x[i].copy())[0]) for i in np.where(ok)[0][:16]])
    like = _like(spec)
    try:
        for layout in ("warp", "lanes", "split"):
            os.environ["C1D_P1D_LAYOUT"] = layout
            got = like.log_prob_and_blobs_vectorized(x)
            np.testing.assert_array_equal(got[~ok, 0], exp[~ok, 0])
            assert np.all(np.abs(got[ok, 0] - exp[ok, 0]) <= _tol(exp[ok, 0]))
            p = like.get_p1d_kms_batch(x[np.where(ok)[0][:16]])
            assert np.max(np.abs(p / pref - 1)) <= 1e-11
    finally:
        os.environ.pop("C1D_P1D_LAYOUT", None)
        like.close()
