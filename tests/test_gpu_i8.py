"""int8 digit-slice emulator (tcgen05 kind::i8, cup1d_b200/csrc/c1d_mlp_i8.cu): the kernel against its
NumPy mirror (tests/i8sim.py, exact integer sums) on network shapes that exercise every block plan, against
the float64 kernel, and end to end on the DESI-DR1-like workload at the north-star tolerances
(model P1D 1e-5 relative, log-likelihood 1e-4 absolute)."""

import numpy as np
import pytest

import i8sim

pytestmark = pytest.mark.gpu

ATOL_LNL = 1e-4
RTOL_P1D = 1e-5


def _custom_nn(dims, seed=3, slope=0.5):
    """an NN emulator spec with the given layer widths (dims[0] = 6 inputs, dims[-1] coefficients)"""
    from cup1d_b200 import synth

    base = synth.synth_nn_emulator(seed=seed, ndeg=dims[-1] - 1)
    rng = np.random.default_rng(seed)
    layers = []
    for il in range(len(dims) - 1):
        fan_in, fan_out = dims[il], dims[il + 1]
        W = rng.normal(0, 1.2 / np.sqrt(fan_in), (fan_out, fan_in))
        # rows of very different size and a few exact zeros: per-row weight scales and empty digits
        W *= 10.0 ** rng.uniform(-2, 0.3, (fan_out, 1))
        W[rng.uniform(size=W.shape) < 0.05] = 0.0
        b = rng.normal(0, 0.3, fan_out)
        layers.append({"W": W, "b": b})
    base["layers"] = layers
    base["slope"] = slope
    return base


def _context(emu):
    from cup1d_b200 import synth
    from cup1d_b200.likelihood import B200Likelihood

    spec = synth.build_spec(emu, grid="chab", rebin_k=1, full_cov=False, nz=2, seed=5)
    return spec, B200Likelihood(spec, device=0, emulator_precision="fp64")


def _coefficients(like, x, flags):
    import torch

    nrows = x.shape[0]
    xin = np.zeros((nrows, 8))
    xin[:, : x.shape[1]] = x
    xd = torch.as_tensor(xin, device="cuda:0")
    out = torch.zeros((nrows, 8), dtype=torch.float64, device="cuda:0")
    like.ctx.emulator_only(xd.data_ptr(), nrows, out.data_ptr(), flags)
    torch.cuda.synchronize()
    return out[:, : like.ctx.cfg["nc"]].cpu().numpy()


SHAPES = [
    [6, 6],                      # output layer alone: input quantisation, one block
    [6, 10, 6],                  # two layers, 32-column block
    [6, 100, 5],                 # 128-column block
    [6, 150, 6],                 # blocks (128, 32) sharing a scale, K = 160 afterwards
    [6, 190, 7],                 # blocks (128, 64), K = 192
    [6, 250, 50, 6],             # blocks (128, 128) with a scale each, then two K-groups
    [6, 40, 230, 64, 30, 6],     # K-groups feeding a 64-wide layer, two more layers behind it
    [6, 10, 70, 130, 190, 250, 50, 6],  # Cabayol23+-shaped
]


@pytest.mark.parametrize("dims", SHAPES, ids=lambda d: "-".join(map(str, d)))
def test_kernel_equals_mirror(dims):
    from cup1d_b200 import _abi

    emu = _custom_nn(dims, seed=len(dims) + dims[1])
    assert i8sim.supported(dims)
    spec, like = _context(emu)
    assert like.ctx.emulator_paths() & _abi.FLAG_EMU_I8
    rng = np.random.default_rng(11)
    nrows = 128 * 5 + 77  # ragged last tile
    lo, hi = np.asarray(emu["xmin"]), np.asarray(emu["xmax"])
    x = lo + (hi - lo) * rng.uniform(-0.2, 1.2, (nrows, len(lo)))
    x[3] = lo  # normalised inputs of exactly -0.5
    x[5, 2] = (lo[2] + hi[2]) / 2
    dev = _coefficients(like, x, _abi.FLAG_EMU_I8)
    f64 = _coefficients(like, x, 0)
    mir = i8sim.mirror_coefficients(emu, x)
    scale = np.max(np.abs(f64)) + 1.0
    r_mir = np.max(np.abs(dev - mir), axis=1) / scale
    e_f64 = np.max(np.abs(dev - f64)) / scale
    print("dims %s: |dev - mirror| median %.2e max %.2e  |dev - fp64| %.2e (scale %.3g)" % (dims, np.median(r_mir), r_mir.max(), e_f64, scale))
    # integer sums are exact on both sides; a last-bit difference of the float64 steps can move one activation
    # across a rounding boundary of its quantisation (one digit unit, ~5e-10 of the row maximum) in rare rows
    assert np.quantile(r_mir, 0.98) < 1e-12, np.quantile(r_mir, 0.98)
    assert r_mir.max() < 1e-7, r_mir.max()
    assert e_f64 < 1e-6, e_f64
    like.close()


def test_unsupported_shapes_keep_fp64():
    from cup1d_b200 import _abi
    from cup1d_b200.likelihood import B200Likelihood

    for dims in ([6, 250, 250, 6], [6, 250, 100, 6]):
        assert not i8sim.supported(dims)
        spec, like = _context(_custom_nn(dims))
        assert not (like.ctx.emulator_paths() & _abi.FLAG_EMU_I8)
        with pytest.raises(ValueError):
            like.set_emulator_precision("i8")
        auto = B200Likelihood(spec, device=0)
        assert auto.emulator_precision == "fp64"
        auto.close()
        like.close()


def test_nonfinite_inputs_propagate():
    from cup1d_b200 import _abi

    emu = _custom_nn([6, 10, 70, 50, 6])
    spec, like = _context(emu)
    lo, hi = np.asarray(emu["xmin"]), np.asarray(emu["xmax"])
    x = np.tile((lo + hi) / 2, (300, 1))
    x[7, 1] = np.nan
    x[130, 0] = np.inf
    dev = _coefficients(like, x, _abi.FLAG_EMU_I8)
    assert np.isnan(dev[7]).all() and np.isnan(dev[130]).all()
    ok = np.ones(300, bool)
    ok[[7, 130]] = False
    assert np.isfinite(dev[ok]).all()
    like.close()


def test_c4_p1d_and_loglike():
    """the tolerance gate of the north star on the configuration the metric is quoted on"""
    from cup1d_b200 import synth
    from cup1d_b200.workloads import make_likelihood

    spec, like = make_likelihood("c4", device=0, emulator_precision="fp64")
    x = synth.walkers_sweep(spec, 8192, seed=16)
    p64 = like.get_p1d_kms_batch(x)
    l64 = like.log_prob_batch(x)
    like.set_emulator_precision("i8")
    p8 = like.get_p1d_kms_batch(x)
    l8 = like.log_prob_batch(x)
    ok = ~np.isnan(p64).any(axis=1)
    assert np.array_equal(ok, ~np.isnan(p8).any(axis=1))
    rel = np.max(np.abs(p8[ok] / p64[ok] - 1))
    dl = np.abs(l8[ok] - l64[ok])
    print("int8 digit slices: max rel dP = %.3g, |dlnL| median %.3g  p99 %.3g  max %.3g  (%d walkers)"
          % (rel, np.median(dl), np.quantile(dl, 0.99), dl.max(), ok.sum()))
    assert rel <= RTOL_P1D
    # 1e-4 absolute wherever a sampler goes (|lnL| <= 1e5); beyond, the quantisation error of the model (3e-9
    # relative) times the pull of the residuals grows like sqrt(chi2): 1e-9 relative
    assert np.all(dl <= ATOL_LNL + 1e-9 * np.abs(l64[ok])), (dl / (ATOL_LNL + 1e-9 * np.abs(l64[ok]))).max()
    assert dl[np.abs(l64[ok]) <= 1e5].max() <= ATOL_LNL
    assert np.array_equal(l8[~ok], l64[~ok])
    like.close()


def test_auto_is_i8_on_c4():
    from cup1d_b200.workloads import make_likelihood

    spec, like = make_likelihood("c4", device=0)
    assert like.emulator_precision == "i8"
    like.close()
