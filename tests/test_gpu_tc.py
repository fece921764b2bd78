"""tcgen05 (3xTF32) emulator path: single-MMA self test, the MLP against the
fp64 CUDA-core path, and end-to-end parity of the model P1D."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tf32(x):
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return b.view(np.float32)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("N", [16, 80, 128, 256])
def test_single_mma_chain(mode, N):
    from cup1d_b200 import _abi

    lib = _abi.load_library()
    rng = np.random.default_rng(N + mode)
    A = rng.normal(0, 1, (128, 32)).astype(np.float32)
    B = rng.normal(0, 1, (N, 32)).astype(np.float32)
    D = np.zeros((128, N), dtype=np.float32)
    rc = lib.c1d_tc_selftest(0, A.ctypes.data, B.ctypes.data, D.ctypes.data, N, mode)
    assert rc == 0
    ref = _tf32(A).astype(np.float64) @ _tf32(B).astype(np.float64).T
    err = np.max(np.abs(D - ref))
    assert err < 2e-5, err


def _c4_like():
    from cup1d_b200.workloads import make_likelihood

    return make_likelihood("c4", device=0, emulator_precision="fp64")


def test_mlp_tc_vs_fp64():
    import torch

    from cup1d_b200 import _abi

    spec, like = _c4_like()
    rng = np.random.default_rng(5)
    nrows = 128 * 37 + 55  # ragged last tile
    lo, hi = np.asarray(spec["emulator"]["xmin"]), np.asarray(spec["emulator"]["xmax"])
    x = np.zeros((nrows, 8))
    x[:, : len(lo)] = lo + (hi - lo) * rng.uniform(0.1, 0.9, (nrows, len(lo)))
    xd = torch.as_tensor(x, device="cuda:0")
    c64 = torch.zeros((nrows, 8), dtype=torch.float64, device="cuda:0")
    ctc = torch.zeros((nrows, 8), dtype=torch.float64, device="cuda:0")
    like.ctx.emulator_only(xd.data_ptr(), nrows, c64.data_ptr(), 0)
    like.ctx.emulator_only(xd.data_ptr(), nrows, ctc.data_ptr(), _abi.FLAG_EMU_TC)
    torch.cuda.synchronize()
    nc = like.ctx.cfg["nc"]
    a, b = c64[:, :nc].cpu().numpy(), ctc[:, :nc].cpu().numpy()
    err = np.max(np.abs(a - b))
    print("max |coef_tc - coef_fp64| =", err, " max|coef| =", np.max(np.abs(a)))
    # coefficients of log10 P: 4e-6 absolute is 1e-5 relative in P
    assert err < 2e-6, err
    like.close()


def test_p1d_and_loglike_tc():
    from cup1d_b200 import synth

    spec, like = _c4_like()
    x = synth.walkers_sweep(spec, 4096, seed=16)
    p64 = like.get_p1d_kms_batch(x)
    l64 = like.log_prob_batch(x)
    like.set_emulator_precision("3xtf32")
    ptc = like.get_p1d_kms_batch(x)
    ltc = like.log_prob_batch(x)
    ok = ~np.isnan(p64).any(axis=1)
    assert np.array_equal(ok, ~np.isnan(ptc).any(axis=1))
    rel = np.max(np.abs(ptc[ok] / p64[ok] - 1))
    dl = np.abs(ltc[ok] - l64[ok])
    print("3xTF32 emulator: max rel dP = %.3g, median |dlnL| = %.3g, max |dlnL| = %.3g" % (rel, np.median(dl), dl.max()))
    assert rel <= 1e-5
    assert np.array_equal(ltc[~ok], l64[~ok])
    like.close()
