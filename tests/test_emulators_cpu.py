"""The restated LaCE emulators (oracle/emulators.py) against independent third-party implementations
of the same published algorithms: torch.nn (Linear + LeakyReLU, what LaCE's NNEmulator is built from)
and scikit-learn's GaussianProcessRegressor (ARD squared-exponential posterior mean).  LaCE itself is
absent (stage 2 stays parity-unpinned against it); this pins the arithmetic of the restatement."""
import numpy as np
import torch

from cup1d_b200 import synth
from oracle.emulators import RefGPEmulator, RefMLPEmulator


def _inputs(spec, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(spec["xmin"]), np.asarray(spec["xmax"])
    return lo + rng.uniform(0.02, 0.98, (n, len(lo))) * (hi - lo)


def test_mlp_restatement_matches_torch_nn():
    for suite, ndeg in (("mpg", 5), ("nyx", 6)):
        spec = synth.synth_nn_emulator(seed=3, suite=suite, ndeg=ndeg)
        emu = RefMLPEmulator(spec)
        mods = []
        for il, layer in enumerate(spec["layers"]):
            W, b = np.asarray(layer["W"], dtype=np.float64), np.asarray(layer["b"], dtype=np.float64)
            lin = torch.nn.Linear(W.shape[1], W.shape[0]).double()
            with torch.no_grad():
                lin.weight.copy_(torch.from_numpy(W))
                lin.bias.copy_(torch.from_numpy(b))
            mods.append(lin)
            if il < len(spec["layers"]) - 1:
                mods.append(torch.nn.LeakyReLU(float(spec["slope"])))
        net = torch.nn.Sequential(*mods)
        x = _inputs(spec, 257, 5)
        xn = (x - np.asarray(spec["xmin"])) / (np.asarray(spec["xmax"]) - np.asarray(spec["xmin"])) - 0.5
        with torch.no_grad():
            ref64 = net(torch.from_numpy(xn)).numpy()
            ref32 = net.float()(torch.from_numpy(xn).float()).double().numpy()
        got = emu.coefficients(x)
        np.testing.assert_allclose(got, ref64, rtol=1e-12, atol=1e-13)
        # LaCE evaluates in float32: the float64 restatement sits within fp32 round-off of it
        assert np.max(np.abs(got - ref32)) < 2e-5


def test_gp_restatement_matches_sklearn():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    for suite, n_out in (("mpg", 5), ("nyx", 6)):
        spec = synth.synth_gp_emulator(seed=8, suite=suite, n_train=120, n_out=n_out)
        emu = RefGPEmulator(spec)
        X, ell, s2 = np.asarray(spec["X_train"]), np.asarray(spec["lengthscales"]), float(spec["sigma2"])
        d = (X[:, None, :] - X[None, :, :]) / ell
        K = s2 * np.exp(-0.5 * np.sum(d * d, axis=2)) + 1e-6 * np.eye(len(X))
        y = K @ np.asarray(spec["alpha"])  # the training targets alpha = K^-1 y was solved from
        gpr = GaussianProcessRegressor(kernel=ConstantKernel(s2, "fixed") * RBF(ell, "fixed"), alpha=1e-6, optimizer=None).fit(X, y)
        x = _inputs(spec, 64, 9)
        u = (x - np.asarray(spec["xmin"])) / (np.asarray(spec["xmax"]) - np.asarray(spec["xmin"]))
        m = gpr.predict(u)
        want = np.asarray(spec["ymean"]) + np.asarray(spec["ystd"]) * m
        mF = x[:, emu.imF]
        lnn = np.polyval(np.asarray(spec["norm_coef"])[::-1], mF)
        want[:, 0] += lnn
        got = emu.coefficients(x)
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-7 * np.max(np.abs(want)))


import pytest  # noqa: E402


@pytest.mark.parametrize("per_output", [False, True])
def test_sklearn_gp_emulator_adapter(per_output):
    """emulator_to_spec on an object shaped like LaCE's default GP emulator (tests/fake_lace.py: a fitted scikit-learn
    GaussianProcessRegressor behind the attributes the reference reads): the extracted spec, evaluated by the oracle
    restatement, reproduces the object's own emulate_p1d_Mpc, i.e. sklearn.predict"""
    from fake_lace import FakeLaceGP
    from cup1d_b200.spec import emulator_to_spec
    from oracle.emulators import RefGPEmulator

    emu = FakeLaceGP(seed=4, per_output=per_output)
    spec = emulator_to_spec(emu)
    assert spec["kind"] == "gp" and spec["norm_table"].shape == (19, 40)
    ref = RefGPEmulator(spec)
    rng = np.random.default_rng(1)
    lims = emu.param_lims
    for _ in range(5):
        X = lims[:, 0] + (lims[:, 1] - lims[:, 0]) * rng.uniform(0.1, 0.9, (11, len(lims)))
        call = {p: X[:, j] for j, p in enumerate(emu.emu_params)}
        k = np.zeros((11, 40))
        for iz in range(11):
            n = 25 + iz
            k[iz, :n] = np.geomspace(0.015, 5.0, n)  # beyond both ends of the norm's k grid
        a, b = emu.emulate_p1d_Mpc(call, k), ref.emulate_p1d_Mpc(call, k)
        good = k > 0
        assert np.max(np.abs(b[good] / a[good] - 1)) < 1e-9


def test_sklearn_gp_adapter_rejects_what_it_cannot_represent():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import Matern

    from fake_lace import FakeLaceGP
    from cup1d_b200.spec import emulator_to_spec

    emu = FakeLaceGP(seed=4, optimise=False)
    U = np.asarray(emu.gp.X_train_)
    emu.gp = GaussianProcessRegressor(kernel=Matern(length_scale=np.ones(U.shape[1])), optimizer=None).fit(U, np.zeros((len(U), 5)))
    with pytest.raises(NotImplementedError, match="unsupported GP kernel"):
        emulator_to_spec(emu)
    emu = FakeLaceGP(seed=4, optimise=False)
    emu.func_poly = lambda x, *c: np.exp(np.polyval(c, x))
    with pytest.raises(NotImplementedError, match="func_poly"):
        emulator_to_spec(emu)
