"""TEST INFRASTRUCTURE - a stand-in shaped like LaCE's default GP emulator (``CH24_mpgcen_gpr``, input_pipeline.py:21)
as far as the reference shows it: ``emu_params``, ``kp_Mpc``, ``kmax_Mpc``, ``list_sim_cube``, ``emulator_label``,
``input_norm["k_Mpc"]``, ``norm_imF(mF)``, ``func_poly`` (notebooks/figures_CM26/Fig3a.py:75-88) and
``emulate_p1d_Mpc(emu_call, k_Mpc)`` (lya_theory.py:690).  Inside, a REAL scikit-learn GaussianProcessRegressor fitted
on a seeded training set predicts the polynomial coefficients - the adapter cup1d_b200.spec.emulator_to_spec must turn
this object into a device context whose model agrees with ``sklearn.predict``."""
import numpy as np
from scipy.interpolate import interp1d
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, WhiteKernel

from cup1d_b200 import synth


class FakeLaceGP(object):
    def __init__(self, seed=3, suite="mpg", n_train=90, n_out=5, per_output=False, white=True, optimise=True):
        rng = np.random.default_rng(seed)
        self.emu_params = list(synth.MPG_PARAMS if suite == "mpg" else synth.NYX_PARAMS)
        n_in = len(self.emu_params)
        self.param_lims = np.array([synth.EMU_RANGES[p] for p in self.emu_params], dtype=float)
        self.kp_Mpc, self.kmax_Mpc = 0.7, 4.0
        self.list_sim_cube = [suite + "_0"]
        self.emulator_label = "synthetic_%s_gp" % suite
        # the norm: a table over (mean flux, k) read with interp1d along mF, np.interp along k
        mF = np.linspace(0.05, 0.95, 19)
        kn = np.geomspace(0.02, 4.5, 40)
        lnn = 0.2 - 0.9 * mF[:, None] - (0.3 + 0.6 * mF[:, None]) * np.log1p(kn[None, :]) + 0.03 * np.cos(2.0 * np.log(kn[None, :]))
        self.input_norm = {"k_Mpc": kn}
        self.norm_imF = interp1d(mF, np.exp(lnn), axis=0)
        # training set on the unit box of the limits and smooth targets = polynomial coefficients
        U = rng.uniform(0.03, 0.97, (n_train, n_in))
        A = rng.normal(0, 0.5, (n_out, n_in))
        Y = np.array([-0.6, -2.8, 1.2, -0.5, 0.1, 0.02, 0.01][:n_out]) + (U - 0.5) @ A.T * np.array([0.25, 0.4, 0.3, 0.15, 0.05, 0.02, 0.01][:n_out])
        Y = Y + 0.05 * np.sin(3 * U[:, :1]) * np.arange(1, n_out + 1)
        kern = ConstantKernel(1.0, (1e-2, 1e2)) * RBF(np.full(n_in, 0.8), (0.2, 5.0))
        if white:
            kern = kern + WhiteKernel(1e-6, (1e-9, 1e-3))
        opt = "fmin_l_bfgs_b" if optimise else None
        if per_output:
            # one regressor per coefficient sharing the kernel hyper-parameters of the first fit
            g0 = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=opt, random_state=0).fit(U, Y[:, 0])
            self.gps = [g0] + [GaussianProcessRegressor(kernel=g0.kernel_, normalize_y=True, optimizer=None).fit(U, Y[:, o]) for o in range(1, n_out)]
        else:
            self.gp = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=opt, random_state=0).fit(U, Y)

    @staticmethod
    def func_poly(x, *coef):
        # highest power first (np.polyval order): the adapter has to find the order out
        return np.polyval(coef, x)

    def _coefficients(self, X):
        U = (X - self.param_lims[:, 0]) / (self.param_lims[:, 1] - self.param_lims[:, 0])
        if hasattr(self, "gp"):
            return np.atleast_2d(self.gp.predict(U))
        return np.stack([g.predict(U) for g in self.gps], axis=1)

    def emulate_p1d_Mpc(self, emu_call, k_Mpc):
        X = np.stack([np.asarray(emu_call[p], dtype=float) for p in self.emu_params], axis=1)
        k_Mpc = np.atleast_2d(np.asarray(k_Mpc, dtype=float))
        coef = self._coefficients(X)
        out = np.zeros_like(k_Mpc)
        imf = self.emu_params.index("mF")
        for iz in range(k_Mpc.shape[0]):
            good = k_Mpc[iz] > 0
            k = k_Mpc[iz][good]
            norm = np.interp(k, self.input_norm["k_Mpc"], self.norm_imF(X[iz, imf]))
            out[iz][good] = norm * np.exp(self.func_poly(k / self.kmax_Mpc, *coef[iz]))
        return out
