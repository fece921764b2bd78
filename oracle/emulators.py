"""TEST INFRASTRUCTURE — CPU restatement of the LaCE emulators (stage 2).

PARITY UNPINNED: LaCE (https://github.com/igmhub/LaCE, no pinned version; not
listed in the reference's pyproject.toml:23-40, installed by ``git clone`` per
README.md:23-30) is absent from /root/reference and cannot be installed here.
No reference test calls ``emulate_p1d_Mpc``.  These classes restate the
*published / recalled* algorithms in NumPy float64 and follow
the call contract that *is* verifiable in the reference:

* call site        cup1d/likelihood/lya_theory.py:672-701
  ``emulator.emulate_p1d_Mpc(emu_call, kin_Mpc)`` with ``emu_call`` a dict
  name -> float64[Nz] (plus ``mF_fid``) and ``kin_Mpc`` a zero-padded
  float64[Nz, Lmax]; returns [Nz, Lmax].
* attributes       ``emu_params`` (lya_theory.py:447), ``kp_Mpc`` (:57),
  ``list_sim_cube`` (:58), ``emulator_label`` (likelihood.py:276),
  ``kmax_Mpc`` (fitter.py:972).

The same object is injected into the *reference* ``Theory`` when golden
vectors are generated (tools/make_golden.py), so parity of the whole path is
measured against the reference's own code for stages 1, 3, 4, 5.

The arithmetic of the restatement is cross-checked against independent implementations of the
same algorithms (tests/test_emulators_cpu.py): torch.nn Linear + LeakyReLU stacks in float64 and
float32 (LaCE's own precision) and scikit-learn's GaussianProcessRegressor with the same fixed
ARD kernel.  That pins the formulas, not LaCE's trained models.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.
"""

import numpy as np


class RefMLPEmulator(object):
    """LaCE ``NNEmulator`` restated [recalled: lace/emulator/nn_emulator.py,
    nn_architecture.py::MDNemulator_polyfit].

    x_norm = (x - lo)/(hi - lo) - 0.5; a chain of Linear + LeakyReLU(slope)
    layers, last Linear without activation giving ndeg+1 coefficients c_j;
    P1D_Mpc(k) = 10**(sum_j c_j log10(k)^j) * yscale.
    """

    def __init__(self, spec, emulator_label="synthetic_nn", suite="mpg"):
        assert spec["kind"] == "nn"
        self.spec = spec
        self.emu_params = list(spec["emu_params"])
        self.kp_Mpc = float(spec["kp_Mpc"])
        self.kmax_Mpc = float(spec["kmax_Mpc"])
        self.ndeg = int(spec["ndeg"])
        self.emulator_label = emulator_label
        self.list_sim_cube = [suite + "_0"]
        self.layers = [
            (np.asarray(l["W"], dtype=np.float64), np.asarray(l["b"], dtype=np.float64))
            for l in spec["layers"]
        ]
        self.slope = float(spec["slope"])
        self.xmin = np.asarray(spec["xmin"], dtype=np.float64)
        self.xmax = np.asarray(spec["xmax"], dtype=np.float64)
        self.yscale = float(spec["yscale"])

    def to_spec(self):
        return self.spec

    def coefficients(self, x):
        """x: [N, n_in] raw emulator inputs -> [N, ndeg+1] (float64)"""
        h = (x - self.xmin) / (self.xmax - self.xmin) - 0.5
        nl = len(self.layers)
        for il, (W, b) in enumerate(self.layers):
            h = h @ W.T + b
            if il < nl - 1:
                h = np.where(h > 0, h, self.slope * h)
        return h

    def emulate_p1d_Mpc(self, emu_call, k_Mpc):
        x = np.stack([np.asarray(emu_call[p], dtype=np.float64) for p in self.emu_params], axis=1)
        k_Mpc = np.atleast_2d(np.asarray(k_Mpc, dtype=np.float64))
        c = self.coefficients(x)
        out = np.zeros_like(k_Mpc)
        for iz in range(k_Mpc.shape[0]):
            good = k_Mpc[iz] > 0  # padded entries (lya_theory.py:685-687)
            lk = np.log10(k_Mpc[iz][good])
            acc = np.zeros_like(lk)
            for j in range(self.ndeg, -1, -1):
                acc = acc * lk + c[iz, j]
            out[iz][good] = 10.0**acc * self.yscale
        return out


class RefGPEmulator(object):
    """LaCE ``GPEmulator`` (CH24_*_gpr) restated [recalled:
    lace/emulator/gp_emulator_multi.py; smooth model visible in the
    reference at notebooks/figures_CM26/Fig3a.py:75-80].

    Inputs are scaled to the unit box, u = (x - lo)/(hi - lo).  The posterior
    mean of output o is m_o = sum_i k(u, U_i) alpha[i, o] with an ARD squared
    exponential kernel k = sigma2 exp(-0.5 sum_d ((u_d - U_id)/ell_d)^2);
    y_o = ymean_o + ystd_o m_o are the coefficients of a polynomial in
    k/kmax, and P1D_Mpc(k) = norm(mF, k) * exp(sum_o y_o (k/kmax)^o).

    The norm follows every use the reference makes of it
    (notebooks/figures_CM26/Fig3a.py:83-85, Fig4b.py:88, FigB1a.py:84):
        norm = np.interp(k_Mpc, emulator.input_norm["k_Mpc"], emulator.norm_imF(mF))
    i.e. ``norm_imF(mF)`` is a table over ``input_norm["k_Mpc"]`` for the given mean flux, interpolated
    linearly in k.  Here ``norm_imF`` interpolates a stored table ``norm_table[mF node, k node]`` linearly
    in mF (clamped outside the nodes) [the interpolation LaCE uses along mF is not visible in the
    reference; an adapter tabulates the real ``norm_imF`` on its own nodes].  Specs written before this
    form (``norm_coef``: ln norm a polynomial in mF, k-independent) are still evaluated.
    """

    def __init__(self, spec, emulator_label="synthetic_gp", suite="mpg"):
        assert spec["kind"] == "gp"
        self.spec = spec
        self.emu_params = list(spec["emu_params"])
        self.kp_Mpc = float(spec["kp_Mpc"])
        self.kmax_Mpc = float(spec["kmax_Mpc"])
        self.emulator_label = emulator_label
        self.list_sim_cube = [suite + "_0"]
        self.xmin = np.asarray(spec["xmin"], dtype=np.float64)
        self.xmax = np.asarray(spec["xmax"], dtype=np.float64)
        self.X = np.asarray(spec["X_train"], dtype=np.float64)
        self.ell = np.asarray(spec["lengthscales"], dtype=np.float64)
        self.sigma2 = float(spec["sigma2"])
        self.alpha = np.asarray(spec["alpha"], dtype=np.float64)
        self.ymean = np.asarray(spec["ymean"], dtype=np.float64)
        self.ystd = np.asarray(spec["ystd"], dtype=np.float64)
        self.norm_coef = np.asarray(spec.get("norm_coef", np.zeros(0)), dtype=np.float64)
        self.imF = self.emu_params.index("mF")
        self.has_ktable = spec.get("norm_table") is not None
        if self.has_ktable:
            self.norm_mF = np.asarray(spec["norm_mF"], dtype=np.float64)
            self.norm_table = np.asarray(spec["norm_table"], dtype=np.float64)
            # the attributes the reference reads (Fig3a.py:75-85)
            self.input_norm = {"k_Mpc": np.asarray(spec["norm_k_Mpc"], dtype=np.float64)}

    def norm_imF(self, mF):
        """norm over input_norm["k_Mpc"] for one mean flux: linear in mF between the table rows"""
        m, T = self.norm_mF, self.norm_table
        j = int(np.clip(np.searchsorted(m, mF, side="right") - 1, 0, len(m) - 2))
        w = float(np.clip((mF - m[j]) / (m[j + 1] - m[j]), 0.0, 1.0))
        return T[j] + w * (T[j + 1] - T[j])

    def func_poly(self, x, *coef):
        """sum_o coef_o x^o (the smooth model of Fig3a.py:86-88)"""
        acc = np.zeros_like(np.asarray(x, dtype=np.float64))
        for c in coef[::-1]:
            acc = acc * x + c
        return acc

    def to_spec(self):
        return self.spec

    def coefficients(self, x):
        """[N, n_in] -> polynomial coefficients [N, n_out]; a k-independent norm (the older ``norm_coef``
        form) is folded into the constant term, a k-table norm is applied in emulate_p1d_Mpc"""
        u = (x - self.xmin) / (self.xmax - self.xmin)
        d = (u[:, None, :] - self.X[None, :, :]) / self.ell
        kern = self.sigma2 * np.exp(-0.5 * np.sum(d * d, axis=2))
        y = self.ymean + self.ystd * (kern @ self.alpha)
        mF = x[:, self.imF]
        lnn = np.zeros_like(mF)
        for c in self.norm_coef[::-1]:
            lnn = lnn * mF + c
        y = y.copy()
        y[:, 0] += lnn
        return y

    def emulate_p1d_Mpc(self, emu_call, k_Mpc):
        x = np.stack([np.asarray(emu_call[p], dtype=np.float64) for p in self.emu_params], axis=1)
        k_Mpc = np.atleast_2d(np.asarray(k_Mpc, dtype=np.float64))
        y = self.coefficients(x)
        out = np.zeros_like(k_Mpc)
        for iz in range(k_Mpc.shape[0]):
            good = k_Mpc[iz] > 0
            t = k_Mpc[iz][good] / self.kmax_Mpc
            acc = np.zeros_like(t)
            for j in range(y.shape[1] - 1, -1, -1):
                acc = acc * t + y[iz, j]
            out[iz][good] = np.exp(acc)
            if self.has_ktable:
                out[iz][good] *= np.interp(k_Mpc[iz][good], self.input_norm["k_Mpc"], self.norm_imF(x[iz, self.imF]))
        return out


def emulator_from_spec(spec, **kw):
    if spec["kind"] == "nn":
        return RefMLPEmulator(spec, **kw)
    if spec["kind"] == "gp":
        return RefGPEmulator(spec, **kw)
    raise ValueError("unknown emulator kind " + str(spec["kind"]))
