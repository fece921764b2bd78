"""Run the *reference* cup1d (from /root/reference, read-only, or from its vendored copy baseline/_ref) so
that its own outputs can pin the oracle and so that `bench.py --impl reference` can time
the reference's own per-walker path on the GPU box's host cores.

This module only works where the reference is present; nothing under tests -m gpu or smoke() imports it.  It

1. stubs the packages cup1d imports at module level but that are not
   installed here (lace, mpi4py, matplotlib),
2. writes synthetic LaCE-schema data files into a
   scratch directory,
3. provides an analytic flat-LCDM ``CAMBModel`` stand-in (CAMB is absent), and
4. builds a reference ``Likelihood`` for a named configuration, with one of
   the oracle's emulator restatements injected as ``theory.emulator``.
"""

import os
import sys
import tempfile
import types
from types import SimpleNamespace

import numpy as np

# the reference itself (build container) or its vendored copy (tools/vendor_reference.py -> baseline/_ref, which
# travels to the GPU box); C1D_REF_ROOT overrides
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.environ.get("C1D_REF_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/cup1d") else os.path.join(_REPO, "baseline", "_ref"))
_STATE = {}


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "cup1d"))


# ---------------------------------------------------------------------------
# synthetic background + histories (shared with nothing: reference-side only)
# ---------------------------------------------------------------------------
def hubble_over_1pz(z, H0=67.66, Om=0.3111):
    z = np.asarray(z, dtype=float)
    return H0 * np.sqrt(Om * (1 + z) ** 3 + 1 - Om) / (1 + z)


def fid_linP(z):
    z = np.asarray(z, dtype=float)
    return {
        "Delta2_p": 0.35 * (4.0 / (1 + z)) ** 2,
        "n_p": -2.30 + 0.0 * z,
        "alpha_p": -0.215 + 0.0 * z,
    }


FID_STAR = {
    "Delta2_star": 0.35,
    "n_star": -2.30,
    "alpha_star": -0.215,
    "f_star": 0.981,
    "g_star": 0.968,
}


def _igm_history(z, rng, amp):
    """One simulation's IGM history; amp=0 gives the central simulation"""
    M = hubble_over_1pz(z)
    e = rng.normal(0, 1, 8) * amp
    tau = 0.0025 * (1 + z) ** 3.7 * np.exp(0.25 * e[0] + 0.1 * e[1] * (z - 3))
    sigT = 9.1 * np.sqrt(1.1 + 0.15 * (z - 3) ** 2 * 0 - 0.08 * (z - 3)) * (1 + 0.2 * e[2] + 0.05 * e[3] * (z - 3))
    gamma = (1.45 + 0.05 * (z - 3)) * (1 + 0.15 * e[4] + 0.03 * e[5] * (z - 3))
    kF = (0.105 + 0.012 * (z - 3)) * (1 + 0.15 * e[6] + 0.03 * e[7] * (z - 3))
    return {
        "z": z.copy(),
        "tau_eff": tau,
        "mF": np.exp(-tau),
        "sigT_kms": sigT,
        "sigT_Mpc": sigT / M,
        "gamma": gamma,
        "kF_kms": kF,
        "kF_Mpc": kF * M,
    }


def write_fake_lace(root, nyx_root, labels, seed=1234):
    """LaCE-schema files"""
    rng = np.random.default_rng(seed)
    z = np.arange(2.0, 4.5 + 1e-6, 0.25)
    d_aus = os.path.join(root, "data", "sim_suites", "Australia20")
    d_cov = os.path.join(root, "data", "covariance")
    os.makedirs(d_aus, exist_ok=True)
    os.makedirs(d_cov, exist_ok=True)
    os.makedirs(nyx_root, exist_ok=True)
    for suite, folder, cosmo_name in (
        ("mpg", d_aus, "mpg_emu_cosmo.npy"),
        ("nyx", nyx_root, "nyx_emu_cosmo_models_Nyx_Mar2025_with_CGAN_val_3axes.npy"),
    ):
        igm = {suite + "_central": _igm_history(z, rng, 0.0)}
        cosmo = {}
        lp = fid_linP(z)
        cosmo[suite + "_central"] = {
            "sim_label": suite + "_central",
            "cosmo_params": {},
            "linP_params": {k: v.copy() for k, v in lp.items()},
            "star_params": dict(FID_STAR),
        }
        for i in range(30):
            lab = "%s_%d" % (suite, i)
            igm[lab] = _igm_history(z, rng, 1.0)
            dA, dn, da = rng.normal(0, 1, 3) * np.array([0.15, 0.05, 0.02])
            cosmo[lab] = {
                "sim_label": lab,
                "cosmo_params": {},
                "linP_params": {
                    "Delta2_p": lp["Delta2_p"] * np.exp(dA),
                    "n_p": lp["n_p"] + dn,
                    "alpha_p": lp["alpha_p"] + da,
                },
                "star_params": {
                    "Delta2_star": FID_STAR["Delta2_star"] * np.exp(dA),
                    "n_star": FID_STAR["n_star"] + dn,
                    "alpha_star": FID_STAR["alpha_star"] + da,
                },
            }
        np.save(os.path.join(folder, "IGM_histories.npy"), igm, allow_pickle=True)
        np.save(os.path.join(folder, cosmo_name), cosmo, allow_pickle=True)
    # emulator leave-one-out covariance (relative), stacked over (z, k)
    for label in labels:
        zz = np.arange(2.0, 4.5 + 1e-6, 0.5)
        kk = np.geomspace(0.05, 5.0, 12)
        zz_zk = np.repeat(zz, len(kk))
        k_zk = np.tile(kk, len(zz))
        m = len(zz_zk)
        a = rng.normal(0, 1, (m, 6)) * 0.004
        cov = a @ a.T + np.diag(np.full(m, 0.003**2))
        np.save(
            os.path.join(d_cov, "l1O_cov_" + label + ".npy"),
            {"zz_zk": zz_zk, "k_Mpc_zk": k_zk, "cov_zk": cov, "zz": zz, "k_Mpc": kk},
            allow_pickle=True,
        )


# ---------------------------------------------------------------------------
# module stubs
# ---------------------------------------------------------------------------
def _install_stubs(lace_root):
    if "cup1d" in sys.modules and _STATE.get("stubs"):
        return

    try:
        import torch  # noqa: F401 - before the stand-ins below: torch's own optional imports must not see them
    except Exception:  # noqa: BLE001
        pass

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Any(object):
        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    mpl = mod("matplotlib", rcParams={}, use=lambda *a, **k: None)
    mpl.pyplot = mod("matplotlib.pyplot")
    mpl.pyplot.__getattr__ = lambda k: _Any()
    mpl.ticker = mod("matplotlib.ticker", MaxNLocator=_Any())
    mpl.colors = mod("matplotlib.colors")
    mpl.colors.__getattr__ = lambda k: _Any()
    mpl.cm = mod("matplotlib.cm")
    mpl.cm.__getattr__ = lambda k: _Any()

    comm = SimpleNamespace(Get_rank=lambda: 0, Get_size=lambda: 1)
    mpi4py = mod("mpi4py")
    mpi4py.MPI = mod("mpi4py.MPI", COMM_WORLD=comm)

    lace = mod("lace")
    lace.__path__ = [os.path.join(lace_root, "lace")]
    lace.cosmo = mod("lace.cosmo")
    lace.cosmo.camb_cosmo = mod("lace.cosmo.camb_cosmo", get_mnu=lambda cosmo: 0.0)
    lace.cosmo.fit_linP = mod("lace.cosmo.fit_linP")
    lace.cosmo.thermal_broadening = mod(
        "lace.cosmo.thermal_broadening",
        thermal_broadening_kms=lambda T: 9.1 * np.sqrt(np.asarray(T) / 1e4),
        T0_from_broadening_kms=lambda s: 1e4 * (np.asarray(s) / 9.1) ** 2,
    )
    lace.emulator = mod("lace.emulator")
    lace.emulator.emulator_manager = mod("lace.emulator.emulator_manager")

    # the fitter's imports (fitter.py:3-8): emcee -> this repo's emcee-compatible ensemble sampler, pyDOE2.lhs -> a
    # plain Latin hypercube (neither package is installed; run_minimizer itself only needs scipy)
    try:
        import emcee  # noqa: F401
    except Exception:  # noqa: BLE001
        from cup1d_b200 import sampler as _smp

        mod("emcee", EnsembleSampler=_smp.EnsembleSampler)
    try:
        import pyDOE2  # noqa: F401
    except Exception:  # noqa: BLE001

        def lhs(n, samples=None, criterion=None, random_state=None):
            rng = np.random.default_rng(random_state)
            samples = samples or n
            u = (rng.permuted(np.tile(np.arange(samples), (n, 1)), axis=1).T + rng.uniform(size=(samples, n))) / samples
            return u

        mod("pyDOE2", lhs=lhs)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _STATE["stubs"] = True


LABELS = ("synthetic_mpg_nn", "synthetic_mpg_gp", "synthetic_nyx_nn", "synthetic_nyx_gp")


def setup(labels=LABELS):
    """Idempotent: stubs + scratch LaCE tree.  Returns the scratch root."""
    if "root" in _STATE:
        return _STATE["root"]
    root = tempfile.mkdtemp(prefix="fake_lace_")
    nyx = os.path.join(root, "nyx")
    write_fake_lace(root, nyx, labels)
    os.environ["NYX_PATH"] = nyx
    _install_stubs(root)
    _STATE["root"] = root
    return root


# ---------------------------------------------------------------------------
# analytic CAMBModel stand-in
# ---------------------------------------------------------------------------
def make_fid_cosmo(zs, z_star=3.0, kp_kms=0.009, As=2.1e-9, ns=0.9665, nrun=0.0):
    from cup1d.likelihood.CAMB_model import CAMBModel

    class AnalyticCAMBModel(CAMBModel):
        def dkms_dMpc(self, z):
            return float(hubble_over_1pz(z))

        def get_linP_params(self):
            return dict(FID_STAR)

        def get_linP_Mpc_params(self, kp_Mpc):
            lp = fid_linP(np.asarray(self.zs))
            return [
                {k: float(lp[k][i]) for k in ("Delta2_p", "n_p", "alpha_p")}
                for i in range(len(self.zs))
            ]

    cosmo = SimpleNamespace(
        ombh2=0.02242,
        omch2=0.11933,
        H0=67.66,
        omnuh2=0.0,
        omk=0.0,
        InitPower=SimpleNamespace(As=As, ns=ns, nrun=nrun, pivot_scalar=0.05),
    )
    return AnalyticCAMBModel(zs=zs, cosmo=cosmo, z_star=z_star, kp_kms=kp_kms)


# ---------------------------------------------------------------------------
# reference Likelihood for a configuration
# ---------------------------------------------------------------------------
def build_reference_likelihood(
    emulator,
    data,
    fit_type="global_opt",
    name_variation=None,
    rebin_k=8,
    emu_cov_type="block",
    use_hull=False,
    use_star_priors=None,
    vary_alphas=False,
    prior_Gauss_rms=None,
    Gauss_priors=None,
    use_ic=True,
    ic_correction=None,
    args_hook=None,
):
    """Follow Pipeline.__init__ (pipeline.py:51-160) with hand-made data and
    a hand-made fiducial cosmology."""
    setup()
    from cup1d.likelihood.input_pipeline import Args
    from cup1d.likelihood.likelihood import Likelihood
    from cup1d.likelihood.lya_theory import Theory
    from cup1d.likelihood.model_contaminants import Contaminants
    from cup1d.likelihood.model_igm import IGM
    from cup1d.likelihood.model_systematics import Systematics
    from cup1d.pipeline.set_like_params import set_free_like_parameters
    from cup1d.utils.hull import Hull

    Args.check_emulator_label = lambda self: None
    args = Args(emulator_label=emulator.emulator_label, rebin_k=rebin_k)
    args.emu_cov_type = "full"  # so that set_baseline accepts name_variation
    args.set_baseline(
        fit_type=fit_type,
        fix_cosmo=False,
        P1D_type="DESIY1_QMLE3",
        name_variation=name_variation,
    )
    args.emu_cov_type = emu_cov_type
    args.vary_alphas = vary_alphas
    args.use_star_priors = use_star_priors
    args.prior_Gauss_rms = prior_Gauss_rms
    if ic_correction is not None:
        args.ic_correction = ic_correction
    if Gauss_priors is not None:
        args.fid_igm["Gauss_priors"] = Gauss_priors
        args.fid_cont["Gauss_priors"] = Gauss_priors
        args.fid_syst["Gauss_priors"] = Gauss_priors
    if not use_ic:
        args.file_ic = None
    if args_hook is not None:
        args_hook(args)
    free = set_free_like_parameters(args, emulator.emulator_label)

    theory = Theory(
        emulator=emulator,
        model_igm=IGM(free_param_names=free, pars_igm=args.fid_igm),
        model_cont=Contaminants(
            free_param_names=free, pars_cont=args.fid_cont, ic_correction=args.ic_correction
        ),
        model_syst=Systematics(free_param_names=free, pars_syst=args.fid_syst),
        use_hull=use_hull,
        use_star_priors=use_star_priors,
        z_star=args.z_star,
        kp_kms=args.kp_kms,
    )
    # Theory.set_fid_cosmo (lya_theory.py:82-136) with the analytic CAMB stand-in
    zs = np.asarray(data.z, dtype=float)
    theory.zs = zs
    theory.zs_hires = None
    suite = emulator.list_sim_cube[0][:3]
    if use_hull:
        theory.hull = Hull(zs=zs, data_hull=theory.hc_points, suite=suite, extra_factor=1.15)
    _zs = np.unique(np.concatenate([zs, [theory.z_star]]))
    cm = make_fid_cosmo(_zs, z_star=theory.z_star, kp_kms=theory.kp_kms)
    theory.fid_cosmo = {
        "zs": _zs,
        "cosmo": cm,
        "linP_Mpc_params": cm.get_linP_Mpc_params(kp_Mpc=theory.emu_kp_Mpc),
        "M_of_zs": cm.get_M_of_zs(),
        "linP_params": cm.get_linP_params(),
    }
    theory.set_cosmo_priors()
    theory.model_igm.set_fid_igm(zs)

    like = Likelihood(
        data,
        theory,
        free_param_names=free,
        cov_factor=args.cov_factor,
        emu_cov_type=emu_cov_type,
        prior_Gauss_rms=prior_Gauss_rms,
        args=args,
    )
    return like, args


def make_data_namespace(zs, k_kms, k_min, k_max, Pk, cov, cov_stat, full_cov=None, full_cov_stat=None):
    """The output contract of cup1d/p1ds (base_p1d_data.py:90-168) as a
    duck-typed object."""
    d = SimpleNamespace(
        z=np.asarray(zs, dtype=float),
        k_kms=[np.asarray(k) for k in k_kms],
        k_kms_min=[np.asarray(k) for k in k_min],
        k_kms_max=[np.asarray(k) for k in k_max],
        Pk_kms=[np.asarray(p) for p in Pk],
        Pksmooth_kms=None,
        cov_Pk_kms=[np.asarray(c) for c in cov],
        covstat_Pk_kms=[np.asarray(c) for c in cov_stat],
        full_Pk_kms=None,
        full_cov_Pk_kms=None,
        full_cov_stat_Pk_kms=None,
        full_zs=None,
        full_k_kms=None,
        apply_blinding=False,
    )
    if full_cov is not None:
        d.full_Pk_kms = np.concatenate(d.Pk_kms)
        d.full_cov_Pk_kms = np.asarray(full_cov)
        d.full_cov_stat_Pk_kms = np.asarray(full_cov_stat)
        d.full_zs = np.concatenate([np.full(len(k), z) for z, k in zip(d.z, d.k_kms)])
        d.full_k_kms = np.concatenate(d.k_kms)
    return d
