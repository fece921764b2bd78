"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/cup1d, imported through tools/ref_harness.py) on seeded
walker sets.  Runs only in the build container; the fixtures travel.

Each fixture holds the model spec extracted from the reference ``Likelihood``
(cup1d_b200.spec.spec_from_likelihood), the walker array and the reference's
own outputs:

  lnprob_blobs [W,7]   Likelihood.log_prob_and_blobs          (likelihood.py:1036)
  p1d          [W,N_d] concat(Likelihood.get_p1d_kms(values)) (:722) NaN if None
  loglike      [W]     get_log_like(...)[0]                    (:822)
  loglike_all  [W,Nz]  get_log_like(...)[1][0]
  chi2         [W]     get_chi2                                (:787)
  emu_call     [W,Nz,n_emu]  Theory.get_emulator_calls         (lya_theory.py:410)
  zmask, lnprob_zmask [W]   log_prob(x, zmask=zmask)           (:1026)

usage: python tools/make_golden.py [name ...]
"""

import os
import sys
import io
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from tools import ref_harness as rh  # noqa: E402
from cup1d_b200 import synth  # noqa: E402
from cup1d_b200.spec import save_spec, spec_from_likelihood  # noqa: E402
from oracle.emulators import RefGPEmulator, RefMLPEmulator  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def small_nn(seed, suite):
    return synth.synth_nn_emulator(seed=seed, suite=suite, nhidden=3, max_neurons=24, head=12)


def small_gp(seed, suite):
    return synth.synth_gp_emulator(seed=seed, suite=suite, n_train=48)


def grids(kind, nz):
    ks, kmins, kmaxs = [], [], []
    for iz in range(nz):
        if kind == "chab":
            k, a, b = synth.chabanier_k_grid(35)
        elif kind == "dr1c":  # DR1 k range, 3x coarser bins
            k, a, b = synth.dr1_k_grid(17 + iz, dk=1.5e-3)
        elif kind == "dr1":
            k, a, b = synth.dr1_k_grid(synth.DR1_NK[iz])
        ks.append(k)
        kmins.append(a)
        kmaxs.append(b)
    return ks, kmins, kmaxs


def synth_cov(p_truth, seed, snr=36.0, corr_amp=0.02):
    nd = len(p_truth)
    rng = np.random.default_rng(seed)
    A = rng.normal(0, corr_amp, (nd, nd))
    rho = np.eye(nd) + A @ A.T
    dg = np.sqrt(np.diag(rho))
    rho = rho / dg[:, None] / dg[None, :]
    D = np.abs(p_truth) / snr
    return rho * D[:, None] * D[None, :]


def walkers(like, n_ball, n_wide, n_out, seed, width=0.15):
    rng = np.random.default_rng(seed)
    x0 = np.clip(like.sampling_point_from_parameters(), 0.03, 0.97)
    ndim = len(x0)
    ball = x0[None, :] + rng.random((n_ball, ndim)) * 0.01
    ball[ball >= 1] = 0.95
    ball[ball <= 0] = 0.05
    wide = np.clip(x0[None, :] + rng.normal(0, width, (n_wide, ndim)), 0.0, 1.0)
    out = np.clip(x0[None, :] + rng.normal(0, 0.05, (n_out, ndim)), 0.01, 0.99)
    for i in range(n_out):
        out[i, rng.integers(ndim)] = -0.003 if i % 2 == 0 else 1.002
    return np.concatenate([ball, wide, out])


def run_reference(like, x, zmask):
    nz = len(like.data.z)
    nd = sum(len(k) for k in like.data.k_kms)
    W = len(x)
    names = like.theory.emulator.emu_params
    res = {
        "lnprob_blobs": np.zeros((W, 7)),
        "p1d": np.full((W, nd), np.nan),
        "loglike": np.zeros(W),
        "loglike_all": np.zeros((W, nz)),
        "chi2": np.zeros(W),
        "emu_call": np.zeros((W, nz, len(names))),
        "minus_log_prob": np.zeros(W),
        "zmask": np.asarray(zmask, dtype=float),
        "lnprob_zmask": np.zeros(W),
    }
    for i in range(W):
        v = x[i].copy()
        res["lnprob_blobs"][i] = like.log_prob_and_blobs(v.copy())
        res["minus_log_prob"][i] = like.minus_log_prob(v.copy())
        incube = not ((v > 1).any() or (v < 0).any())
        if incube:
            r = like.get_p1d_kms(values=v.copy())
            if r is not None:
                res["p1d"][i] = np.concatenate(r[0])
            lp = like.parameters_from_sampling_point(v.copy())
            call = like.theory.get_emulator_calls(like.data.z, like_params=lp, return_M_of_z=False)
            for j, nm in enumerate(names):
                res["emu_call"][i, :, j] = call[nm]
        ll = like.get_log_like(v.copy(), return_blob=True)
        res["loglike"][i] = ll[0]
        res["loglike_all"][i] = np.asarray(ll[1]).reshape(-1)[:nz] if np.ndim(ll[1]) else ll[1]
        res["chi2"][i] = like.get_chi2(v.copy())
        try:
            res["lnprob_zmask"][i] = like.log_prob(v.copy(), zmask=res["zmask"])
        except TypeError:
            # the reference itself fails for zmask + ic_correction (a list is
            # negated at model_contaminants.py:319): recorded as NaN, not tested
            res["lnprob_zmask"][i] = np.nan
    return res


def make(name, emu, grid, nz=11, rebin_k=1, full=False, noise=True, seed=100, n=(24, 20, 4), width=0.15, **kw):
    zs = np.round(2.2 + 0.2 * np.arange(nz), 10)
    ks, kmins, kmaxs = grids(grid, nz)
    nd = sum(len(k) for k in ks)

    def data_ns(Pk_full, cov_full):
        Pk, cov, off = [], [], 0
        for k in ks:
            Pk.append(Pk_full[off : off + len(k)])
            cov.append(cov_full[off : off + len(k), off : off + len(k)])
            off += len(k)
        cov_stat = [0.8 * c for c in cov]
        if full:
            return rh.make_data_namespace(zs, ks, kmins, kmaxs, Pk, cov, cov_stat, full_cov=cov_full, full_cov_stat=0.8 * cov_full)
        return rh.make_data_namespace(zs, ks, kmins, kmaxs, Pk, cov, cov_stat)

    emu_cov_type = kw.pop("emu_cov_type", "block")
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        like0, _ = rh.build_reference_likelihood(emu, data_ns(np.ones(nd), np.eye(nd) * 0.01), rebin_k=rebin_k, emu_cov_type="block", **kw)
        x_fid = np.clip(like0.sampling_point_from_parameters(), 0.03, 0.97)
        r = like0.get_p1d_kms(values=x_fid, apply_hull=False)
        p_truth = np.concatenate(r[0])
        cov = synth_cov(p_truth, seed)
        dvec = p_truth.copy()
        if noise:
            dvec = dvec + np.linalg.cholesky(cov) @ np.random.default_rng(seed + 1).normal(0, 1, nd)
        like, args = rh.build_reference_likelihood(emu, data_ns(dvec, cov), rebin_k=rebin_k, emu_cov_type=emu_cov_type, **kw)
        x = walkers(like, n[0], n[1], n[2], seed + 2, width=width)
        res = run_reference(like, x, zmask=[zs[1], zs[min(4, nz - 1)]])
    spec = spec_from_likelihood(like)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    save_spec(path, spec, x=x, **res)
    lp = res["lnprob_blobs"][:, 0]
    print(
        "%-18s ndim=%3d N_d=%4d W=%d  lnP range [%.3g, %.3g]  rejected=%d  size=%.0f kB"
        % (name, x.shape[1], nd, len(x), lp[lp > -1e99].min() if (lp > -1e99).any() else np.nan, lp.max(), int((lp <= -1e99).sum()), os.path.getsize(path) / 1e3)
    )


def hook_spl_zmax(args):
    args.fid_igm["tau_eff_ztype"] = "interp_spl"
    args.fid_igm["gamma_ztype"] = "interp_spl"
    args.fid_cont["z_max"] = {k: 3.9 for k in args.cont_params}


CONFIGS = {
    # C1: sam_sim-style mock, Chabanier grid, block-diagonal covariance, NN emulator
    "c1_chab_nn": lambda: make("c1_chab_nn", RefMLPEmulator(small_nn(21, "mpg"), "synthetic_mpg_nn", "mpg"), "chab", rebin_k=1, noise=False, seed=100),
    # C4-like: DR1 k range, rebin 8, dense covariance through emu_cov_type="full"
    "dr1_full_nn": lambda: make("dr1_full_nn", RefMLPEmulator(small_nn(22, "mpg"), "synthetic_mpg_nn", "mpg"), "dr1c", rebin_k=8, full=True, emu_cov_type="full", seed=200),
    # C2/C3-like: nyx suite (7 inputs, nrun free), GP emulator, hull + star priors + Gaussian priors + IC correction, interp_spl + z_max
    "nyx_gp_priors": lambda: make(
        "nyx_gp_priors", RefGPEmulator(small_gp(23, "nyx"), "synthetic_nyx_gp", "nyx"), "dr1c", rebin_k=4, full=True, seed=300,
        use_hull=True, vary_alphas=True, name_variation="more_igm", prior_Gauss_rms=0.5, ic_correction=True,
        use_star_priors={"Delta2_star": [0.2, 0.6], "alpha_star": [-0.25, -0.18]}, args_hook=hook_spl_zmax, width=0.06, use_ic=False,
    ),
    # variations: SiVid metals + one-parameter pivot families
    "sivid_pivot": lambda: make("sivid_pivot", RefMLPEmulator(small_nn(24, "mpg"), "synthetic_mpg_nn", "mpg"), "chab", rebin_k=1, seed=400, fit_type="at_a_time_global", name_variation="Metals_Ma2025", use_ic=False),
    # variations: BOSS HCD model
    "hcd_boss": lambda: make("hcd_boss", RefMLPEmulator(small_nn(25, "mpg"), "synthetic_mpg_nn", "mpg"), "dr1c", nz=6, rebin_k=2, seed=500, name_variation="HCD_BOSS"),
    # variations: free HCD_const (pivot, otype const)
    "hcd_const": lambda: make("hcd_const", RefGPEmulator(small_gp(26, "mpg"), "synthetic_mpg_gp", "mpg"), "chab", nz=5, rebin_k=1, full=True, seed=600, name_variation="HCD0"),
    # 11 nodes per family: 189 free parameters
    "chab2019_real": lambda: make_real_chabanier(),
    "global_all": lambda: make("global_all", RefMLPEmulator(small_nn(27, "mpg"), "synthetic_mpg_nn", "mpg"), "chab", rebin_k=1, seed=700, fit_type="global_all", use_ic=False, n=(8, 6, 2), width=0.05),
}


# further baselines of Args.set_baseline (input_pipeline.py:830-1112, 685-766, 1114-1180): parameter sets and
# node layouts that differ structurally from the ones above (families without free parameters, one-node
# pivots, four-node LLS, no contaminants / no resolution, IGM-only and one-z-at-a-time fits)
VARIATIONS = {
    "var_metal_trad": ("global_opt", "metal_trad"),
    "var_no_contaminants": ("global_opt", "no_contaminants"),
    "var_gaikwad21": ("global_opt", "Gaikwad21"),
    "var_turner24": ("global_opt", "Turner24"),
    "var_lls_nz4": ("global_opt", "LLS_nz4"),
    "var_zmin": ("global_opt", "zmin"),
    "var_global_igm": ("global_igm", None),
    "var_at_a_time": ("at_a_time_global", None),
}
def hook_fix_cosmo(args):
    args.fix_cosmo = True


# no cosmological parameter free (fix_cosmo=True, set_like_params.py): the rescaling of stage 1 is the identity
CONFIGS["var_fix_cosmo"] = lambda: make(
    "var_fix_cosmo", RefMLPEmulator(small_nn(49, "mpg"), "synthetic_mpg_nn", "mpg"), "dr1c", nz=5, rebin_k=2, seed=990,
    use_ic=False, n=(8, 8, 2), width=0.1, args_hook=hook_fix_cosmo)

for _i, (_nm, (_ft, _var)) in enumerate(VARIATIONS.items()):
    CONFIGS[_nm] = (lambda nm=_nm, ft=_ft, var=_var, i=_i: make(
        nm, RefMLPEmulator(small_nn(40 + i, "mpg"), "synthetic_mpg_nn", "mpg"), "dr1c", nz=5, rebin_k=2, seed=900 + 10 * i,
        fit_type=ft, name_variation=var, use_ic=False, n=(8, 8, 2), width=0.1))


def make_real_chabanier(name="chab2019_real", nz=11, rebin_k=2, seed=800, n=(24, 20, 4)):
    """The reference's own measured data set: Chabanier et al. (2019) P1D, statistical + systematic
    covariance with its correlation matrices, parsed by the reference's read_from_file
    (cup1d/p1ds/data_Chabanier2019.py:43-120; its P1D_Chabanier2019 class itself fails in
    _drop_zbins, base_p1d_data.py:72, so the z cut is made here), 11 z x 35 k, block-diagonal."""
    rh.setup()
    from cup1d.p1ds import data_Chabanier2019 as dc

    tot = dc.read_from_file(add_syst=True)
    stat = dc.read_from_file(add_syst=False)
    zs = np.asarray(tot[0][:nz], dtype=float)
    ks = [np.asarray(k, dtype=float) for k in tot[1][:nz]]
    kmins, kmaxs = [], []
    for k in ks:
        edges = np.concatenate([[k[0] - 0.5 * (k[1] - k[0])], 0.5 * (k[1:] + k[:-1]), [k[-1] + 0.5 * (k[-1] - k[-2])]])
        kmins.append(edges[:-1])
        kmaxs.append(edges[1:])
    Pk = [np.asarray(p, dtype=float) for p in tot[2][:nz]]
    cov = [np.asarray(c, dtype=float) for c in tot[3][:nz]]
    cov_stat = [np.asarray(c, dtype=float) for c in stat[3][:nz]]
    data = rh.make_data_namespace(zs, ks, kmins, kmaxs, Pk, cov, cov_stat)
    emu = RefMLPEmulator(small_nn(28, "mpg"), "synthetic_mpg_nn", "mpg")
    nd = sum(len(k) for k in ks)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        like, args = rh.build_reference_likelihood(emu, data, rebin_k=rebin_k, emu_cov_type="block")
        x = walkers(like, n[0], n[1], n[2], seed + 2, width=0.1)
        res = run_reference(like, x, zmask=[zs[1], zs[4]])
    spec = spec_from_likelihood(like)
    path = os.path.join(OUT, name + ".npz")
    save_spec(path, spec, x=x, **res)
    lp = res["lnprob_blobs"][:, 0]
    print("%-18s ndim=%3d N_d=%4d W=%d  lnP range [%.3g, %.3g]  rejected=%d  size=%.0f kB"
          % (name, x.shape[1], nd, len(x), lp[lp > -1e99].min(), lp.max(), int((lp <= -1e99).sum()), os.path.getsize(path) / 1e3))


def make_agn_kat():
    """AGN_Model.get_contamination of the reference (AGN_model.py:93-120) on a few (z, coeff)"""
    rh.setup()
    from cup1d.contaminants.AGN_model import AGN_Model

    k = np.linspace(1e-3, 0.04, 200)
    cases, outs = [], []
    for z in (2.0, 2.2, 3.0, 3.7, 4.2, 4.4):
        for coeff in ([0.0, -1.5], [0.8, -0.3], [-2.0], [0.0, -5.5], [0.3, -6.0]):
            m = AGN_Model(ln_AGN_coeff=list(coeff))
            c = m.get_contamination(z, k)
            cases.append([z, len(coeff)] + list(coeff) + [0.0] * (2 - len(coeff)))
            outs.append(np.ones_like(k) * c)
    m = AGN_Model()
    np.savez_compressed(os.path.join(OUT, "agn_kat.npz"), k=k, cases=np.array(cases), out=np.array(outs),
                        agn_z=m.AGN_z, agn_expansion=m.AGN_expansion, null=m.null_value, z0=m.z_0)
    print("agn_kat          %d cases" % len(cases))


def make_mcdonald_kat():
    """HCD_Model_McDonald2005.get_contamination of the reference (hcd_model_McDonald2005.py:83-92) on a few
    (z, coefficients), one redshift per call (its multi-redshift call only runs nulled), free parameters through
    like_params as well as constructor coefficients"""
    rh.setup()
    from cup1d.contaminants.hcd_model_McDonald2005 import HCD_Model_McDonald2005
    from cup1d.likelihood.likelihood_parameter import LikelihoodParameter

    k = np.linspace(1e-3, 0.04, 200)
    cases, outs = [], []
    for z in (2.0, 2.2, 3.0, 3.7, 4.4):
        for coeff in ([0.0, -1.2], [0.7, -0.3], [-2.0], [0.0, -6.0], [0.3, -6.5], [1.5, 2.0]):
            m = HCD_Model_McDonald2005(ln_A_damp_coeff=list(coeff))
            c = m.get_contamination(z, k)
            cases.append([z, len(coeff)] + list(coeff) + [0.0] * (2 - len(coeff)))
            outs.append(np.ones_like(k) * c)
    # through like_params: parameter i feeds coefficient [n-1-i]
    m = HCD_Model_McDonald2005(free_param_names=["ln_A_damp_0", "ln_A_damp_1"])
    lp = [LikelihoodParameter(name="ln_A_damp_0", value=-0.8, min_value=-7, max_value=2.5),
          LikelihoodParameter(name="ln_A_damp_1", value=1.1, min_value=-10, max_value=10)]
    cases.append([3.4, 2, 1.1, -0.8])
    outs.append(np.ones_like(k) * m.get_contamination(3.4, k, like_params=lp))
    np.savez_compressed(os.path.join(OUT, "mcdonald_kat.npz"), k=k, cases=np.array(cases), out=np.array(outs),
                        null=m.null_value, z0=m.z_0)
    print("mcdonald_kat     %d cases" % len(cases))


def make_theory_kat(name, emu, grid, nz, rebin_k, seed, **kw):
    """Theory.get_p1d_kms (lya_theory.py:610-814) and Likelihood.get_p1d_kms (likelihood.py:722-785) of the
    reference off the likelihood's own grid: a redshift subset, a custom k grid, full / partial / empty
    like_params (absent parameters keep the models' coefficients, the resolution term needs an R_coeff
    parameter in the list), points outside the unit cube, values=None."""
    zs = np.round(2.2 + 0.2 * np.arange(nz), 10)
    ks, kmins, kmaxs = grids(grid, nz)
    nd = sum(len(k) for k in ks)

    def data_ns(Pk_full, cov_full):
        Pk, cov, off = [], [], 0
        for k in ks:
            Pk.append(Pk_full[off : off + len(k)])
            cov.append(cov_full[off : off + len(k), off : off + len(k)])
            off += len(k)
        return rh.make_data_namespace(zs, ks, kmins, kmaxs, Pk, cov, [0.8 * c for c in cov])

    rng = np.random.default_rng(seed)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        like, args = rh.build_reference_likelihood(emu, data_ns(np.ones(nd), np.eye(nd) * 0.01), rebin_k=rebin_k, **kw)
        x = walkers(like, 4, 4, 0, seed + 2, width=0.08)
        x = np.concatenate([x, x[:2]])
        x[-2, 3] = 1.07   # outside the unit cube: Theory has no cube test
        x[-1, 0] = -0.04
        izs = [1, 3, 2] if nz > 3 else [0]
        zsub = zs[izs]
        kc = [np.linspace(0.0012 + 0.0003 * j, 0.035 - 0.002 * j, 23 + 5 * j) for j in range(len(izs))]
        ndc = sum(len(k) for k in kc)
        out = {"x": x, "izs": np.array(izs), "kc": np.concatenate(kc), "kc_len": np.array([len(k) for k in kc]),
               "p_full": np.full((len(x), ndc), np.nan), "blob_full": np.zeros((len(x), 6)),
               "emu_full": np.zeros((len(x), len(izs), len(emu.emu_params))),
               "p_empty": np.zeros(ndc), "p_partial": np.full((len(x), ndc), np.nan), "p_res_only": np.full((len(x), ndc), np.nan),
               "p_like_sub": np.full((len(x), sum(len(ks[i]) for i in izs)), np.nan), "p_like_none": np.zeros(nd)}
        names = [p.name for p in like.free_params]
        # whole families only: the reference raises "number of params mismatch" otherwise (base_igm.py:306-308)
        fams = sorted({n.rsplit("_", 1)[0] for n in names if n not in ("As", "ns", "nrun")})
        keep = [n for n in names if (n in ("As",)) or (n.rsplit("_", 1)[0] in fams[::2] and not n.startswith("R_coeff"))]
        out["partial_names"] = np.array(keep, dtype=str)
        for i, v in enumerate(x):
            lp = like.parameters_from_sampling_point(v.copy())
            r = like.theory.get_p1d_kms(zsub, kc, like_params=lp, return_emu_params=True)
            if r is not None:
                out["p_full"][i] = np.concatenate(r[0])
                out["blob_full"][i] = r[1]
                for j, nm in enumerate(emu.emu_params):
                    out["emu_full"][i, :, j] = r[2][nm]
            r = like.theory.get_p1d_kms(zsub, kc, like_params=[p for p in lp if p.name in keep], return_blob=False)
            if r is not None:
                out["p_partial"][i] = np.concatenate(r)
            r = like.theory.get_p1d_kms(zsub, kc, like_params=[p for p in lp if p.name.startswith("R_coeff")], return_blob=False)
            if r is not None:
                out["p_res_only"][i] = np.concatenate(r)
            r = like.get_p1d_kms(zs=zsub, _k_kms=(kc if rebin_k == 1 else None), values=v.copy())
            if r is not None and rebin_k != 1:
                out["p_like_sub"][i] = np.concatenate(r[0])
            elif r is not None:
                out["p_like_sub"] = out["p_like_sub"] if out["p_like_sub"].shape[1] == ndc else np.full((len(x), ndc), np.nan)
                out["p_like_sub"][i] = np.concatenate(r[0])
        out["p_empty"] = np.concatenate(like.theory.get_p1d_kms(zsub, kc, like_params=[], return_blob=False))
        out["p_like_none"] = np.concatenate(like.get_p1d_kms()[0])
    spec = spec_from_likelihood(like)
    save_spec(os.path.join(OUT, name + ".npz"), spec, **out)
    print("%-18s ndim=%3d  %d points, %d rejected, custom grid %d bins over z = %s" % (name, x.shape[1], len(x), int(np.isnan(out["p_full"]).all(axis=1).sum()), ndc, zsub))


CONFIGS["theory_kat_chab"] = lambda: make_theory_kat("theory_kat_chab", RefMLPEmulator(small_nn(51, "mpg"), "synthetic_mpg_nn", "mpg"), "chab", 6, 1, 1100)
CONFIGS["theory_kat_nyx"] = lambda: make_theory_kat(
    "theory_kat_nyx", RefGPEmulator(small_gp(52, "nyx"), "synthetic_nyx_gp", "nyx"), "dr1c", 5, 4, 1200,
    use_hull=True, vary_alphas=True, use_star_priors={"Delta2_star": [0.2, 0.6]}, use_ic=False)
CONFIGS["agn_kat"] = make_agn_kat
CONFIGS["mcdonald_kat"] = make_mcdonald_kat

if __name__ == "__main__":
    assert rh.available(), "needs /root/reference"
    todo = sys.argv[1:] or list(CONFIGS)
    for nm in todo:
        CONFIGS[nm]()
