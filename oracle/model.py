"""TEST INFRASTRUCTURE — CPU restatement (NumPy/SciPy float64) of cup1d's
per-walker P1D log-likelihood, stages 1, 3, 4 and 5 of the hot path.

It evaluates a *model spec* (cup1d_b200/spec.py) one walker at a time with the
same structure as the reference (per-z loops over NumPy vectors in k), and
every function cites the reference lines it follows.  It is pinned against

* data/zenodo/fig_5.npy, fig_6.npy of the reference (contaminant templates,
  tests/test_kat_contaminants.py), and
* golden outputs of the reference's own ``Likelihood.log_prob_and_blobs`` /
  ``get_p1d_kms`` / ``get_log_like`` run in the build container
  (tools/make_golden.py -> tests/golden/*.npz, tests/test_oracle_golden.py).

Stage 2 (the LaCE emulator) is restated in oracle/emulators.py — parity
unpinned, see its header.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product never does.
"""

import math

import numpy as np
from scipy.interpolate import make_interp_spline

from .emulators import emulator_from_spec

C_KMS = 2.99792458e5
LYA_AA = 1215.67


# ---------------------------------------------------------------------------
# coefficient families
# ---------------------------------------------------------------------------
def family_coeffs(fam, pvalues):
    """Coefficient vector with free parameters substituted
    (base_igm.py:289-319, base_contaminants.py:215-245)."""
    coeff = np.array(fam["coeffs"], dtype=float)
    if pvalues is None:
        return coeff
    idx = np.asarray(fam["param_idx"])
    for j in range(len(coeff)):
        if idx[j] >= 0:
            coeff[j] = pvalues[idx[j]]
    return coeff


def family_value(fam, z, pvalues):
    """g(z) of one family, after the z_max override and the output type
    (base_igm.py:245-280, base_contaminants.py:158-206)."""
    z = np.atleast_1d(np.asarray(z, dtype=float))
    coeff = family_coeffs(fam, pvalues)
    ztype = fam["ztype"]
    if ztype == "pivot":
        xz = np.log((1 + z) / (1 + fam["z0"]))
        ln_out = np.polyval(coeff, xz)
    elif ztype == "interp_lin":
        ln_out = np.interp(z, fam["znodes"], coeff)
    elif ztype == "interp_spl":
        ln_out = make_interp_spline(fam["znodes"], coeff, k=1)(z)
    else:
        raise ValueError("unsupported ztype " + ztype)
    ln_out = np.array(ln_out, dtype=float)
    if not np.isnan(fam["zmax"]):
        ln_out[z >= fam["zmax"]] = fam["null"]
    if fam["otype"] == "exp":
        return np.exp(ln_out)
    return ln_out


def nulled_value(fam, z, pvalues):
    """Family value with the ``<= null`` entries zeroed
    (si_mult.py:165-176, si_add.py:138-149, hcd_model_rogers_class.py:100-111)."""
    val = family_value(fam, z, pvalues)
    if not np.isnan(fam["null"]):
        null = fam["null"] if fam["otype"] == "const" else np.exp(fam["null"])
        val[val <= null] = 0
    return val


# ---------------------------------------------------------------------------
# contaminant templates (stage 3)
# ---------------------------------------------------------------------------
SIMULT_OFF_DEFAULT = {
    "SiIII_Lya": 1,
    "SiIIa_Lya": 1,
    "SiIIb_Lya": 1,
    "SiIIc_Lya": 0,
    "SiIII_SiIIa": 1,
    "SiIII_SiIIb": 1,
    "SiIII_SiIIc": 0,
    "SiIIc_SiIIb": 0,
    "SiIIc_SiIIa": 0,
    "SiIIb_SiIIa": 0,
}


def simult_switches(off, remove=None):
    """Per-call switch table of SiMult (si_mult.py:185-206)."""
    off = dict(off)
    if remove is not None:
        for key in remove:
            if key in off:
                off[key] = remove[key]
    # SiII-SiII terms are additive only
    off["SiIIb_SiIIa"] = 0
    off["SiIIc_SiIIa"] = 0
    off["SiIIc_SiIIb"] = 0
    return off


def si_mult_contamination(sm, vals, k_kms, mF, remove=None):
    """SiIII/SiII multiplicative oscillations (si_mult.py:157-343).

    sm: dict(dv, rat, off); vals: family name -> [Nz] nulled values;
    k_kms: list of Nz arrays; mF: [Nz].  Returns list of Nz arrays."""
    dv, rat = sm["dv"], sm["rat"]
    off = simult_switches(sm["off"], remove)
    ra3, rb3, rc3 = rat["SiIIa_SiIII"], rat["SiIIb_SiIII"], rat["SiIIc_SiIII"]
    out = []
    for iz in range(len(k_kms)):
        k = k_kms[iz]
        G3 = 2 - 2 / (1 + np.exp(-vals["s_Lya_SiIII"][iz] * k))
        G2 = 2 - 2 / (1 + np.exp(-vals["s_Lya_SiII"][iz] * k))
        Gmm = vals["f_SiIIb_SiIII"][iz]
        a3 = vals["f_Lya_SiIII"][iz] / (1 - mF[iz])
        a2 = rb3 * vals["f_Lya_SiII"][iz] / (1 - mF[iz])
        r = ra3 * vals["f_SiIIa_SiIII"][iz] / rb3
        c = rc3 / rb3
        C0 = a3**2 * off["SiIII_Lya"] + a2**2 * (
            r**2 * off["SiIIa_Lya"] + off["SiIIb_Lya"] + c**2 * off["SiIIc_Lya"]
        )
        C3 = 2 * a3 * G3 * off["SiIII_Lya"] * np.cos(dv["SiIII_Lya"] * k)
        C2 = (
            2
            * a2
            * G2
            * (
                off["SiIIa_Lya"] * r * np.cos(dv["SiIIa_Lya"] * k)
                + off["SiIIb_Lya"] * np.cos(dv["SiIIb_Lya"] * k)
                + off["SiIIc_Lya"] * c * np.cos(dv["SiIIc_Lya"] * k)
            )
        )
        Cmm = (
            2
            * a3
            * a2
            * Gmm
            * (
                off["SiIII_SiIIc"] * c * np.cos(dv["SiIII_SiIIc"] * k)
                + off["SiIII_SiIIb"] * np.cos(dv["SiIII_SiIIb"] * k)
                + off["SiIII_SiIIa"] * r * np.cos(dv["SiIII_SiIIa"] * k)
            )
        )
        # the SiII-SiII term (Cm) is switched off unconditionally (:203-206)
        out.append(1 + C0 + (C3 + C2) + Cmm)
    return out


def si_vid_contamination(sm, vals, k_kms, mF):
    """Ma+2025 variant (si_vid_final.py:157-221)."""
    dv = sm["dv"]
    out = []
    for iz in range(len(k_kms)):
        k = k_kms[iz]
        p3 = vals["f_Lya_SiIII"][iz] / (1 - mF[iz]) * np.exp(vals["s_Lya_SiIII"][iz] * k)
        p2 = (
            vals["f_Lya_SiII"][iz]
            / (1 - mF[iz])
            * np.cos(dv["SiIII_Lya"] * k)
            * np.exp(-vals["s_Lya_SiII"][iz] * k)
        )
        out.append(1 + p3 + p2)
    return out


def si_add_ktot(sa, k, remove=None):
    """Walker-independent shape of the additive SiII-SiII term
    (si_add.py:172-219)."""
    dv, rat = sa["dv"], sa["rat"]
    off = dict(sa["off"])
    if remove is not None:
        for key in remove:
            if key in off:
                off[key] = remove[key]
    rac, rbc, rab = rat["SiIIa_SiIIc"], rat["SiIIb_SiIIc"], rat["SiIIa_SiIIb"]
    dca, dcb, dba = dv["SiIIc_SiIIa"], dv["SiIIc_SiIIb"], dv["SiIIb_SiIIa"]
    Cac = 1 + rac**2 + 2 * rac * np.cos(dca * k)
    Cbc = 1 + rbc**2 + 2 * rbc * np.cos(dcb * k)
    Cba = rbc**2 * (1 + rab**2 + 2 * rab * np.cos(dba * k))
    t_acbc = 2 * (1 + rac * rbc) * np.cos(0.5 * (dca - dcb) * k) + 2 * (rac + rbc) * np.cos(
        0.5 * (dca + dcb) * k
    )
    t_acab = 2 * rbc * (1 + rac * rab) * np.cos(0.5 * (dca - dba) * k) + 2 * rbc * (
        rac + rab
    ) * np.cos(0.5 * (dca + dba) * k)
    t_bcab = 2 * rbc * (1 + rbc * rab) * np.cos(0.5 * (dcb - dba) * k) + 2 * rbc * (
        rbc + rab
    ) * np.cos(0.5 * (dcb + dba) * k)
    return (
        off["SiIIc_SiIIa"] * Cac
        + off["SiIIc_SiIIb"] * Cbc
        + off["SiIIb_SiIIa"] * Cba
        + off["SiIIacbc"] * t_acbc
        + off["SiIIacab"] * t_acab
        + off["SiIIbcab"] * t_bcab
    )


def si_add_contamination(sa, vals, k_kms, remove=None):
    """Additive SiII-SiII term in km/s units (si_add.py:130-221)."""
    out = []
    for iz in range(len(k_kms)):
        k = k_kms[iz]
        a = vals["f_SiIIa_SiIIb"][iz]
        s = vals["s_SiIIa_SiIIb"][iz]
        out.append(a**2 * si_add_ktot(sa, k, remove) * np.exp(-1 * s**2 * k**2))
    return out


def rogers_profiles(rg, z, k):
    """The four damping-wing profiles 1/(a e^{bk} - 1)^2
    (hcd_model_rogers_class.py:5-6, 115-131)."""
    prof = []
    for it in range(4):
        a_z = rg["a0"][it] * ((1 + z) / (1 + rg["z0"])) ** rg["a1"][it]
        b_z = rg["b0"][it] * ((1 + z) / (1 + rg["z0"])) ** rg["b1"][it]
        prof.append(1 / (a_z * np.exp(k * b_z) - 1) ** 2)
    return prof


def hcd_rogers_contamination(rg, vals, zs, k_kms):
    """hcd_model_rogers_class.py:94-140"""
    out = []
    for iz in range(len(k_kms)):
        k = k_kms[iz]
        cont = 1 + vals["HCD_const"][iz] + np.zeros_like(k)
        prof = rogers_profiles(rg, zs[iz], k)
        for it in range(4):
            cont += vals["HCD_damp%d" % (it + 1)][iz] * prof[it]
        out.append(cont)
    return out


def hcd_mcdonald_contamination(fam, zs, k_kms, pvalues):
    """hcd_model_McDonald2005.py:71-92: 1 + A_damp(z) (0.018 + 1/(15000 k - 8.9)), A_damp = exp(poly(ln((1+z)/(1+z_0)))),
    the whole term 1 when the constant coefficient is <= the null value (:75-76).  The reference's own multi-redshift
    call only runs in that nulled state (``if A_damp == 0`` on an array raises otherwise); the formula is applied
    per redshift bin here, as its single-redshift form reads."""
    coeff = family_coeffs(fam, pvalues)
    out = []
    for iz in range(len(k_kms)):
        if coeff[-1] <= fam["null"]:
            out.append(np.ones_like(k_kms[iz]))
            continue
        xz = np.log((1 + zs[iz]) / (1 + fam["z0"]))
        a_damp = np.exp(np.polyval(coeff, xz))
        out.append(1 + a_damp * (0.018 + 1 / (15000 * k_kms[iz] - 8.9)))
    return out


def hcd_boss_contamination(vals, k_kms):
    """hcd_boss.py:5-7, 66-91"""
    out = []
    for iz in range(len(k_kms)):
        k = k_kms[iz]
        out.append(1 + 1 / (1 - (1 / (15000 * k - 8.9))) * vals["HCD_damp1"][iz])
    return out


def resolution_Rz(z):
    """resolution_class.py:26-34"""
    return C_KMS * 0.8 / (1 + z) / LYA_AA


def resolution_contamination(vals, zs, k_kms):
    """resolution_class.py:88-114"""
    return [
        1 + 2 * vals["R_coeff"][iz] * resolution_Rz(zs[iz]) ** 2 * k_kms[iz] ** 2
        for iz in range(len(k_kms))
    ]


def nyx_ic_correction(zs, k_kms):
    """model_contaminants.py:310-330"""
    pz = np.array([0.15261529, -2.30600644, 2.61877894])
    kc = 0.003669741766936781
    out = []
    for iz in range(len(k_kms)):
        anc = (pz[0] * zs[iz] ** 2 + pz[1] * zs[iz] + pz[2]) * (1 - np.exp(-k_kms[iz] / kc))
        out.append(1 / (1 - anc / 100))
    return out


def agn_delta(agn, z, k):
    """AGN feedback shape delta(z, k) (AGN_model.py:99-116): a + b exp(-c k) tabulated at the
    redshifts of data/nuisance/AGN_corr.dat, interpolated linearly in z, extrapolated from the
    first two rows above the table"""
    from scipy.interpolate import interp1d

    az, ex = np.asarray(agn["z"], dtype=float), np.asarray(agn["expansion"], dtype=float)
    if z <= np.max(az):
        yy = ex[:, 0][None, :] + ex[:, 1][None, :] * np.exp(-ex[:, 2][None, :] * k[:, None])
        return interp1d(az, yy)(z)
    upper = ex[0, 0] + ex[0, 1] * np.exp(-ex[0, 2] * k)
    lower = ex[1, 0] + ex[1, 1] * np.exp(-ex[1, 2] * k)
    return (upper - lower) / (az[0] - az[1]) * (z - az[0]) + upper


def agn_contamination(agn, fam, zs, k_kms, pvalues):
    """1 + f_AGN(z) delta(z, k) (AGN_model.py:81-120).  The whole term is switched off when
    the constant coefficient is <= the null value (:85-86).  Disabled in the reference
    (model_contaminants.py:138-154, 242-252); here an optional term."""
    coeff = family_coeffs(fam, pvalues)
    out = []
    for iz in range(len(k_kms)):
        if coeff[-1] <= fam["null"]:
            out.append(np.ones_like(k_kms[iz]))
            continue
        xz = np.log((1 + zs[iz]) / (1 + fam["z0"]))
        f_agn = np.exp(np.polyval(coeff, xz))
        out.append(1 + agn_delta(agn, zs[iz], k_kms[iz]) * f_agn)
    return out


# ---------------------------------------------------------------------------
# Theory (stages 1, 2 call, 3, 5a/5b)
# ---------------------------------------------------------------------------
class OracleTheory(object):
    def __init__(self, spec, emulator=None):
        self.spec = spec
        self.zs = np.asarray(spec["zs"], dtype=float)
        self.emulator = emulator if emulator is not None else emulator_from_spec(spec["emulator"])
        self.names = list(spec["params"]["names"])

    # -- stage 1 ----------------------------------------------------------
    def _primordial_shifts(self, pvalues):
        """ratio_As, delta_ns, delta_nrun (lya_theory.py:281-294, 542-555)"""
        c = self.spec["cosmo"]
        ratio_As, d_ns, d_nrun = 1.0, 0.0, 0.0
        if pvalues is not None:
            if "As" in self.names:
                ratio_As = pvalues[self.names.index("As")] / c["As_fid"]
            if "ns" in self.names:
                d_ns = pvalues[self.names.index("ns")] - c["ns_fid"]
            if "nrun" in self.names:
                d_nrun = pvalues[self.names.index("nrun")] - c["nrun_fid"]
        return ratio_As, d_ns, d_nrun

    def linP_Mpc_params(self, iz_sel, pvalues):
        """lya_theory.py:270-351"""
        c = self.spec["cosmo"]
        ratio_As, d_ns, d_nrun = self._primordial_shifts(pvalues)
        ln_kp_ks = np.log(c["kp_emu_Mpc"] / c["ks_Mpc"])
        d_alpha = d_nrun
        d_n = d_ns + d_nrun * ln_kp_ks
        ln_ratio = np.log(ratio_As) + (d_ns + 0.5 * d_nrun * ln_kp_ks) * ln_kp_ks
        fid = c["linP_fid"][iz_sel]
        return {
            "Delta2_p": fid[:, 0] * np.exp(ln_ratio),
            "n_p": fid[:, 1] + d_n,
            "alpha_p": fid[:, 2] + d_alpha,
        }

    def blob(self, pvalues):
        """lya_theory.py:536-584"""
        c = self.spec["cosmo"]
        ratio_As, d_ns, d_nrun = self._primordial_shifts(pvalues)
        ln_kp_ks = np.log(c["kp_star_Mpc"] / c["ks_Mpc"])
        fb = c["blob_fid"]
        ln_ratio = np.log(ratio_As) + (d_ns + 0.5 * d_nrun * ln_kp_ks) * ln_kp_ks
        return (
            fb[0] * np.exp(ln_ratio),
            fb[1] + d_ns + d_nrun * ln_kp_ks,
            fb[2] + d_nrun,
            fb[3],
            fb[4],
            fb[5],
        )

    def emulator_calls(self, iz_sel, pvalues):
        """lya_theory.py:410-498; IGM histories mean_flux_class.py:51-61,
        thermal_class.py:52-72, pressure_class.py:51-56"""
        s = self.spec
        z = self.zs[iz_sel]
        M = s["cosmo"]["M_of_z"][iz_sel]
        fam, fid = s["families"], s["igm_fid"]
        linP = self.linP_Mpc_params(iz_sel, pvalues)
        call = {}
        for key in self.emulator.emu_params:
            if key in ("Delta2_p", "n_p", "alpha_p"):
                call[key] = linP[key]
            elif key == "mF":
                tau = family_value(fam["tau_eff"], z, pvalues) * fid["tau_eff"][iz_sel]
                call[key] = np.exp(-tau)
                tau0 = family_value(fam["tau_eff"], z, None) * fid["tau_eff"][iz_sel]
                call["mF_fid"] = np.exp(-tau0)
            elif key == "gamma":
                call[key] = family_value(fam["gamma"], z, pvalues) * fid["gamma"][iz_sel]
            elif key == "sigT_Mpc":
                call[key] = family_value(fam["sigT_kms"], z, pvalues) * fid["sigT_kms"][iz_sel] / M
            elif key == "kF_Mpc":
                call[key] = family_value(fam["kF_kms"], z, pvalues) * fid["kF_kms"][iz_sel] * M
            else:
                raise ValueError("Not a theory model for emulator parameter " + key)
        return call, M

    # -- stage 5a/5b ------------------------------------------------------
    def in_star_box(self, blob):
        """star-prior box on the blob (lya_theory.py:647-654)"""
        sp = self.spec["cosmo"]["star_priors"]
        if sp is None:
            return True
        for i in range(3):
            if not np.isnan(sp[i, 0]):
                if (blob[i] > sp[i, 1]) or (blob[i] < sp[i, 0]):
                    return False
        return True

    def in_hulls(self, call):
        """all 2-D hull projections at all z (hull.py:9-10, 157-165)"""
        hull = self.spec["hull"]
        if hull is None:
            return True
        p0 = np.stack([call[key] for key in hull["params"]], axis=1)
        for h in hull["hulls"]:
            p = p0[:, [h["dim0"], h["dim1"]]]
            inside = np.all(h["A"] @ p.T + h["b"][:, None] <= hull["tol"], 0)
            if not inside.all():
                return False
        return True

    # -- the model P1D ----------------------------------------------------
    def get_p1d_kms(self, iz_sel, k_kms, pvalues, apply_hull=True, remove=None, parts=False):
        """lya_theory.py:610-814 for the z-bins ``iz_sel``.  Returns
        (p1d list | None, blob, emu_call)."""
        s = self.spec
        z = self.zs[iz_sel]
        call, M = self.emulator_calls(iz_sel, pvalues)
        blob = self.blob(pvalues)
        if not self.in_star_box(blob):
            return None, blob, call
        if apply_hull and not self.in_hulls(call):
            return None, blob, call
        nz = len(z)
        length = max(len(k) for k in k_kms)
        kin = np.zeros((nz, length))
        for iz in range(nz):
            kin[iz, : len(k_kms[iz])] = k_kms[iz] * M[iz]
        p_Mpc = self.emulator.emulate_p1d_Mpc(call, kin)
        p_emu = [p_Mpc[iz][: len(k_kms[iz])] * M[iz] for iz in range(nz)]

        fam, cont = s["families"], s["cont"]
        if cont["apply_resolution"]:
            vals = {"R_coeff": family_value(fam["R_coeff"], z, pvalues)}
            syst = resolution_contamination(vals, z, k_kms)
        else:
            syst = np.ones(nz)

        mF = call["mF"]
        if cont["metal_model"] == "SiMult":
            names = ["f_Lya_SiIII", "s_Lya_SiIII", "f_Lya_SiII", "s_Lya_SiII", "f_SiIIa_SiIII", "f_SiIIb_SiIII"]
            vals = {n: nulled_value(fam[n], z, pvalues) for n in names}
            mul = si_mult_contamination(cont["si_mult"], vals, k_kms, mF, remove)
        else:
            names = ["f_Lya_SiIII", "s_Lya_SiIII", "f_Lya_SiII", "s_Lya_SiII"]
            vals = {n: nulled_value(fam[n], z, pvalues) for n in names}
            mul = si_vid_contamination(cont["si_mult"], vals, k_kms, mF)
        vals = {n: nulled_value(fam[n], z, pvalues) for n in ["f_SiIIa_SiIIb", "s_SiIIa_SiIIb"]}
        add = si_add_contamination(cont["si_add"], vals, k_kms, remove)
        if cont["hcd_model"] == "rogers":
            names = ["HCD_damp1", "HCD_damp2", "HCD_damp3", "HCD_damp4", "HCD_const"]
            vals = {n: nulled_value(fam[n], z, pvalues) for n in names}
            hcd = hcd_rogers_contamination(cont["rogers"], vals, z, k_kms)
        elif cont["hcd_model"] == "mcdonald":
            hcd = hcd_mcdonald_contamination(fam["HCD_damp1"], z, k_kms, pvalues)
        else:
            vals = {"HCD_damp1": nulled_value(fam["HCD_damp1"], z, pvalues)}
            hcd = hcd_boss_contamination(vals, k_kms)
        if cont["ic_correction"]:
            ic = nyx_ic_correction(z, k_kms)
        else:
            ic = np.ones(nz)

        if cont.get("agn") is not None and "ln_AGN" in fam:
            agn = agn_contamination(cont["agn"], fam["ln_AGN"], z, k_kms, pvalues)
        else:
            agn = np.ones(nz)

        out = []
        for iz in range(nz):
            # lya_theory.py:778-784 (with the multiplicative AGN term of the older
            # mult_cont_total = metals * HCD * SN * AGN * IC, model_contaminants.py:259-262)
            out.append((hcd[iz] * mul[iz] * agn[iz] * ic[iz] * p_emu[iz] + add[iz]) * syst[iz])
        if parts:
            return out, blob, call, {"p_emu": p_emu, "mul": mul, "add": add, "hcd": hcd, "syst": syst}
        return out, blob, call


# ---------------------------------------------------------------------------
# Likelihood (stages 0, 4, 5)
# ---------------------------------------------------------------------------
class OracleLikelihood(object):
    """Scalar API with the reference's names and return conventions
    (likelihood.py:722-1072)."""

    def __init__(self, spec, emulator=None):
        self.spec = spec
        self.theory = OracleTheory(spec, emulator)
        p = spec["params"]
        self.pmin = np.asarray(p["min"], dtype=float)
        self.pmax = np.asarray(p["max"], dtype=float)
        self.min_log_like = float(p["min_log_like"])
        self.Gauss_priors = p["gauss_sigma_cube"]
        # LikelihoodParameter.value_in_cube (likelihood_parameter.py:25-28)
        self.fid_cube = (np.asarray(p["value"], dtype=float) - self.pmin) / (self.pmax - self.pmin)
        self.zs = np.asarray(spec["zs"], dtype=float)
        self.nz = len(self.zs)

    def parameters_from_sampling_point(self, values):
        """likelihood.py:561-574, likelihood_parameter.py:58-61"""
        if values is None:
            return None
        values = np.asarray(values, dtype=float)
        assert len(values) == len(self.pmin), "size mismatch"
        return self.pmin + values * (self.pmax - self.pmin)

    def rebinning(self, iz_sel, p_fine):
        """likelihood.py:147-161"""
        rb = self.spec["rebin"]
        out = []
        for ii, iz in enumerate(iz_sel):
            out.append(np.sum(rb["cover"][iz] * p_fine[ii][np.newaxis, :], axis=1) / rb["sum_cover"][iz])
        return out

    def get_p1d_kms(self, values=None, iz_sel=None, apply_hull=True, remove=None, return_blob=False, return_emu_params=False):
        """likelihood.py:722-785 (on the data redshifts ``iz_sel``)"""
        if iz_sel is None:
            iz_sel = np.arange(self.nz)
        iz_sel = np.atleast_1d(iz_sel)
        s = self.spec
        if s["rebin"] is not None:
            k_kms = [s["rebin"]["k_kms"][iz] for iz in iz_sel]
        else:
            k_kms = [s["data"]["k_kms"][iz] for iz in iz_sel]
        pvalues = self.parameters_from_sampling_point(values)
        p1d, blob, call = self.theory.get_p1d_kms(iz_sel, k_kms, pvalues, apply_hull=apply_hull, remove=remove)
        if p1d is None:
            return None
        if s["rebin"] is not None:
            p1d = self.rebinning(iz_sel, p1d)
        out = [p1d]
        if return_blob:
            out.append(blob)
        if return_emu_params:
            out.append(call)
        return out

    def _zsel(self, zmask):
        if zmask is None:
            return np.arange(self.nz)
        zmask = np.atleast_1d(zmask)
        return np.array([iz for iz in range(self.nz) if np.any(np.abs(zmask - self.zs[iz]) < 1e-3)], dtype=int)

    def get_log_like(self, values=None, ignore_log_det_cov=True, return_blob=False, zmask=None):
        """likelihood.py:822-972"""
        null_out = [-np.inf, -np.inf]
        if return_blob:
            null_out.append((0, 0, 0, 0, 0, 0))
        if values is not None:
            values = np.asarray(values, dtype=float)
            if (values > 1.0).any() | (values < 0.0).any():
                return null_out
        d = self.spec["data"]
        sel = self._zsel(zmask)
        blob = None
        model = [None] * self.nz
        if zmask is not None:
            # one Theory call per selected z (:844-863): priors are tested per z
            for iz in sel:
                res = self.get_p1d_kms(values, [iz], return_blob=True)
                if res is None:
                    return null_out
                model[iz] = res[0][0]
                blob = res[1]
        else:
            res = self.get_p1d_kms(values, None, return_blob=True)
            if res is None:
                return null_out
            for iz in range(self.nz):
                model[iz] = res[0][iz]
            blob = res[1]

        log_like_all = np.zeros((1, self.nz))
        for iz in sel:
            diff = d["Pk_kms"][iz] - np.array(model[iz]).reshape(-1)
            chi2_z = np.dot(np.dot(d["icov"][iz], diff), diff)
            if ignore_log_det_cov:
                log_like_all[0, iz] = -0.5 * chi2_z
            else:
                log_det_cov = np.log(np.abs(1 / np.linalg.det(d["icov"][iz])))
                log_like_all[0, iz] = -0.5 * (chi2_z + log_det_cov)
        log_like = 0
        if (d["full_icov"] is None) | (zmask is not None):
            log_like += np.sum(log_like_all[0])
        else:
            diff = d["full_Pk_kms"] - np.concatenate(model)
            chi2_all = np.dot(np.dot(d["full_icov"], diff), diff)
            if ignore_log_det_cov:
                log_like += -0.5 * chi2_all
            else:
                log_det_cov = np.log(np.abs(1 / np.linalg.det(d["full_icov"])))
                log_like += -0.5 * (chi2_all + log_det_cov)
        if np.isnan(log_like):
            return null_out
        out = [log_like, log_like_all]
        if return_blob:
            out.append(blob)
        return out

    def get_chi2(self, values=None, return_all=False, zmask=None):
        """likelihood.py:787-798"""
        log_like, log_like_all = self.get_log_like(values, ignore_log_det_cov=True, zmask=zmask)
        if return_all:
            return -2.0 * log_like, -2.0 * log_like_all
        return -2.0 * log_like

    def regulate_log_like(self, log_like):
        """likelihood.py:974-980"""
        if (log_like is None) or math.isnan(log_like):
            return self.min_log_like
        return max(self.min_log_like, log_like)

    def get_log_prior(self, values):
        """likelihood.py:1049-1064"""
        if max(values) > 1:
            return self.min_log_like
        if min(values) < 0:
            return self.min_log_like
        return -np.sum((self.fid_cube - values) ** 2 / (2 * self.Gauss_priors**2))

    def compute_log_prob(self, values, return_blob=False, ignore_log_det_cov=True, zmask=None):
        """likelihood.py:982-1024"""
        values = np.asarray(values, dtype=float)
        if (np.max(values) > 1.0) or (np.min(values) < 0.0):
            if return_blob:
                return self.min_log_like, (np.nan,) * 6
            return self.min_log_like
        log_prior = self.get_log_prior(values) if self.Gauss_priors is not None else 0
        res = self.get_log_like(values, ignore_log_det_cov=ignore_log_det_cov, return_blob=True, zmask=zmask)
        log_like = self.regulate_log_like(res[0])
        if return_blob:
            return log_like + log_prior, res[2]
        return log_like + log_prior

    def log_prob(self, values, ignore_log_det_cov=True, zmask=None):
        return self.compute_log_prob(values, False, ignore_log_det_cov, zmask)

    def log_prob_and_blobs(self, values, ignore_log_det_cov=True, zmask=None):
        """likelihood.py:1036-1047"""
        lnprob, blob = self.compute_log_prob(values, True, ignore_log_det_cov, zmask)
        return (lnprob, *blob)

    def minus_log_prob(self, values, zmask=None, ind_fix=None, pfix=None):
        """likelihood.py:1066-1072"""
        if ind_fix is not None:
            values[ind_fix] = pfix
        return -1.0 * self.log_prob(values, zmask=zmask)

    # convenience for tests / the CPU baseline: loop the scalar path
    def log_prob_and_blobs_many(self, values2d):
        out = np.zeros((len(values2d), 7))
        for i, v in enumerate(values2d):
            out[i] = self.log_prob_and_blobs(v)
        return out
