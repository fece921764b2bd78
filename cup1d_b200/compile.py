"""Context compiler: lower a model spec (cup1d_b200/spec.py) to the flat
constant tables of the device context (include/cup1d_b200.h).

Every walker-independent quantity is evaluated here once, on the host, in
float64: the cos(dv k) oscillation tables, the HCD
damping profiles, the resolution and IC shapes, the rebinning operator, the
interpolation weights that turn each coefficient family into a linear
function of the free parameters.  What is left for the device per walker is
the arithmetic that actually depends on the walker.
"""

from __future__ import annotations

import numpy as np

from .spec import model_k_grid, spec_sizes

C_KMS = 2.99792458e5
LYA_AA = 1215.67

NQ = 20
NCOL = 16
MAX_EMU = 8
MAX_COEF = 8
MAX_LAYERS = 12

# c1d_quantity order
QUANTITIES = [
    "tau_eff", "sigT_kms", "gamma", "kF_kms",
    "f_Lya_SiIII", "s_Lya_SiIII", "f_Lya_SiII", "s_Lya_SiII",
    "f_SiIIa_SiIII", "f_SiIIb_SiIII",
    "f_SiIIa_SiIIb", "s_SiIIa_SiIIb",
    "HCD_damp1", "HCD_damp2", "HCD_damp3", "HCD_damp4", "HCD_const",
    "R_coeff", "ln_AGN", "ln_AGN_const",
]
(COL_K, COL_S, COL_T, COL_T1, COL_T2A, COL_T2BC, COL_T3A, COL_T3BC,
 COL_D1, COL_D2, COL_D3, COL_D4, COL_KT, COL_Q, COL_AGN, COL_SPARE) = range(NCOL)

EMU_CODES = {"Delta2_p": 0, "n_p": 1, "alpha_p": 2, "mF": 3, "sigT_Mpc": 4, "gamma": 5, "kF_Mpc": 6, "lambda_P": 7}

# c1d_field ids
FIELDS = [
    "PMIN", "PSCALE", "PRIOR_X0", "PRIOR_W", "ZBASE", "ZWGT", "ZIDX", "ZNULL", "ZOTYPE",
    "LINP", "MZ", "IGMFID", "EMU_LO", "EMU_HI", "HULL_ABC", "HULL_DIM", "ZOFF_F", "ZOFF_D",
    "TAB", "RB_START", "RB_WGT", "DATA", "ICOV_BLOCKS", "ICOV_FULL", "NN_W", "NN_B",
    "GP_X", "GP_INVELL", "GP_ALPHA", "GP_YMEAN", "GP_YSTD", "GP_NORM", "GP_NORM_MF", "GP_NORM_TAB", "ZMASK",
]
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}
INT_FIELDS = {"ZIDX", "ZOTYPE", "HULL_DIM", "ZOFF_F", "ZOFF_D", "RB_START", "ZMASK"}


def family_z_weights(fam, zs):
    """Weights W[z, j] with g_lin(z) = sum_j W[z, j] coeff_j for one family
    (before the output type), reproducing
      pivot       np.poly1d(coeff)(ln((1+z)/(1+z0)))       base_igm.py:248-251
      interp_lin  np.interp(z, znodes, coeff) (clamped)     :253-254
      interp_spl  make_interp_spline(znodes, coeff, k=1)(z) (linear extrapolation) :255-261
    and the rows forced to the null value by z_max
    (base_contaminants.py:192-199) as a boolean mask."""
    zs = np.asarray(zs, dtype=float)
    n = len(fam["coeffs"])
    Wz = np.zeros((len(zs), n))
    ztype = fam["ztype"]
    if ztype == "pivot":
        xz = np.log((1 + zs) / (1 + fam["z0"]))
        for j in range(n):
            Wz[:, j] = xz ** (n - 1 - j)
    elif ztype in ("interp_lin", "interp_spl"):
        nodes = np.asarray(fam["znodes"], dtype=float)
        if len(nodes) == 0 and n == 1:
            # a family switched off in the baseline (n_<name> = 0: no nodes, one fixed coefficient, e.g. R_coeff
            # under name_variation="no_contaminants", input_pipeline.py:1096-1098): the reference never
            # evaluates it, the constant stands in
            Wz[:, 0] = 1.0
            nodes = None
        elif len(nodes) != n:
            raise ValueError("znodes / coefficient count mismatch")
        if nodes is not None and n > 1 and np.any(np.diff(nodes) <= 0):
            # repeated nodes (a one-redshift data set): whatever the reference's routine returns for them is
            # linear in the coefficients, so its response to unit vectors gives the weights
            if ztype == "interp_spl":
                raise ValueError("interp_spl needs strictly increasing znodes (scipy.interpolate.make_interp_spline)")
            for j in range(n):
                Wz[:, j] = np.interp(zs, nodes, np.eye(n)[j])
            nodes = None
        for iz, z in enumerate(zs if nodes is not None else []):
            if n == 1:
                Wz[iz, 0] = 1.0
                continue
            if ztype == "interp_lin" and z <= nodes[0]:
                Wz[iz, 0] = 1.0
            elif ztype == "interp_lin" and z >= nodes[-1]:
                Wz[iz, -1] = 1.0
            else:
                j = int(np.clip(np.searchsorted(nodes, z, side="right") - 1, 0, n - 2))
                t = (z - nodes[j]) / (nodes[j + 1] - nodes[j])
                Wz[iz, j] = 1.0 - t
                Wz[iz, j + 1] = t
    else:
        raise NotImplementedError("ztype " + str(ztype))
    forced = np.zeros(len(zs), dtype=bool)
    if not np.isnan(fam["zmax"]):
        forced = zs >= fam["zmax"]
    return Wz, forced


def _neutral_family(value, otype):
    return {
        "ztype": "pivot", "otype": otype, "znodes": np.zeros(0), "coeffs": np.array([float(value)]),
        "param_idx": np.array([-1]), "z0": 3.0, "null": np.nan, "zmax": np.nan,
    }


def compile_z_functions(spec):
    """ZBASE/ZWGT/ZIDX/ZNULL/ZOTYPE for the NQ quantities."""
    zs = np.asarray(spec["zs"], dtype=float)
    nz = len(zs)
    fams = spec["families"]
    rows = []
    ell = 1
    for q in QUANTITIES:
        if q == "ln_AGN" and q in fams:
            # AGN_Model nulls on the constant coefficient, not on the value (AGN_model.py:85-86)
            fam = dict(fams[q])
            fam["null"] = np.nan
        elif q == "ln_AGN_const" and "ln_AGN" in fams:
            src = fams["ln_AGN"]
            fam = dict(src)
            fam.update(ztype="pivot", znodes=np.zeros(0), coeffs=np.asarray(src["coeffs"], dtype=float)[-1:],
                       param_idx=np.asarray(src["param_idx"])[-1:], zmax=np.nan, otype="exp")
        elif q == "HCD_damp1" and spec["cont"]["hcd_model"] == "mcdonald":
            # HCD_Model_McDonald2005 nulls on the constant coefficient, not on the value (hcd_model_McDonald2005.py:75-76)
            fam = dict(fams[q])
            fam["null"] = np.nan
        elif q == "HCD_damp2" and spec["cont"]["hcd_model"] == "mcdonald":
            # the gate: exp(constant coefficient), 0 when that coefficient is <= the null value
            src = fams["HCD_damp1"]
            fam = dict(src)
            fam.update(ztype="pivot", znodes=np.zeros(0), coeffs=np.asarray(src["coeffs"], dtype=float)[-1:],
                       param_idx=np.asarray(src["param_idx"])[-1:], zmax=np.nan, otype="exp")
        elif q in ("HCD_damp3", "HCD_damp4") and spec["cont"]["hcd_model"] == "mcdonald":
            fam = _neutral_family(0.0, "const")
        elif q == "R_coeff" and not spec["cont"]["apply_resolution"]:
            # the resolution term is applied only if some R_coeff is free (lya_theory.py:712-722)
            fam = _neutral_family(0.0, "const")
        elif q == "HCD_const" and spec["cont"]["hcd_model"] != "rogers":
            # HCD_BOSS has the one amplitude HCD_damp1 (hcd_boss.py:66-91): a stray HCD_const family has no effect
            fam = _neutral_family(0.0, "const")
        elif q in fams:
            fam = fams[q]
        elif q in ("tau_eff", "sigT_kms", "gamma", "kF_kms"):
            raise ValueError("missing IGM family " + q)
        else:
            # families a model variant does not have (SiVid, BOSS, AGN off): value 0
            fam = _neutral_family(0.0, "const")
        Wz, forced = family_z_weights(fam, zs)
        coeffs = np.asarray(fam["coeffs"], dtype=float)
        pidx = np.asarray(fam["param_idx"], dtype=np.int64)
        free = pidx >= 0
        base = Wz[:, ~free] @ coeffs[~free] if (~free).any() else np.zeros(nz)
        ent = []
        for iz in range(nz):
            if forced[iz]:
                ent.append([])
                base[iz] = fam["null"]
            else:
                e = [(int(pidx[j]), float(Wz[iz, j])) for j in np.nonzero(free)[0] if Wz[iz, j] != 0.0]
                ent.append(e)
                ell = max(ell, len(e))
        null = -np.inf if np.isnan(fam["null"]) else float(fam["null"])
        rows.append((base, ent, null, 1 if fam["otype"] == "exp" else 0))
    zbase = np.zeros((NQ, nz))
    zwgt = np.zeros((NQ, nz, ell))
    zidx = np.full((NQ, nz, ell), -1, dtype=np.int32)
    znull = np.zeros(NQ)
    zotype = np.zeros(NQ, dtype=np.int32)
    for iq, (base, ent, null, ot) in enumerate(rows):
        zbase[iq] = base
        znull[iq] = null
        zotype[iq] = ot
        for iz in range(nz):
            for e, (pi, wv) in enumerate(ent[iz]):
                zidx[iq, iz, e] = pi
                zwgt[iq, iz, e] = wv
    return zbase, zwgt, zidx, znull, zotype, ell


def compile_tables(spec):
    """The NCOL walker-independent columns over the packed model k grid."""
    zs = np.asarray(spec["zs"], dtype=float)
    kgrid = model_k_grid(spec)
    cont = spec["cont"]
    emu = spec["emulator"]
    M = np.asarray(spec["cosmo"]["M_of_z"], dtype=float)
    nf = int(sum(len(k) for k in kgrid))
    tab = np.zeros((NCOL, nf))
    sm, sa = cont["si_mult"], cont["si_add"]
    dv, rat, off = sm["dv"], sm["rat"], dict(sm["off"])
    # SiII-SiII cross terms are additive only (si_mult.py:203-206)
    c = rat["SiIIc_SiIII"] / rat["SiIIb_SiIII"]
    o = 0
    for iz, z in enumerate(zs):
        k = np.asarray(kgrid[iz], dtype=float)
        s = slice(o, o + len(k))
        o += len(k)
        tab[COL_K, s] = k
        k_Mpc = k * M[iz]
        if emu["kind"] == "nn":
            tab[COL_T, s] = np.log10(k_Mpc)
            scale = float(emu["yscale"]) * M[iz]
        else:
            tab[COL_T, s] = k_Mpc / float(emu["kmax_Mpc"])
            scale = M[iz]
            if emu.get("norm_table") is not None:
                # norm = np.interp(k_Mpc, input_norm["k_Mpc"], norm_imF(mF)) (Fig3a.py:83-85): the walker-independent
                # half of it, interval index c and weight t in the norm's k grid, packed as c + t; beyond the last node
                # c = n - 1, t = 0 reads the padded (repeated) last column
                kn = np.asarray(emu["norm_k_Mpc"], dtype=float)
                cidx = np.clip(np.searchsorted(kn, k_Mpc, side="right") - 1, 0, len(kn) - 1)
                hi = np.minimum(cidx + 1, len(kn) - 1)
                den = np.where(hi > cidx, kn[hi] - kn[cidx], 1.0)
                tt = np.clip(np.where(hi > cidx, (k_Mpc - kn[cidx]) / den, 0.0), 0.0, np.nextafter(1.0, 0.0))
                tt = np.where(k_Mpc <= kn[0], 0.0, tt)
                tab[COL_SPARE, s] = cidx + tt
        S = np.full(len(k), scale)
        if cont["ic_correction"]:
            # model_contaminants.py:310-330
            anc = (0.15261529 * z**2 - 2.30600644 * z + 2.61877894) * (1 - np.exp(-k / 0.003669741766936781))
            S = S * (1 / (1 - anc / 100))
        tab[COL_S, s] = S
        if cont["metal_model"] == "SiMult":
            tab[COL_T1, s] = off["SiIII_Lya"] * np.cos(dv["SiIII_Lya"] * k)
            tab[COL_T2A, s] = off["SiIIa_Lya"] * np.cos(dv["SiIIa_Lya"] * k)
            tab[COL_T2BC, s] = off["SiIIb_Lya"] * np.cos(dv["SiIIb_Lya"] * k) + off["SiIIc_Lya"] * c * np.cos(dv["SiIIc_Lya"] * k)
            tab[COL_T3A, s] = off["SiIII_SiIIa"] * np.cos(dv["SiIII_SiIIa"] * k)
            tab[COL_T3BC, s] = off["SiIII_SiIIb"] * np.cos(dv["SiIII_SiIIb"] * k) + off["SiIII_SiIIc"] * c * np.cos(dv["SiIII_SiIIc"] * k)
        else:
            tab[COL_T1, s] = np.cos(dv["SiIII_Lya"] * k)  # si_vid_final.py:212-216
        if cont["hcd_model"] == "rogers":
            rg = cont["rogers"]
            for it in range(4):
                a_z = rg["a0"][it] * ((1 + z) / (1 + rg["z0"])) ** rg["a1"][it]
                b_z = rg["b0"][it] * ((1 + z) / (1 + rg["z0"])) ** rg["b1"][it]
                tab[COL_D1 + it, s] = 1 / (a_z * np.exp(k * b_z) - 1) ** 2
        elif cont["hcd_model"] == "boss":
            tab[COL_D1, s] = 1 / (1 - (1 / (15000 * k - 8.9)))  # hcd_boss.py:5-7
        elif cont["hcd_model"] == "mcdonald":
            tab[COL_D1, s] = 0.018 + 1 / (15000 * k - 8.9)  # hcd_model_McDonald2005.py:91
        else:
            raise NotImplementedError("hcd model " + str(cont["hcd_model"]))
        # si_add.py:172-217
        soff, sdv, srat = sa["off"], sa["dv"], sa["rat"]
        rac, rbc, rab = srat["SiIIa_SiIIc"], srat["SiIIb_SiIIc"], srat["SiIIa_SiIIb"]
        dca, dcb, dba = sdv["SiIIc_SiIIa"], sdv["SiIIc_SiIIb"], sdv["SiIIb_SiIIa"]
        kt = (
            soff["SiIIc_SiIIa"] * (1 + rac**2 + 2 * rac * np.cos(dca * k))
            + soff["SiIIc_SiIIb"] * (1 + rbc**2 + 2 * rbc * np.cos(dcb * k))
            + soff["SiIIb_SiIIa"] * (rbc**2 * (1 + rab**2 + 2 * rab * np.cos(dba * k)))
            + soff["SiIIacbc"] * (2 * (1 + rac * rbc) * np.cos(0.5 * (dca - dcb) * k) + 2 * (rac + rbc) * np.cos(0.5 * (dca + dcb) * k))
            + soff["SiIIacab"] * (2 * rbc * (1 + rac * rab) * np.cos(0.5 * (dca - dba) * k) + 2 * rbc * (rac + rab) * np.cos(0.5 * (dca + dba) * k))
            + soff["SiIIbcab"] * (2 * rbc * (1 + rbc * rab) * np.cos(0.5 * (dcb - dba) * k) + 2 * rbc * (rbc + rab) * np.cos(0.5 * (dcb + dba) * k))
        )
        tab[COL_KT, s] = kt
        Rz = C_KMS * 0.8 / (1 + z) / LYA_AA  # resolution_class.py:26-34
        tab[COL_Q, s] = 2 * Rz**2 * k**2
        if cont.get("agn") is not None:
            tab[COL_AGN, s] = agn_delta(cont["agn"], z, k)
    return tab


def agn_delta(agn, z, k):
    """delta(z, k) of the AGN feedback template: a + b exp(-c k) tabulated at 9 redshifts,
    linear in z between them and extrapolated from the first two rows above the table
    (AGN_model.py:99-116; table data/nuisance/AGN_corr.dat)"""
    az = np.asarray(agn["z"], dtype=float)
    ex = np.asarray(agn["expansion"], dtype=float)
    if z <= np.max(az):
        yy = ex[:, 0][None, :] + ex[:, 1][None, :] * np.exp(-ex[:, 2][None, :] * k[:, None])
        order = np.argsort(az)
        out = np.zeros(len(k))
        for i in range(len(k)):
            out[i] = np.interp(z, az[order], yy[i, order])
        return out
    up = ex[0, 0] + ex[0, 1] * np.exp(-ex[0, 2] * k)
    lo = ex[1, 0] + ex[1, 1] * np.exp(-ex[1, 2] * k)
    return (up - lo) / (az[0] - az[1]) * (z - az[0]) + up


def compile_rebin(spec):
    """Banded (ELL) form of the rebinning operator cover/sum_cover
    (likelihood.py:83-106, 147-161); identity when rebin_k == 1."""
    d = spec["data"]
    nd = int(sum(len(k) for k in d["k_kms"]))
    rb = spec["rebin"]
    if rb is None:
        start = np.concatenate([np.arange(len(k), dtype=np.int32) for k in d["k_kms"]])
        return start.astype(np.int32), np.ones((nd, 1)), 1
    starts, spans = [], []
    for iz in range(len(d["k_kms"])):
        cover = np.asarray(rb["cover"][iz], dtype=float)
        for j in range(cover.shape[0]):
            nzc = np.nonzero(cover[j])[0]
            if len(nzc) == 0:
                starts.append(0)
                spans.append(1)
            else:
                starts.append(int(nzc[0]))
                spans.append(int(nzc[-1] - nzc[0] + 1))
    width = max(spans)
    wgt = np.zeros((nd, width))
    o = 0
    for iz in range(len(d["k_kms"])):
        cover = np.asarray(rb["cover"][iz], dtype=float)
        sc = np.asarray(rb["sum_cover"][iz], dtype=float)
        nfz = cover.shape[1]
        for j in range(cover.shape[0]):
            st = starts[o + j]
            seg = cover[j, st : min(st + width, nfz)] / sc[j]
            wgt[o + j, : len(seg)] = seg
        o += cover.shape[0]
    return np.asarray(starts, dtype=np.int32), wgt, width


def compile_spec(spec):
    """Returns (config: dict of scalars for c1d_config, fields: name -> ndarray)"""
    nz, ndim, nd, nf = spec_sizes(spec)
    p = spec["params"]
    cosmo = spec["cosmo"]
    cont = spec["cont"]
    emu = spec["emulator"]
    names = list(p["names"])
    pmin = np.asarray(p["min"], dtype=float)
    pmax = np.asarray(p["max"], dtype=float)
    fields = {}
    fields["PMIN"] = pmin
    fields["PSCALE"] = pmax - pmin
    fields["PRIOR_X0"] = (np.asarray(p["value"], dtype=float) - pmin) / (pmax - pmin)
    gs = p["gauss_sigma_cube"]
    has_gauss = gs is not None
    fields["PRIOR_W"] = 1.0 / (2.0 * np.asarray(gs, dtype=float) ** 2) if has_gauss else np.zeros(ndim)

    zbase, zwgt, zidx, znull, zotype, ell = compile_z_functions(spec)
    fields.update(ZBASE=zbase, ZWGT=zwgt, ZIDX=zidx, ZNULL=znull, ZOTYPE=zotype)
    fields["LINP"] = np.asarray(cosmo["linP_fid"], dtype=float)
    fields["MZ"] = np.asarray(cosmo["M_of_z"], dtype=float)
    fields["IGMFID"] = np.stack([np.asarray(spec["igm_fid"][k], dtype=float) for k in ("tau_eff", "sigT_kms", "gamma", "kF_kms")])

    emu_params = list(emu["emu_params"])
    n_emu = len(emu_params)
    if n_emu > MAX_EMU:
        raise ValueError("too many emulator inputs")
    fields["EMU_LO"] = np.asarray(emu["xmin"], dtype=float)
    fields["EMU_HI"] = np.asarray(emu["xmax"], dtype=float)
    emu_code = [EMU_CODES[n] for n in emu_params] + [0] * (MAX_EMU - n_emu)

    hull = spec["hull"]
    if hull is not None:
        abc, dims = [], []
        for h in hull["hulls"]:
            d0 = emu_params.index(hull["params"][h["dim0"]])
            d1 = emu_params.index(hull["params"][h["dim1"]])
            A = np.asarray(h["A"], dtype=float)
            b = np.asarray(h["b"], dtype=float)
            for f in range(A.shape[0]):
                abc.append((A[f, 0], A[f, 1], b[f]))
                dims.append((d0, d1))
        fields["HULL_ABC"] = np.asarray(abc, dtype=float).reshape(-1, 3)
        fields["HULL_DIM"] = np.asarray(dims, dtype=np.int32).reshape(-1, 2)
        n_facets = len(abc)
        hull_tol = float(hull["tol"])
    else:
        fields["HULL_ABC"] = np.zeros((0, 3))
        fields["HULL_DIM"] = np.zeros((0, 2), dtype=np.int32)
        n_facets, hull_tol = 0, 0.0

    kgrid = model_k_grid(spec)
    d = spec["data"]
    fields["ZOFF_F"] = np.concatenate([[0], np.cumsum([len(k) for k in kgrid])]).astype(np.int32)
    fields["ZOFF_D"] = np.concatenate([[0], np.cumsum([len(k) for k in d["k_kms"]])]).astype(np.int32)
    fields["TAB"] = compile_tables(spec)
    start, wgt, rwidth = compile_rebin(spec)
    fields["RB_START"], fields["RB_WGT"] = start, wgt
    fields["DATA"] = np.concatenate([np.asarray(x, dtype=float) for x in d["Pk_kms"]])
    fields["ICOV_BLOCKS"] = np.concatenate([np.asarray(c, dtype=float).reshape(-1) for c in d["icov"]])
    has_full = d["full_icov"] is not None
    if has_full:
        if not np.array_equal(np.asarray(d["full_Pk_kms"], dtype=float), fields["DATA"]):
            raise ValueError("full_Pk_kms differs from the concatenated per-z data vector")
        # d^T A d only sees the symmetric part of A (np.linalg.inv output is symmetric to rounding only); the chi2
        # kernel visits one triangle of the matrix and doubles it
        A = np.asarray(d["full_icov"], dtype=float)
        fields["ICOV_FULL"] = 0.5 * (A + A.T)
    else:
        fields["ICOV_FULL"] = np.zeros(0)

    # evenly spaced model grids allow exp(-s k) by recurrence on the device
    uniform = True
    for k in kgrid:
        k = np.asarray(k, dtype=float)
        if len(k) < 3:
            uniform = False
            break
        lin = k[0] + np.arange(len(k)) * ((k[-1] - k[0]) / (len(k) - 1))
        if np.max(np.abs(k - lin)) > 1e-13 * np.abs(k[-1]):
            uniform = False
            break
    cfg = {
        "uniform_k": int(uniform),
        "has_agn": int(cont.get("agn") is not None and "ln_AGN" in spec["families"]),
        "hcd_gate": int(cont["hcd_model"] == "mcdonald"),
        "nz": nz, "ndim": ndim, "n_emu": n_emu, "nd": nd, "nf": nf,
        "ell_width": int(ell), "rebin_width": int(rwidth),
        "n_layers": 0, "layer_dims": [0] * (MAX_LAYERS + 1),
        "gp_ntrain": 0, "gp_nnorm": 0, "gp_imf": 0, "gp_nm": 0, "gp_nkn": 0,
        "metal_model": 0 if cont["metal_model"] == "SiMult" else 1,
        "apply_resolution": int(bool(cont["apply_resolution"])),
        "has_full": int(has_full), "has_gauss": int(has_gauss), "n_hull_facets": int(n_facets),
        "idx_As": names.index("As") if "As" in names else -1,
        "idx_ns": names.index("ns") if "ns" in names else -1,
        "idx_nrun": names.index("nrun") if "nrun" in names else -1,
        "emu_code": emu_code,
        "As_fid": float(cosmo["As_fid"]), "ns_fid": float(cosmo["ns_fid"]), "nrun_fid": float(cosmo["nrun_fid"]),
        "L_emu": float(np.log(cosmo["kp_emu_Mpc"] / cosmo["ks_Mpc"])),
        "L_star": float(np.log(cosmo["kp_star_Mpc"] / cosmo["ks_Mpc"])),
        "blob_fid": [float(v) for v in cosmo["blob_fid"]],
        "star_lo": [-np.inf] * 3, "star_hi": [np.inf] * 3,
        "min_log_like": float(p["min_log_like"]),
        "hull_tol": hull_tol, "nn_slope": 0.0, "gp_sigma2": 0.0,
    }
    sp = cosmo["star_priors"]
    if sp is not None:
        for i in range(3):
            if not np.isnan(sp[i, 0]):
                cfg["star_lo"][i] = float(sp[i, 0])
                cfg["star_hi"][i] = float(sp[i, 1])
    rat, off = cont["si_mult"]["rat"], cont["si_mult"]["off"]
    cc = rat["SiIIc_SiIII"] / rat["SiIIb_SiIII"]
    cfg.update(
        rb3=float(rat["SiIIb_SiIII"]),
        ra3_over_rb3=float(rat["SiIIa_SiIII"] / rat["SiIIb_SiIII"]),
        c0_o1=float(off["SiIII_Lya"]),
        c0_o2=float(off["SiIIa_Lya"]),
        c0_o3c=float(off["SiIIb_Lya"] + cc**2 * off["SiIIc_Lya"]),
    )

    for nm in ("NN_W", "NN_B", "GP_X", "GP_INVELL", "GP_ALPHA", "GP_YMEAN", "GP_YSTD", "GP_NORM", "GP_NORM_MF", "GP_NORM_TAB"):
        fields[nm] = np.zeros(0)
    if emu["kind"] == "nn":
        layers = emu["layers"]
        if len(layers) > MAX_LAYERS:
            raise ValueError("too many MLP layers")
        dims = [n_emu] + [int(np.asarray(l["W"]).shape[0]) for l in layers]
        for il, l in enumerate(layers):
            if np.asarray(l["W"]).shape != (dims[il + 1], dims[il]):
                raise ValueError("layer %d shape mismatch" % il)
        nc = dims[-1]
        if nc != int(emu["ndeg"]) + 1 or nc > MAX_COEF:
            raise ValueError("polynomial degree / head width mismatch")
        cfg.update(emu_kind=0, nc=nc, n_layers=len(layers), nn_slope=float(emu["slope"]))
        cfg["layer_dims"] = dims + [0] * (MAX_LAYERS + 1 - len(dims))
        # in-major [K][N] so that consecutive output columns are contiguous
        fields["NN_W"] = np.concatenate([np.asarray(l["W"], dtype=float).T.reshape(-1) for l in layers])
        fields["NN_B"] = np.concatenate([np.asarray(l["b"], dtype=float).reshape(-1) for l in layers])
    elif emu["kind"] == "gp":
        alpha = np.asarray(emu["alpha"], dtype=float)
        nc = alpha.shape[1]
        if nc > MAX_COEF:
            raise ValueError("too many GP outputs")
        ncoef = np.asarray(emu.get("norm_coef", np.zeros(0)), dtype=float).reshape(-1)
        cfg.update(
            emu_kind=1, nc=nc, gp_ntrain=int(alpha.shape[0]), gp_nnorm=len(ncoef),
            gp_imf=emu_params.index("mF"), gp_sigma2=float(emu["sigma2"]),
        )
        if emu.get("norm_table") is not None:
            if len(ncoef):
                raise ValueError("GP emulator: give either norm_coef or norm_table")
            nm_, T = np.asarray(emu["norm_mF"], dtype=float), np.asarray(emu["norm_table"], dtype=float)
            kn = np.asarray(emu["norm_k_Mpc"], dtype=float)
            if T.shape != (len(nm_), len(kn)) or len(nm_) < 2 or len(kn) < 2 or np.any(np.diff(nm_) <= 0) or np.any(np.diff(kn) <= 0):
                raise ValueError("GP emulator: norm_table must be [len(norm_mF), len(norm_k_Mpc)] over increasing nodes")
            cfg.update(gp_nm=len(nm_), gp_nkn=len(kn) + 1)
            fields["GP_NORM_MF"] = nm_
            fields["GP_NORM_TAB"] = np.concatenate([T, T[:, -1:]], axis=1)  # last column repeated, see COL_SPARE
        fields["GP_X"] = np.asarray(emu["X_train"], dtype=float)
        fields["GP_INVELL"] = 1.0 / np.asarray(emu["lengthscales"], dtype=float)
        fields["GP_ALPHA"] = alpha
        fields["GP_YMEAN"] = np.asarray(emu["ymean"], dtype=float)
        fields["GP_YSTD"] = np.asarray(emu["ystd"], dtype=float)
        fields["GP_NORM"] = ncoef
    else:
        raise ValueError("unknown emulator kind")

    out = {}
    for nm, arr in fields.items():
        dt = np.int32 if nm in INT_FIELDS else np.float64
        out[nm] = np.ascontiguousarray(arr, dtype=dt)
    return cfg, out
