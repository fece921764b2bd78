"""Synthetic, seeded workloads for the P1D likelihood hot path.

LaCE, CAMB, the DESI DR1 FITS files and the trained emulators are not
available offline, so the benchmark and the GPU parity tests run
on *synthetic mock P1D data of the reference's shapes* (SURVEY.md §8d configs
C1-C5).  Everything is generated from ``np.random.default_rng(seed)`` and the
seeds are part of the config.  The builders return model specs
(cup1d_b200/spec.py); mock data vectors are produced by evaluating the model at
the truth point (as ``Mock_P1D`` does, cup1d/p1ds/mock_data.py:96-100) with
whatever evaluator the caller passes (the B200 context in the product, the
oracle in tests).
"""

from __future__ import annotations

import numpy as np

from .spec import SPEC_VERSION

C_KMS = 299792.458

# atomic data of the lines (same physical constants as si_mult.py:32-44)
WAV = {"SiIII": 1206.51, "SiIIc": 1260.42, "SiIIb": 1193.28, "SiIIa": 1190.42, "Lya": 1215.67}
OSC = {"SiIII": 1.67, "SiIIc": 1.22, "SiIIb": 0.575, "SiIIa": 0.277}

# DESI-DR1 QMLE3 k-grid shape (data/zenodo/fig_1.npy of the reference)
DR1_NK = (49, 52, 55, 58, 61, 63, 65, 67, 69, 70, 72)


def _dv(a, b):
    return abs(np.log(WAV[b] / WAV[a])) * C_KMS


def _rs(a, b):
    return (WAV[a] * OSC[a]) / (WAV[b] * OSC[b])


def silicon_constants():
    """dv / rat / off tables of the SiMult and SiAdd models
    (si_mult.py:46-103, si_add.py:44-83)"""
    pairs = [
        ("SiIII", "Lya"), ("SiIIc", "Lya"), ("SiIIb", "Lya"), ("SiIIa", "Lya"),
        ("SiIII", "SiIIc"), ("SiIII", "SiIIb"), ("SiIII", "SiIIa"),
        ("SiIIc", "SiIIb"), ("SiIIc", "SiIIa"), ("SiIIb", "SiIIa"),
    ]
    dv = {a + "_" + b: float(_dv(a, b)) for a, b in pairs}
    si_mult = {
        "dv": dv,
        "rat": {
            "SiIIa_SiIII": _rs("SiIIa", "SiIII"),
            "SiIIb_SiIII": _rs("SiIIb", "SiIII"),
            "SiIIc_SiIII": _rs("SiIIc", "SiIII"),
            "SiIIa_SiIIc": _rs("SiIIa", "SiIIc"),
            "SiIIb_SiIIc": _rs("SiIIb", "SiIIc"),
        },
        "off": {
            "SiIII_Lya": 1.0, "SiIIa_Lya": 1.0, "SiIIb_Lya": 1.0, "SiIIc_Lya": 0.0,
            "SiIII_SiIIa": 1.0, "SiIII_SiIIb": 1.0, "SiIII_SiIIc": 0.0,
            "SiIIc_SiIIb": 0.0, "SiIIc_SiIIa": 0.0, "SiIIb_SiIIa": 0.0,
        },
    }
    si_add = {
        "dv": {k: dv[k] for k in ("SiIIc_SiIIb", "SiIIc_SiIIa", "SiIIb_SiIIa")},
        "rat": {
            "SiIIa_SiIIc": _rs("SiIIa", "SiIIc"),
            "SiIIb_SiIIc": _rs("SiIIb", "SiIIc"),
            "SiIIa_SiIIb": _rs("SiIIa", "SiIIb"),
        },
        "off": {
            "SiIIc_SiIIb": 0.0, "SiIIc_SiIIa": 0.0, "SiIIb_SiIIa": 1.0,
            "SiIIacbc": 0.0, "SiIIacab": 0.0, "SiIIbcab": 0.0,
        },
    }
    return si_mult, si_add


ROGERS = {
    # hcd_model_rogers_class.py:76-79; effective pivot z0 = 3
    "a0": np.array([2.2001, 1.5083, 1.1415, 0.8633]),
    "a1": np.array([0.0134, 0.0994, 0.0937, 0.2943]),
    "b0": np.array([36.449, 81.388, 162.95, 429.58]),
    "b1": np.array([-0.0674, -0.2287, 0.0126, -0.4964]),
    "z0": 3.0,
}

# AGN feedback correction table of Chabanier et al. 2020 (rows: z, a, b, c with
# delta = a + b exp(-c k)); same numbers as the reference's data/nuisance/AGN_corr.dat
AGN_TABLE = np.array([
    [4.254242298971526, 0.001948707260760598, -0.005113025363647983, 297.76336326532356],
    [3.7399508480460177, -0.002395192935503769, -0.014496865270097973, 189.05445743949753],
    [3.33497418318008, -0.0015888346971373486, -0.0280564876368884, 166.29319089261327],
    [3.0053240496893006, -0.0066211056484300895, -0.03933342912116104, 159.0177616872931],
    [2.7310610558920296, -0.007901357724366368, -0.04622024793203778, 153.92538415800533],
    [2.4979310724428525, -0.011606137378406078, -0.05222578599582309, 151.0360486518846],
    [2.29696817249267, -0.01397956154253014, -0.058065380798205515, 138.5389744852627],
    [2.1217615536671572, -0.019499510059870485, -0.061213371048927705, 124.88130335751495],
    [1.966806396137914, -0.025420244243385796, -0.06663923678784489, 120.3535253491937],
])

EMU_RANGES = {
    "Delta2_p": (0.10, 1.00),
    "n_p": (-2.60, -2.00),
    "alpha_p": (-0.32, -0.10),
    "mF": (0.02, 0.98),
    "sigT_Mpc": (0.03, 0.30),
    "gamma": (0.6, 2.6),
    "kF_Mpc": (3.0, 30.0),
}
MPG_PARAMS = ["Delta2_p", "n_p", "mF", "sigT_Mpc", "gamma", "kF_Mpc"]
NYX_PARAMS = ["Delta2_p", "n_p", "alpha_p", "mF", "sigT_Mpc", "gamma", "kF_Mpc"]


# ---------------------------------------------------------------------------
# emulators
# ---------------------------------------------------------------------------
def synth_nn_emulator(seed=11, suite="mpg", nhidden=5, max_neurons=250, ndeg=5, head=50):
    """Seeded weights of a LaCE-shaped polynomial-coefficient MLP (SURVEY.md
    §2.2): Linear(n_in,10), widths linspace(10, max_neurons, nhidden), a
    ``means`` head Linear(max,head)+Linear(head, ndeg+1); LeakyReLU(0.5)
    between layers.  Weights are float32-representable, as torch would hold
    them."""
    rng = np.random.default_rng(seed)
    emu_params = MPG_PARAMS if suite == "mpg" else NYX_PARAMS
    n_in = len(emu_params)
    widths = [int(w) for w in np.linspace(10, max_neurons, nhidden).astype(int)]
    dims = [n_in] + widths + [head, ndeg + 1]
    gain = np.sqrt(2.0 / (1.0 + 0.25))
    layers = []
    for il in range(len(dims) - 1):
        fan_in, fan_out = dims[il], dims[il + 1]
        last = il == len(dims) - 2
        scale = (0.25 if last else gain) / np.sqrt(fan_in)
        W = rng.normal(0, scale, (fan_out, fan_in))
        b = rng.normal(0, 0.05, fan_out)
        if last:
            base = np.zeros(ndeg + 1)
            base[: min(5, ndeg + 1)] = [-0.75, -0.85, -0.45, -0.05, 0.02][: min(5, ndeg + 1)]
            b = base + rng.normal(0, 0.01, fan_out)
            W[3:] *= 0.2  # keep the high-order coefficients tame
        layers.append(
            {
                "W": W.astype(np.float32).astype(np.float64),
                "b": b.astype(np.float32).astype(np.float64),
            }
        )
    return {
        "kind": "nn",
        "emu_params": list(emu_params),
        "xmin": np.array([EMU_RANGES[p][0] for p in emu_params]),
        "xmax": np.array([EMU_RANGES[p][1] for p in emu_params]),
        "layers": layers,
        "slope": 0.5,
        "ndeg": int(ndeg),
        "yscale": 1.0,
        "kmax_Mpc": 4.0,
        "kp_Mpc": 0.7,
    }


def synth_gp_emulator(seed=12, suite="mpg", n_train=1320, n_out=5, norm="ktable", n_norm_mF=24, n_norm_k=48):
    """Seeded GP restatement: ARD squared-exponential kernel on unit-box
    inputs, alpha = K^-1 y for a smooth synthetic target.  norm="ktable": the norm of the
    reference's own uses, a table over (mF, k_Mpc) read with np.interp in k (Fig3a.py:83-85); norm="poly": the
    k-independent ln norm(mF) polynomial of the first fixtures."""
    rng = np.random.default_rng(seed)
    emu_params = MPG_PARAMS if suite == "mpg" else NYX_PARAMS
    n_in = len(emu_params)
    X = rng.uniform(0.05, 0.95, (n_train, n_in))
    ell = rng.uniform(0.6, 1.6, n_in)
    sigma2 = 1.0
    # smooth target: random quadratic form per output
    A = rng.normal(0, 0.6, (n_out, n_in))
    B = rng.normal(0, 0.3, (n_out, n_in, n_in))
    Xc = X - 0.5
    y = Xc @ A.T + np.einsum("ni,oij,nj->no", Xc, B, Xc)
    d = (X[:, None, :] - X[None, :, :]) / ell
    K = sigma2 * np.exp(-0.5 * np.sum(d * d, axis=2)) + 1e-6 * np.eye(n_train)
    alpha = np.linalg.solve(K, y)
    ymean = np.zeros(n_out)
    ymean[: min(5, n_out)] = [-0.55, -3.2, 1.5, -0.6, 0.1][: min(5, n_out)]
    ystd = np.array([0.25, 0.4, 0.3, 0.15, 0.05, 0.02, 0.01][:n_out])
    spec = {
        "kind": "gp",
        "emu_params": list(emu_params),
        "xmin": np.array([EMU_RANGES[p][0] for p in emu_params]),
        "xmax": np.array([EMU_RANGES[p][1] for p in emu_params]),
        "X_train": X,
        "lengthscales": ell,
        "sigma2": sigma2,
        "alpha": alpha,
        "ymean": ymean,
        "ystd": ystd,
        "kmax_Mpc": 4.0,
        "kp_Mpc": 0.7,
    }
    if norm == "poly":
        spec["norm_coef"] = np.array([0.3, -1.1, 0.4])  # ln norm(mF) = c0 + c1 mF + c2 mF^2
        return spec
    # a median-P1D-like shape: falls with k, steeper for high mean flux; unevenly spaced nodes in both directions
    lo, hi = EMU_RANGES["mF"]
    mF = lo + (hi - lo) * np.sort(np.concatenate([[0.0, 1.0], rng.uniform(0, 1, n_norm_mF - 2)]))
    kn = np.geomspace(0.03, 4.2, n_norm_k) * (1.0 + 0.02 * rng.uniform(-1, 1, n_norm_k))
    kn = np.sort(kn)
    lnn = (0.3 - 1.1 * mF[:, None] + 0.4 * mF[:, None] ** 2) - (0.35 + 0.5 * mF[:, None]) * np.log1p(kn[None, :] / 0.8) \
        + 0.05 * np.sin(3.0 * np.log(kn[None, :]) + 2.0 * mF[:, None])
    spec["norm_coef"] = np.zeros(0)
    spec["norm_mF"] = mF
    spec["norm_k_Mpc"] = kn
    spec["norm_table"] = np.exp(lnn)
    return spec


# ---------------------------------------------------------------------------
# background, histories, grids
# ---------------------------------------------------------------------------
def hubble_over_1pz(z, H0=67.66, Om=0.3111):
    z = np.asarray(z, dtype=float)
    return H0 * np.sqrt(Om * (1 + z) ** 3 + 1 - Om) / (1 + z)


def fid_histories(z):
    z = np.asarray(z, dtype=float)
    return {
        "tau_eff": 0.0025 * (1 + z) ** 3.7,
        "sigT_kms": 9.1 * np.sqrt(1.1 - 0.08 * (z - 3)),
        "gamma": 1.45 + 0.05 * (z - 3),
        "kF_kms": 0.105 + 0.012 * (z - 3),
    }


def dr1_k_grid(nk, dk=5e-4):
    """k0 = 0.00125, dk = 5e-4 up to 0.03 s/km, then log-spaced steps of
    dlnk = 0.023026 (shape of data/zenodo/fig_1.npy); a coarser ``dk`` gives
    the same k range with fewer bins (small golden fixtures)"""
    edges = [0.001]
    while len(edges) < nk + 1:
        if edges[-1] < 0.03 - 1e-9:
            edges.append(edges[-1] + dk)
        else:
            edges.append(edges[-1] * np.exp(0.023026 * dk / 5e-4))
    edges = np.array(edges)
    kmin, kmax = edges[:-1], edges[1:]
    return 0.5 * (kmin + kmax), kmin, kmax


def chabanier_k_grid(nk=35):
    """35 linear bins of width 5.42e-4 s/km starting at 1.084e-3
    (data/p1d_measurements/Chabanier2019/Pk1D_data.dat)"""
    k = 0.001084 + 0.000542 * np.arange(nk)
    return k, k - 0.000271, k + 0.000271


def bin_coverage(xmin_o, xmax_o, xmin_n, xmax_n):
    """fraction of each fine cell inside each data bin (likelihood.py:27-37)"""
    lo = np.maximum(xmin_o[None, :], xmin_n[:, None])
    hi = np.minimum(xmax_o[None, :], xmax_n[:, None])
    return np.maximum((hi - lo) / (xmax_o - xmin_o)[None, :], 0.0)


def rebin_tables(kmin, kmax, rebin_k):
    """fine grid + coverage of one z-bin (likelihood.py:83-106)"""
    nelem = len(kmin) * rebin_k
    kf = np.linspace(kmin[0] * 0.95, kmax[-1] * 1.05, nelem)
    dk = kf[1] - kf[0]
    cover = bin_coverage(kf - 0.5 * dk, kf + 0.5 * dk, kmin, kmax)
    return kf, cover, cover.sum(axis=1)


# ---------------------------------------------------------------------------
# parameterisation ("global_opt"-like, input_pipeline.py:815-1110)
# ---------------------------------------------------------------------------
def _interp_family(nodes, value, lo, hi, names, pmin, pmax, pval, name, otype, z0=3.0, null=np.nan, free=True, rng=None, jitter=0.0):
    n = len(nodes)
    coeffs = np.full(n, float(value))
    idx = np.full(n, -1, dtype=np.int64)
    if free:
        for i in range(n):
            v = float(value)
            if rng is not None and jitter > 0:
                v = float(np.clip(v + rng.normal(0, jitter), lo + 0.05 * (hi - lo), hi - 0.05 * (hi - lo)))
            coeffs[i] = v
            idx[i] = len(names)
            names.append("%s_%d" % (name, i))
            pmin.append(lo)
            pmax.append(hi)
            pval.append(v)
    if n >= 2:
        ztype, znodes = "interp_lin", np.asarray(nodes, dtype=float)
    else:
        # a single coefficient is a 'pivot' family (input_pipeline.py:1041-1044)
        ztype, znodes = "pivot", np.zeros(0)
    return {
        "ztype": ztype,
        "otype": otype,
        "znodes": znodes,
        "coeffs": coeffs,
        "param_idx": idx,
        "z0": float(z0),
        "null": float(null),
        "zmax": np.nan,
    }


def _fixed_pivot(value, otype, null=np.nan):
    return {
        "ztype": "pivot",
        "otype": otype,
        "znodes": np.zeros(0),
        "coeffs": np.array([0.0, float(value)]),
        "param_idx": np.array([-1, -1], dtype=np.int64),
        "z0": 3.0,
        "null": float(null),
        "zmax": np.nan,
    }


def build_spec(
    emulator,
    grid="dr1",
    rebin_k=8,
    full_cov=True,
    nz=11,
    nz_igm=4,
    vary_nrun=False,
    resolution=True,
    hcd_const=False,
    ic_correction=False,
    gauss_priors=False,
    agn=False,
    seed=14,
    snr=36.0,
    nk=None,
):
    """A model spec with the reference's baseline structure and synthetic
    numbers.  ``data.Pk_kms`` is left at zero: fill it with
    :func:`attach_mock_data`."""
    rng = np.random.default_rng(seed)
    zs = np.round(2.2 + 0.2 * np.arange(nz), 10)
    z_min, z_max = zs[0], zs[-1]
    names, pmin, pmax, pval = [], [], [], []

    # cosmology (priors as Theory.set_cosmo_priors would give, order of magnitude)
    As_fid, ns_fid, nrun_fid = 2.1e-9, 0.9665, 0.0
    names += ["As", "ns"]
    pmin += [1.0e-9, 0.80]
    pmax += [3.6e-9, 1.12]
    pval += [As_fid, ns_fid]
    if vary_nrun:
        names.append("nrun")
        pmin.append(-0.08)
        pmax.append(0.08)
        pval.append(nrun_fid)

    fam = {}
    nodes_igm = np.geomspace(z_min, z_max, nz_igm)
    for name, otype, val, lo, hi in (
        ("tau_eff", "exp", 0.0, -0.45, 0.45),
        ("sigT_kms", "const", 1.0, 0.55, 1.55),
        ("gamma", "const", 1.0, 0.70, 1.35),
        ("kF_kms", "const", 1.0, 0.60, 1.50),
    ):
        fam[name] = _interp_family(nodes_igm, val, lo, hi, names, pmin, pmax, pval, name, otype, rng=rng, jitter=0.02)
    nodes2 = np.geomspace(z_min, z_max, 2)
    # flat priors input_pipeline.py:366-448, centres from data/ics/mpg_ic_global_red.npy scale
    for name, val, lo, hi, null in (
        ("f_Lya_SiIII", -4.17, -6.0, -2.0, -20.0),
        ("s_Lya_SiIII", 4.90, 2.0, 7.0, -20.0),
        ("f_Lya_SiII", -3.66, -6.0, -2.0, -20.0),
        ("s_Lya_SiII", 5.65, 2.0, 7.0, -20.0),
        ("f_SiIIa_SiIIb", 0.40, -3.0, 3.0, -20.0),
        ("s_SiIIa_SiIIb", 4.43, 0.0, 7.5, -20.0),
        ("f_SiIIa_SiIII", 0.84, -1.0, 3.0, -20.0),
        ("f_SiIIb_SiIII", 0.59, -1.0, 5.0, -20.0),
        ("HCD_damp1", -1.27, -10.0, -0.03, -21.5),
        ("HCD_damp2", -5.0, -10.0, -1.0, -21.5),
        ("HCD_damp3", -4.0, -10.0, -1.0, -21.5),
        ("HCD_damp4", -4.5, -10.0, -1.0, -21.5),
    ):
        fam[name] = _interp_family(nodes2, val, lo, hi, names, pmin, pmax, pval, name, "exp", null=null, rng=rng, jitter=0.1)
    if hcd_const:
        fam["HCD_const"] = _interp_family([3.0], 0.0, -0.2, 0.2, names, pmin, pmax, pval, "HCD_const", "const")
        fam["HCD_const"]["coeffs"] = np.array([0.0])
    else:
        fam["HCD_const"] = _fixed_pivot(0.0, "const")
    if resolution:
        fam["R_coeff"] = _interp_family(zs, 0.0, -0.1, 0.1, names, pmin, pmax, pval, "R_coeff", "const", rng=rng, jitter=0.01)
    else:
        fam["R_coeff"] = _fixed_pivot(0.0, "const")

    if agn:
        # optional AGN feedback amplitude, one pivot parameter (AGN_model.py:52-72: prior [-5, 1])
        fam["ln_AGN"] = _interp_family([3.0], -1.5, -5.0, 1.0, names, pmin, pmax, pval, "ln_AGN", "exp", null=-5.5)

    ndim = len(names)
    M = hubble_over_1pz(zs)
    hist = fid_histories(zs)
    kp_star = 0.009 * float(hubble_over_1pz(3.0))
    linP = np.stack([0.35 * (4.0 / (1 + zs)) ** 2, np.full(nz, -2.30), np.full(nz, -0.215)], axis=1)

    # k grids, rebinning, covariance skeleton
    k_kms, kmins, kmaxs = [], [], []
    for iz in range(nz):
        if grid == "dr1":
            n = DR1_NK[iz % len(DR1_NK)] if nk is None else nk
            k, kmin, kmax = dr1_k_grid(n)
        else:
            k, kmin, kmax = chabanier_k_grid(35 if nk is None else nk)
        k_kms.append(k)
        kmins.append(kmin)
        kmaxs.append(kmax)
    if rebin_k != 1:
        rb = {"rebin_k": int(rebin_k), "k_kms": [], "cover": [], "sum_cover": []}
        for iz in range(nz):
            kf, cover, sc = rebin_tables(kmins[iz], kmaxs[iz], rebin_k)
            rb["k_kms"].append(kf)
            rb["cover"].append(cover)
            rb["sum_cover"].append(sc)
    else:
        rb = None

    si_mult, si_add = silicon_constants()
    spec = {
        "version": SPEC_VERSION,
        "zs": zs,
        "params": {
            "names": names,
            "min": np.array(pmin, dtype=float),
            "max": np.array(pmax, dtype=float),
            "value": np.array(pval, dtype=float),
            "gauss_sigma_cube": (np.where(np.arange(ndim) % 3 == 0, 0.35, 1e4) if gauss_priors else None),
            "min_log_like": -1e100,
        },
        "cosmo": {
            "As_fid": As_fid,
            "ns_fid": ns_fid,
            "nrun_fid": nrun_fid,
            "ks_Mpc": 0.05,
            "kp_emu_Mpc": float(emulator["kp_Mpc"]),
            "kp_star_Mpc": kp_star,
            "linP_fid": linP,
            "M_of_z": M,
            "blob_fid": np.array([0.35, -2.30, -0.215, 0.981, 0.968, 67.66]),
            "star_priors": None,
        },
        "families": fam,
        "igm_fid": {k: np.asarray(v, dtype=float) for k, v in hist.items()},
        "cont": {
            "metal_model": "SiMult",
            "hcd_model": "rogers",
            "ic_correction": bool(ic_correction),
            "apply_resolution": bool(resolution),
            "si_mult": si_mult,
            "si_add": si_add,
            "rogers": {k: (v.copy() if hasattr(v, "copy") else v) for k, v in ROGERS.items()},
            "agn": ({"z": AGN_TABLE[:, 0].copy(), "expansion": AGN_TABLE[:, 1:].copy()} if agn else None),
        },
        "data": {
            "k_kms": k_kms,
            "Pk_kms": [np.zeros(len(k)) for k in k_kms],
            "icov": [np.eye(len(k)) for k in k_kms],
            "full_Pk_kms": np.zeros(sum(len(k) for k in k_kms)) if full_cov else None,
            "full_icov": np.eye(sum(len(k) for k in k_kms)) if full_cov else None,
        },
        "rebin": rb,
        "hull": None,
        "emulator": emulator,
    }
    spec["_synth"] = {"seed": int(seed), "snr": float(snr), "k_min": kmins, "k_max": kmaxs}
    return spec


def use_mcdonald_hcd(spec, n_hcd=2):
    """Switch a built spec to HCD_Model_McDonald2005 (hcd_model_McDonald2005.py): the two HCD_damp1 parameters
    become ln_A_damp_0 (constant coefficient, prior [-7, 2.5] so that walkers cross the null value -6) and, with
    ``n_hcd=2``, ln_A_damp_1 (slope in ln((1+z)/(1+z_0)), held fixed otherwise); the other Rogers amplitudes stay in the
    parameter vector without effect."""
    fam = spec["families"]
    old = fam["HCD_damp1"]
    i0, i1 = int(old["param_idx"][0]), int(old["param_idx"][1])
    p = spec["params"]
    p["min"][i0], p["max"][i0], p["value"][i0] = -7.0, 2.5, -1.2
    p["min"][i1], p["max"][i1], p["value"][i1] = -3.0, 3.0, 0.7
    fam["HCD_damp1"] = {
        "ztype": "pivot", "otype": "exp", "znodes": np.zeros(0),
        "coeffs": np.array([0.7, -1.2]) if n_hcd == 2 else np.array([-1.2]),
        "param_idx": np.array([i1, i0], dtype=np.int64) if n_hcd == 2 else np.array([i0], dtype=np.int64),
        "z0": 3.0, "null": -6.0, "zmax": np.nan,
    }
    spec["cont"]["hcd_model"] = "mcdonald"
    return spec


def truth_point(spec):
    """cube coordinates of the fiducial parameter values"""
    p = spec["params"]
    return (p["value"] - p["min"]) / (p["max"] - p["min"])


def attach_mock_data(spec, p_truth, seed_cov=14, seed_noise=15, add_noise=True, corr_amp=0.02):
    """Fill the data vector and the metric from the model at the truth.

    Covariance: Sigma = D rho D, D = P_truth/snr,
    rho = I + A A^T normalised to unit diagonal, A ~ N(0, corr_amp^2); the
    per-z blocks are the diagonal blocks of Sigma.  ``full_icov`` is used when
    the spec was built with ``full_cov=True`` (likelihood.py:953-963), and a
    Gaussian noise realisation is added (base_p1d_mock.py:57-75) so that
    pulls are O(1)."""
    snr = spec["_synth"]["snr"]
    p_truth = np.asarray(p_truth, dtype=float)
    nd = len(p_truth)
    rng = np.random.default_rng(seed_cov)
    A = rng.normal(0, corr_amp, (nd, nd))
    rho = np.eye(nd) + A @ A.T
    dg = np.sqrt(np.diag(rho))
    rho = rho / dg[:, None] / dg[None, :]
    D = np.abs(p_truth) / snr
    cov = rho * D[:, None] * D[None, :]
    data = p_truth.copy()
    if add_noise:
        rng2 = np.random.default_rng(seed_noise)
        L = np.linalg.cholesky(cov)
        data = data + L @ rng2.normal(0, 1, nd)
    d = spec["data"]
    off = 0
    for iz, k in enumerate(d["k_kms"]):
        n = len(k)
        d["Pk_kms"][iz] = data[off : off + n].copy()
        d["icov"][iz] = np.linalg.inv(cov[off : off + n, off : off + n])
        off += n
    if d["full_icov"] is not None:
        d["full_Pk_kms"] = data.copy()
        d["full_icov"] = np.linalg.inv(cov)
    return spec


def walkers_ball(spec, n, seed=0, rms=0.01):
    """Initial ensemble as Fitter.get_initial_walkers (fitter.py:746-767):
    pini + U(0,1)*rms, out-of-cube entries moved to 0.95 / 0.05"""
    rng = np.random.default_rng(seed)
    p0 = truth_point(spec)[None, :] + rng.random((n, len(spec["params"]["names"]))) * rms
    p0[p0 >= 1.0] = 0.95
    p0[p0 <= 0.0] = 0.05
    return p0


def walkers_sweep(spec, n, seed=16, width=0.08, frac_out=0.01):
    """Throughput-sweep ensemble: Gaussian scatter of
    ``width`` around the truth clipped to [0.02, 0.98], plus a fraction of
    rows deliberately outside the unit cube."""
    rng = np.random.default_rng(seed)
    ndim = len(spec["params"]["names"])
    x = truth_point(spec)[None, :] + rng.normal(0, width, (n, ndim))
    x = np.clip(x, 0.02, 0.98)
    nout = int(round(frac_out * n))
    if nout > 0:
        rows = rng.choice(n, nout, replace=False)
        cols = rng.integers(0, ndim, nout)
        x[rows, cols] = np.where(rng.random(nout) < 0.5, -0.01, 1.01)
    return x
