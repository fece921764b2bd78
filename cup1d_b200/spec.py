"""Neutral, serialisable description ("model spec") of one cup1d P1D likelihood.

The spec is the seam between the reference's object graph and the B200
context: :func:`spec_from_likelihood` walks any object exposing the
reference ``Likelihood`` attributes (SURVEY.md §8b "Construction") and
flattens it to plain numpy arrays / scalars / strings.  The context compiler
(:mod:`cup1d_b200.compile`) lowers a spec to device tables; the test oracle
(``oracle/``) evaluates the same spec on the CPU.

Nothing here touches the GPU and nothing here imports the oracle.

Spec layout (all floats are float64)::

  version      int
  zs           [Nz]
  params       names[ndim], min, max, value, gauss_sigma_cube (or None), min_log_like
  cosmo        As_fid ns_fid nrun_fid ks_Mpc kp_emu_Mpc kp_star_Mpc
               linP_fid[Nz,3] M_of_z[Nz] blob_fid[6] star_priors[3,2] (nan = open)
  families     {name: ztype otype znodes coeffs param_idx z0 null zmax}
  igm_fid      {tau_eff, sigT_kms, gamma, kF_kms: [Nz]}
  cont         metal_model hcd_model ic_correction apply_resolution
               si_mult{dv,rat,off} si_add{dv,rat,off} rogers{a0,a1,b0,b1,z0}
  data         k_kms[list] Pk_kms[list] icov[list] full_Pk_kms full_icov
  rebin        None | rebin_k k_kms[list] cover[list] sum_cover[list]
  hull         None | params[list] tol hulls[list of dim0 dim1 A b]
  emulator     kind ("nn"|"gp") emu_params[...] + arrays
"""

from __future__ import annotations

import numpy as np

SPEC_VERSION = 1

# coefficient families the hot path knows about, with the model that owns them
# (reference: cup1d/igm/*.py list_coeffs; cup1d/contaminants/*.py list_coeffs)
IGM_FAMILIES = ("tau_eff", "sigT_kms", "gamma", "kF_kms")
SIMULT_FAMILIES = (
    "f_Lya_SiIII",
    "s_Lya_SiIII",
    "f_Lya_SiII",
    "s_Lya_SiII",
    "f_SiIIa_SiIII",
    "f_SiIIb_SiIII",
)
SIADD_FAMILIES = ("f_SiIIa_SiIIb", "s_SiIIa_SiIIb")
HCD_FAMILIES = ("HCD_damp1", "HCD_damp2", "HCD_damp3", "HCD_damp4", "HCD_const")
RES_FAMILIES = ("R_coeff",)
AGN_FAMILIES = ("ln_AGN",)


# --------------------------------------------------------------------------
# (de)serialisation: nested dict/list of arrays <-> flat npz
# --------------------------------------------------------------------------
def _flatten(obj, prefix, out):
    if isinstance(obj, dict):
        out[prefix + "/__dict__"] = np.array(sorted(obj.keys()), dtype=str)
        for k, v in obj.items():
            _flatten(v, prefix + "/" + str(k), out)
    elif isinstance(obj, (list, tuple)):
        out[prefix + "/__list__"] = np.array(len(obj))
        for i, v in enumerate(obj):
            _flatten(v, prefix + "/" + str(i), out)
    elif obj is None:
        out[prefix + "/__none__"] = np.array(0)
    elif isinstance(obj, str):
        out[prefix + "/__str__"] = np.array(obj)
    elif isinstance(obj, (bool, np.bool_)):
        out[prefix + "/__bool__"] = np.array(bool(obj))
    elif isinstance(obj, (int, np.integer)):
        out[prefix + "/__int__"] = np.array(int(obj))
    elif isinstance(obj, (float, np.floating)):
        out[prefix + "/__float__"] = np.array(float(obj))
    else:
        out[prefix] = np.asarray(obj)


def _unflatten(flat, prefix):
    if prefix + "/__dict__" in flat:
        return {
            str(k): _unflatten(flat, prefix + "/" + str(k))
            for k in flat[prefix + "/__dict__"]
        }
    if prefix + "/__list__" in flat:
        n = int(flat[prefix + "/__list__"])
        return [_unflatten(flat, prefix + "/" + str(i)) for i in range(n)]
    if prefix + "/__none__" in flat:
        return None
    if prefix + "/__str__" in flat:
        return str(flat[prefix + "/__str__"])
    if prefix + "/__bool__" in flat:
        return bool(flat[prefix + "/__bool__"])
    if prefix + "/__int__" in flat:
        return int(flat[prefix + "/__int__"])
    if prefix + "/__float__" in flat:
        return float(flat[prefix + "/__float__"])
    return np.array(flat[prefix])


def save_spec(path, spec, **extra):
    """Write a spec (plus optional extra arrays, e.g. golden outputs) to .npz"""
    flat = {}
    _flatten(spec, "spec", flat)
    for k, v in extra.items():
        _flatten(v, "extra/" + k, flat)
    np.savez_compressed(path, **flat)


def load_spec(path, with_extra=False):
    with np.load(path, allow_pickle=False) as f:
        flat = {k: f[k] for k in f.files}
    spec = _unflatten(flat, "spec")
    if not with_extra:
        return spec
    names = sorted(
        {k.split("/")[1] for k in flat if k.startswith("extra/")}
    )
    extra = {n: _unflatten(flat, "extra/" + n) for n in names}
    return spec, extra


# --------------------------------------------------------------------------
# helpers on specs
# --------------------------------------------------------------------------
def spec_sizes(spec):
    """Return (Nz, ndim, N_d, N_f) of a spec"""
    nz = len(spec["zs"])
    ndim = len(spec["params"]["names"])
    nd = int(sum(len(k) for k in spec["data"]["k_kms"]))
    if spec["rebin"] is None:
        nf = nd
    else:
        nf = int(sum(len(k) for k in spec["rebin"]["k_kms"]))
    return nz, ndim, nd, nf


def model_k_grid(spec):
    """k grid (per z) on which the theory is evaluated: the fine grid when
    rebin_k != 1 (likelihood.py:743-748), else the data grid."""
    if spec["rebin"] is None:
        return spec["data"]["k_kms"]
    return spec["rebin"]["k_kms"]


# --------------------------------------------------------------------------
# extraction from reference-API objects
# --------------------------------------------------------------------------
def _family_from_model(model, name, free_names, is_igm):
    """One coefficient family of an IGM_model / Contaminant.

    Mirrors the slot bookkeeping of ``get_coeff`` (base_igm.py:289-319,
    base_contaminants.py:215-245): parameters are matched by the substring
    ``name + "_"`` and parameter ``name_i`` feeds ``coeff[-(i+1)]`` for a
    ``pivot`` family and ``coeff[i]`` otherwise.
    """
    ztype = model.prop_coeffs[name + "_ztype"]
    otype = model.prop_coeffs[name + "_otype"]
    coeffs = np.array(model.coeffs[name], dtype=float).reshape(-1)
    ncoef = len(coeffs)
    matched = [p for p in free_names if (name + "_") in p]
    param_idx = np.full(ncoef, -1, dtype=np.int64)
    if len(matched) > 0:
        n_pars = model.n_pars[name] if hasattr(model, "n_pars") else len(matched)
        if len(matched) != n_pars:
            raise ValueError("number of params mismatch for: " + name)
        for ii in range(len(matched)):
            pname = name + "_" + str(ii)
            if pname not in free_names:
                raise ValueError("free parameter not found: " + pname)
            slot = ncoef - (ii + 1) if ztype == "pivot" else ii
            param_idx[slot] = free_names.index(pname)
    if ztype.startswith("interp"):
        znodes = np.array(model.prop_coeffs[name + "_znodes"], dtype=float)
    else:
        znodes = np.zeros(0)
    null = np.nan
    zmax = np.nan
    if not is_igm:
        nv = getattr(model, "null_vals", None)
        if nv is not None and name in nv:
            null = float(nv[name])
        zm = getattr(model, "z_max", None)
        if zm is not None:
            zmax = float(zm[name])
            if np.isnan(null):
                raise ValueError("z_max without null value for " + name)
    return {
        "ztype": str(ztype),
        "otype": str(otype),
        "znodes": znodes,
        "coeffs": coeffs,
        "param_idx": param_idx,
        "z0": float(model.z_0),
        "null": null,
        "zmax": zmax,
    }


def _plain(d):
    return {str(k): float(v) for k, v in d.items()}


def emulator_to_spec(emulator):
    """Flatten an emulator object.  Objects that know how to describe
    themselves (``to_spec``) are used as is; a LaCE ``NNEmulator`` is read
    through its public attributes [recalled layout, see SURVEY.md §2.2]."""
    if hasattr(emulator, "to_spec"):
        return emulator.to_spec()
    if hasattr(emulator, "nn") and hasattr(emulator, "paramLims"):
        import torch  # LaCE NN models are torch modules

        layers = []
        mods = [m for m in emulator.nn.modules() if isinstance(m, torch.nn.Linear)]
        # means head only (covariance return is commented out, lya_theory.py:691-709)
        names = [n for n, m in emulator.nn.named_modules() if isinstance(m, torch.nn.Linear)]
        for n, m in zip(names, mods):
            if "std" in n.lower():
                continue
            layers.append(
                {
                    "W": m.weight.detach().cpu().double().numpy(),
                    "b": m.bias.detach().cpu().double().numpy(),
                }
            )
        lims = np.asarray(emulator.paramLims, dtype=float)
        return {
            "kind": "nn",
            "emu_params": [str(p) for p in emulator.emu_params],
            "xmin": lims[:, 0],
            "xmax": lims[:, 1],
            "layers": layers,
            "slope": 0.5,
            "ndeg": int(emulator.ndeg),
            "yscale": float(np.asarray(emulator.yscalings)),
            "kmax_Mpc": float(emulator.kmax_Mpc),
            "kp_Mpc": float(emulator.kp_Mpc),
        }
    gp = _gp_emulator_to_spec(emulator)
    if gp is not None:
        return gp
    raise NotImplementedError(
        "cannot flatten emulator of type %s: give it a to_spec() method"
        % type(emulator).__name__
    )


def _find_gprs(emulator):
    """the fitted scikit-learn GaussianProcessRegressor(s) an emulator object holds: one multi-output regressor, or a
    list / tuple / dict of single-output ones (one per polynomial coefficient, in coefficient order)"""
    try:
        from sklearn.gaussian_process import GaussianProcessRegressor as GPR
    except Exception:  # noqa: BLE001
        return None
    found = []
    for name, v in sorted(vars(emulator).items()):
        if isinstance(v, GPR):
            found.append((name, [v]))
        elif isinstance(v, (list, tuple)) and len(v) and all(isinstance(g, GPR) for g in v):
            found.append((name, list(v)))
        elif isinstance(v, dict) and len(v) and all(isinstance(g, GPR) for g in v.values()):
            found.append((name, [v[k] for k in sorted(v)]))
    if not found:
        return None
    if len(found) > 1:
        raise NotImplementedError("emulator holds several Gaussian-process attributes (%s): give it a to_spec() method" % [n for n, _ in found])
    return found[0][1]


def _rbf_of_kernel(kernel):
    """(sigma2, lengthscales) of a fitted kernel of the form [ConstantKernel *] RBF [+ WhiteKernel]; anything else is
    rejected by name (the device kernel evaluates an ARD squared exponential only)"""
    from sklearn.gaussian_process import kernels as K

    k = kernel
    if isinstance(k, K.Sum):
        parts = [q for q in (k.k1, k.k2) if not isinstance(q, K.WhiteKernel)]
        if len(parts) != 1:
            raise NotImplementedError("unsupported GP kernel %r (sums other than kernel + WhiteKernel)" % (kernel,))
        k = parts[0]  # the white-noise term does not enter k(x*, X_train) of the predictive mean
    sigma2 = 1.0
    if isinstance(k, K.Product):
        consts = [q for q in (k.k1, k.k2) if isinstance(q, K.ConstantKernel)]
        rest = [q for q in (k.k1, k.k2) if not isinstance(q, K.ConstantKernel)]
        if len(consts) != 1 or len(rest) != 1:
            raise NotImplementedError("unsupported GP kernel %r" % (kernel,))
        sigma2 = float(consts[0].constant_value)
        k = rest[0]
    if type(k) is not K.RBF:  # Matern and others derive from RBF
        raise NotImplementedError("unsupported GP kernel %r: only [Constant *] RBF [+ White] is evaluated on the device" % (kernel,))
    return sigma2, np.atleast_1d(np.asarray(k.length_scale, dtype=float))


def _input_limits(emulator, names):
    """[n_in, 2] limits of the unit-box scaling of the emulator inputs"""
    for attr in ("param_lims", "paramLims", "param_limits"):
        if hasattr(emulator, attr):
            lims = np.asarray(getattr(emulator, attr), dtype=float)
            if lims.shape == (len(names), 2):
                return lims
    if hasattr(emulator, "xmin") and hasattr(emulator, "xmax"):
        return np.stack([np.asarray(emulator.xmin, dtype=float), np.asarray(emulator.xmax, dtype=float)], axis=1)
    norm = getattr(emulator, "input_norm", None)
    if isinstance(norm, dict) and all(n in norm for n in names):
        out = []
        for n in names:
            v = norm[n]
            out.append([float(v["min"]), float(v["max"])] if isinstance(v, dict) else [float(v[0]), float(v[1])])
        return np.asarray(out)
    raise NotImplementedError("GP emulator: cannot find the limits of the input scaling (param_lims / paramLims [n_in, 2], xmin / xmax, "
                              "or input_norm[<parameter>] = (min, max))")


def _poly_order(func_poly, n_out):
    """which power of x each argument of func_poly(x, *coef) multiplies (the device evaluates sum_o c_o x^o)"""
    x = np.array([0.3, 0.7, 1.3])
    zero = np.asarray(func_poly(x, *([0.0] * n_out)), dtype=float)
    powers = []
    for o in range(n_out):
        e = [0.0] * n_out
        e[o] = 1.0
        b = np.asarray(func_poly(x, *e), dtype=float) - zero
        ok = [p for p in range(n_out) if np.allclose(b, x**p, rtol=1e-12, atol=1e-14)]
        if len(ok) != 1:
            raise NotImplementedError("GP emulator: func_poly is not a plain polynomial in its arguments")
        powers.append(ok[0])
    if sorted(powers) != list(range(n_out)) or np.any(np.abs(zero) > 1e-14):
        raise NotImplementedError("GP emulator: func_poly is not a plain polynomial in its arguments")
    return powers


def _gp_emulator_to_spec(emulator):
    """A LaCE-style GP emulator (the reference's default, ``CH24_mpgcen_gpr``: input_pipeline.py:21,
    set_emulator.py:4-38) read through what the reference itself uses of it - ``emu_params``, ``kp_Mpc``, ``kmax_Mpc``,
    ``input_norm["k_Mpc"]``, ``norm_imF(mF)``, ``func_poly`` (notebooks/figures_CM26/Fig3a.py:75-88) - and through its
    fitted scikit-learn ``GaussianProcessRegressor`` (``X_train_``, ``alpha_``, ``kernel_``, the target normalisation).
    Returns None when the object holds no such regressor.  [LaCE's source is absent: attribute names beyond the ones
    the reference reads are looked up generically; anything that cannot be represented is rejected loudly.]"""
    gprs = _find_gprs(emulator)
    if gprs is None:
        return None
    names = [str(p) for p in emulator.emu_params]
    X = np.asarray(gprs[0].X_train_, dtype=float)
    if X.shape[1] != len(names):
        raise NotImplementedError("GP emulator: X_train_ has %d columns for %d emulator parameters" % (X.shape[1], len(names)))
    sig, ell = [], []
    alphas, ymean, ystd = [], [], []
    for g in gprs:
        if np.asarray(g.X_train_).shape != X.shape or not np.array_equal(np.asarray(g.X_train_, dtype=float), X):
            raise NotImplementedError("GP emulator: the regressors do not share one training set")
        s2, l = _rbf_of_kernel(g.kernel_)
        sig.append(s2)
        ell.append(np.broadcast_to(l, (X.shape[1],)).astype(float))
        a = np.asarray(g.alpha_, dtype=float)
        a = a[:, None] if a.ndim == 1 else a
        alphas.append(a)
        ymean.append(np.broadcast_to(np.atleast_1d(np.asarray(getattr(g, "_y_train_mean", 0.0), dtype=float)), (a.shape[1],)))
        ystd.append(np.broadcast_to(np.atleast_1d(np.asarray(getattr(g, "_y_train_std", 1.0), dtype=float)), (a.shape[1],)))
    if any(not np.array_equal(l, ell[0]) for l in ell):
        raise NotImplementedError("GP emulator: per-output length scales (one kernel-vector product per output) are not supported")
    # different amplitudes per regressor fold into alpha (the posterior mean is linear in sigma2 * alpha)
    alpha = np.concatenate([a * (s / sig[0]) for a, s in zip(alphas, sig)], axis=1)
    ymean, ystd = np.concatenate(ymean), np.concatenate(ystd)
    n_out = alpha.shape[1]
    powers = _poly_order(emulator.func_poly, n_out)
    order = np.argsort(powers)  # device coefficient o multiplies x^o
    lims = _input_limits(emulator, names)
    # the regressor works on the unit box of the limits: check it on the training set itself
    if X.min() < -1e-6 or X.max() > 1 + 1e-6:
        raise NotImplementedError("GP emulator: X_train_ is not in the unit box of the input limits")
    kn = np.asarray(emulator.input_norm["k_Mpc"], dtype=float)
    f = emulator.norm_imF
    if hasattr(f, "x") and hasattr(f, "y") and str(getattr(f, "_kind", "linear")) == "linear" and np.ndim(f.y) == 2:
        # a scipy interp1d over mean flux: its own nodes, exactly
        mF = np.asarray(f.x, dtype=float)
        T = np.asarray(f.y, dtype=float)
        T = T if T.shape == (len(mF), len(kn)) else T.T
    else:
        # any other callable: tabulated on a fine mean-flux grid (read linearly in between)
        imf = names.index("mF")
        mF = np.linspace(lims[imf, 0], lims[imf, 1], 1025)
        T = np.stack([np.asarray(f(m), dtype=float).reshape(-1) for m in mF])
    if T.shape != (len(mF), len(kn)):
        raise NotImplementedError("GP emulator: norm_imF(mF) does not return one value per input_norm['k_Mpc'] node")
    return {
        "kind": "gp",
        "emu_params": names,
        "xmin": lims[:, 0].copy(),
        "xmax": lims[:, 1].copy(),
        "X_train": X,
        "lengthscales": ell[0].copy(),
        "sigma2": float(sig[0]),
        "alpha": alpha[:, order],
        "ymean": ymean[order],
        "ystd": ystd[order],
        "norm_coef": np.zeros(0),
        "norm_mF": mF,
        "norm_k_Mpc": kn,
        "norm_table": T,
        "kmax_Mpc": float(emulator.kmax_Mpc),
        "kp_Mpc": float(emulator.kp_Mpc),
    }


def spec_from_likelihood(like):
    """Flatten a reference-API ``Likelihood`` (cup1d/likelihood/likelihood.py)
    into a model spec.  Every walker-independent number is read from, or
    evaluated with, the reference objects themselves."""
    theory = like.theory
    data = like.data
    zs = np.array(data.z, dtype=float)
    nz = len(zs)
    free = like.free_params
    free_names = [p.name for p in free]
    for nm in free_names:
        if nm in ("ombh2", "omch2", "H0", "mnu", "cosmomc_theta"):
            raise NotImplementedError(
                "parameter %s changes the background expansion (CAMB per "
                "walker, lya_theory.py:433-443): outside the hot path" % nm
            )

    # ---- parameters and priors (likelihood.py:163-186, 1049-1064) ----
    gp = like.Gauss_priors
    params = {
        "names": free_names,
        "min": np.array([p.min_value for p in free], dtype=float),
        "max": np.array([p.max_value for p in free], dtype=float),
        "value": np.array([p.value for p in free], dtype=float),
        "gauss_sigma_cube": None if gp is None else np.array(gp, dtype=float),
        "min_log_like": float(like.min_log_like),
    }

    # ---- cosmology constants (lya_theory.py:270-351, 536-584) ----
    fid = theory.fid_cosmo
    cm = fid["cosmo"]
    init = cm.cosmo.InitPower
    fzs = np.asarray(fid["zs"], dtype=float)
    linP = np.zeros((nz, 3))
    M_of_z = np.zeros(nz)
    for iz, z in enumerate(zs):
        j = np.argwhere(fzs == z)[0, 0]
        lp = fid["linP_Mpc_params"][j]
        linP[iz] = (lp["Delta2_p"], lp["n_p"], lp["alpha_p"])
        M_of_z[iz] = fid["M_of_zs"][j]
    pstar = cm.get_linP_params()
    blob_fid = np.array(
        [
            pstar["Delta2_star"],
            pstar["n_star"],
            pstar["alpha_star"],
            pstar["f_star"],
            pstar["g_star"],
            cm.cosmo.H0,
        ],
        dtype=float,
    )
    star = np.full((3, 2), np.nan)
    sp = getattr(theory, "star_priors", None)
    if sp is not None:
        for i, key in enumerate(("Delta2_star", "n_star", "alpha_star")):
            if key in sp:
                star[i] = (sp[key][0], sp[key][1])
    cosmo = {
        "As_fid": float(init.As),
        "ns_fid": float(init.ns),
        "nrun_fid": float(init.nrun),
        "ks_Mpc": float(init.pivot_scalar),
        "kp_emu_Mpc": float(theory.emu_kp_Mpc),
        "kp_star_Mpc": float(theory.kp_kms * cm.dkms_dMpc(theory.z_star)),
        "linP_fid": linP,
        "M_of_z": M_of_z,
        "blob_fid": blob_fid,
        "star_priors": None if sp is None else star,
    }

    # ---- coefficient families ----
    families = {}
    igm_fid = {}
    igm_models = theory.model_igm.models
    for mkey in ("F_model", "T_model", "P_model"):
        model = igm_models[mkey]
        for name in model.list_coeffs:
            ztype = model.prop_coeffs[name + "_ztype"]
            if ztype.endswith("_smspl"):
                raise NotImplementedError(
                    "interp_smspl solves a smoothing-spline system per call "
                    "(base_igm.py:262-266): not supported on the device"
                )
            families[name] = _family_from_model(model, name, free_names, True)
            igm_fid[name] = np.array(model.fid_interp[name](zs), dtype=float)

    mc = theory.model_cont
    si_mult = mc.metal_models["Si_mult"]
    si_add = mc.metal_models["Si_add"]
    hcd = mc.hcd_model
    res = theory.model_syst.resolution_model
    cname = type(si_mult).__name__
    if cname == "SiVid":
        metal_model = "SiVid"
    elif cname == "SiMult":
        metal_model = "SiMult"
    else:
        raise NotImplementedError("metal model " + cname)
    hname = type(hcd).__name__
    if hname == "HCD_Model_Rogers":
        hcd_model = "rogers"
    elif hname == "HCD_BOSS":
        hcd_model = "boss"
    elif hname == "HCD_Model_McDonald2005":
        hcd_model = "mcdonald"
    else:
        raise NotImplementedError("HCD model " + hname)
    if hcd_model == "mcdonald":
        # not a Contaminant: pivot polynomial in ln_A_damp_i with parameter i feeding coefficient [n-1-i]
        # (hcd_model_McDonald2005.py:41-63, 94-131), A_damp = exp(poly(ln((1+z)/(1+z_0)))), nulled on the constant coefficient
        coeffs = np.array(hcd.ln_A_damp_coeff, dtype=float)
        n = len(coeffs)
        pidx = np.full(n, -1, dtype=np.int64)
        matched = [p for p in free_names if "ln_A_damp" in p]
        if matched and len(matched) != n:
            raise ValueError("number of params mismatch in get_A_damp_coeffs")  # hcd_model_McDonald2005.py:117-121
        for i in range(n):
            nm = "ln_A_damp_" + str(i)
            if nm in free_names:
                pidx[n - 1 - i] = free_names.index(nm)
        families["HCD_damp1"] = {
            "ztype": "pivot", "otype": "exp", "znodes": np.zeros(0), "coeffs": coeffs, "param_idx": pidx,
            "z0": float(hcd.z_0), "null": float(hcd.null_value), "zmax": np.nan,
        }
    for model in (si_mult, si_add, res) + (() if hcd_model == "mcdonald" else (hcd,)):
        for name in model.list_coeffs:
            if model.prop_coeffs[name + "_ztype"].endswith("_smspl"):
                raise NotImplementedError("interp_smspl not supported: " + name)
            families[name] = _family_from_model(model, name, free_names, False)

    cont = {
        "metal_model": metal_model,
        "hcd_model": hcd_model,
        "ic_correction": bool(mc.ic_correction),
        # lya_theory.py:712-722: systematics only if an R_coeff parameter is free
        "apply_resolution": any(n.startswith("R_coeff") for n in free_names),
        "si_mult": {
            "dv": _plain(si_mult.dv),
            "rat": _plain(si_mult.rat),
            "off": _plain(si_mult.off),
        },
        "si_add": {
            "dv": _plain(si_add.dv),
            "rat": _plain(si_add.rat),
            "off": _plain(si_add.off),
        },
    }
    agn_model = getattr(mc, "agn_model", None)
    cont["agn"] = None
    if agn_model is not None:
        # AGN_Model is not a Contaminant: pivot polynomial in ln_AGN_i with parameter i
        # feeding coefficient [n-1-i] (AGN_model.py:52-72, 137-167)
        coeffs = np.array(agn_model.ln_AGN_coeff, dtype=float)
        n = len(coeffs)
        pidx = np.full(n, -1, dtype=np.int64)
        for i in range(n):
            nm = "ln_AGN_" + str(i)
            if nm in free_names:
                pidx[n - 1 - i] = free_names.index(nm)
        families["ln_AGN"] = {
            "ztype": "pivot", "otype": "exp", "znodes": np.zeros(0), "coeffs": coeffs, "param_idx": pidx,
            "z0": float(agn_model.z_0), "null": float(agn_model.null_value), "zmax": np.nan,
        }
        cont["agn"] = {"z": np.array(agn_model.AGN_z, dtype=float), "expansion": np.array(agn_model.AGN_expansion, dtype=float)}
    if hcd_model == "rogers":
        cont["rogers"] = {
            "a0": np.array(hcd.a_0, dtype=float),
            "a1": np.array(hcd.a_1, dtype=float),
            "b0": np.array(hcd.b_0, dtype=float),
            "b1": np.array(hcd.b_1, dtype=float),
            # the base-class ctor overwrites z_0 after the subclass sets it
            # (hcd_model_rogers_class.py:80 vs base_contaminants.py:24)
            "z0": float(hcd.z_0),
        }

    # ---- data and metric (likelihood.py:916-963) ----
    if like.extra_data is not None:
        raise NotImplementedError("extra (high-resolution) data set")
    dspec = {
        "k_kms": [np.array(k, dtype=float) for k in data.k_kms],
        "Pk_kms": [np.array(p, dtype=float) for p in data.Pk_kms],
        "icov": [np.array(c, dtype=float) for c in like.icov_Pk_kms],
        "full_Pk_kms": None,
        "full_icov": None,
    }
    if like.full_icov_Pk_kms is not None:
        dspec["full_Pk_kms"] = np.array(data.full_Pk_kms, dtype=float)
        dspec["full_icov"] = np.array(like.full_icov_Pk_kms, dtype=float)

    if like.args.rebin_k != 1:
        rebin = {
            "rebin_k": int(like.args.rebin_k),
            "k_kms": [np.array(k, dtype=float) for k in like.rebin["k_kms"]],
            "cover": [np.array(c, dtype=float) for c in like.rebin["cover"]],
            "sum_cover": [np.array(c, dtype=float) for c in like.rebin["sum_cover"]],
        }
    else:
        rebin = None

    # ---- hull prior (hull.py:121-165) ----
    hull = None
    if getattr(theory, "use_hull", False):
        h = theory.hull
        hull = {
            "params": [str(p) for p in h.params],
            "tol": float(h.tol),
            "hulls": [
                {
                    "dim0": int(hh.dim0),
                    "dim1": int(hh.dim1),
                    "A": np.array(hh.eq, dtype=float),
                    "b": np.array(hh.eq2[:, 0], dtype=float),
                }
                for hh in h.hulls
            ],
        }

    return {
        "version": SPEC_VERSION,
        "zs": zs,
        "params": params,
        "cosmo": cosmo,
        "families": families,
        "igm_fid": igm_fid,
        "cont": cont,
        "data": dspec,
        "rebin": rebin,
        "hull": hull,
        "emulator": emulator_to_spec(theory.emulator),
    }
