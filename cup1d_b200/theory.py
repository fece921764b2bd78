"""Drop-in for ``cup1d.likelihood.lya_theory.Theory.get_p1d_kms`` (lya_theory.py:610-814):
the model P1D on any subset of the compiled redshift bins and on any k grid, for any
parameter values.

The device tables of a context are tied to one (z, k) grid, so a request on another grid
compiles a derived context (the likelihood's spec with the redshift-dependent arrays
sub-selected, the requested k bins as "data" and no rebinning) and keeps it in a small
cache keyed by the grid; the evaluation itself is the same ``c1d_p1d_batch`` the
likelihood uses.  The reference's conventions are kept:

* ``like_params`` is a list of objects with ``name`` and ``value``
  (``LikelihoodParameter``); a parameter that is not in the list keeps the value held by
  its model (the fixed coefficient, or the fiducial cosmology) — base_igm.py:289-319,
  base_contaminants.py:215-245, lya_theory.py:281-307.  A flat array of ``ndim`` values in
  the order of the free parameters (``B200Likelihood.parameters_from_sampling_point``) and
  a ``{name: value}`` dict are accepted too;
* the resolution term is applied only when some ``R_coeff*`` parameter is in the list
  (lya_theory.py:712-722);
* there is no unit-cube test here (it lives in ``Likelihood``, likelihood.py:989-994);
* ``None`` when the star-prior box or (``apply_hull``) the convex hulls reject the point
  (:647-670); otherwise ``[p1ds(, blob)(, emu_call)]``, a bare list of per-z arrays when
  nothing else is requested; ``return_blob`` defaults to True as in the reference.

``remove`` (a dict of metal-term switches, si_mult.py:185-206, si_add.py:156-159) compiles a derived
context with those switches.  Not available on the device path: ``return_covar``, ``hires`` and
``return_contaminants`` (plot-only).
"""

from __future__ import annotations

from collections import OrderedDict

import numpy as np

BLOB_NAMES = ("Delta2_star", "n_star", "alpha_star", "f_star", "g_star", "H0")


def grid_spec(spec, iz_sel, k_kms, apply_resolution, remove=None, keep_data=False):
    """The likelihood's spec restricted to the redshift bins ``iz_sel`` with ``k_kms`` (one array
    per bin) as the evaluation grid: no rebinning, unit metric, zero data.  ``k_kms=None`` keeps the
    likelihood's own k bins and rebinning tables for those redshifts.  ``remove`` overrides the switch
    tables of the metal models for this context (si_mult.py:185-206, si_add.py:156-159).  ``keep_data``
    (own grid only) keeps the data vector and the inverse covariances of the selected redshifts, the dense
    one when every redshift is selected."""
    iz = np.asarray(iz_sel, dtype=int)
    s = dict(spec)
    s["zs"] = np.asarray(spec["zs"], dtype=float)[iz]
    cosmo = dict(spec["cosmo"])
    cosmo["linP_fid"] = np.asarray(spec["cosmo"]["linP_fid"], dtype=float)[iz]
    cosmo["M_of_z"] = np.asarray(spec["cosmo"]["M_of_z"], dtype=float)[iz]
    s["cosmo"] = cosmo
    s["igm_fid"] = {k: np.asarray(v, dtype=float)[iz] for k, v in spec["igm_fid"].items()}
    cont = dict(spec["cont"])
    cont["apply_resolution"] = bool(apply_resolution)
    if remove:
        for model in ("si_mult", "si_add"):
            m = dict(cont[model])
            m["off"] = {k: (remove[k] if k in remove else v) for k, v in cont[model]["off"].items()}
            cont[model] = m
    s["cont"] = cont
    if k_kms is None:
        ks = [np.asarray(spec["data"]["k_kms"][i], dtype=float) for i in iz]
        rb = spec["rebin"]
        s["rebin"] = None if rb is None else {"rebin_k": rb["rebin_k"], "k_kms": [rb["k_kms"][i] for i in iz],
                                              "cover": [rb["cover"][i] for i in iz], "sum_cover": [rb["sum_cover"][i] for i in iz]}
    else:
        ks = [np.ascontiguousarray(k, dtype=float) for k in k_kms]
        s["rebin"] = None
    s["data"] = {
        "k_kms": ks,
        "Pk_kms": [np.zeros(len(k)) for k in ks],
        "icov": [np.eye(len(k)) for k in ks],
        "full_Pk_kms": None,
        "full_icov": None,
    }
    if keep_data:
        if k_kms is not None:
            raise ValueError("keep_data needs the likelihood's own k bins")
        d = spec["data"]
        s["data"]["Pk_kms"] = [np.asarray(d["Pk_kms"][i], dtype=float) for i in iz]
        s["data"]["icov"] = [np.asarray(d["icov"][i], dtype=float) for i in iz]
        if d["full_icov"] is not None and list(iz) == list(range(len(spec["zs"]))):
            s["data"]["full_Pk_kms"] = d["full_Pk_kms"]
            s["data"]["full_icov"] = d["full_icov"]
        return s
    params = dict(spec["params"])
    params["gauss_sigma_cube"] = None
    s["params"] = params
    return s


class B200Theory(object):
    MAX_CACHED = 4

    def __init__(self, like):
        self._like = like
        self._cache = OrderedDict()
        spec = like.spec
        p = spec["params"]
        self._names = list(p["names"])
        self._pmin = np.asarray(p["min"], dtype=float)
        self._pmax = np.asarray(p["max"], dtype=float)
        # value each free parameter takes when it is absent from like_params: the coefficient its model
        # holds, or the fiducial cosmology
        held = np.full(len(self._names), np.nan)
        for fam in spec["families"].values():
            idx = np.asarray(fam["param_idx"], dtype=int)
            coeffs = np.asarray(fam["coeffs"], dtype=float)
            for j in np.nonzero(idx >= 0)[0]:
                held[idx[j]] = coeffs[j]
        for name, key in (("As", "As_fid"), ("ns", "ns_fid"), ("nrun", "nrun_fid")):
            if name in self._names:
                held[self._names.index(name)] = float(spec["cosmo"][key])
        self._held = held
        self.zs = np.asarray(spec["zs"], dtype=float)
        self.emu_params = list(spec["emulator"]["emu_params"])

    def get_blobs_dtype(self):
        """lya_theory.py:500-512"""
        return [(n, float) for n in BLOB_NAMES]

    # ------------------------------------------------------------------
    def cube_from_like_params(self, like_params):
        """-> (cube coordinates [ndim], apply_resolution); coordinates may lie outside [0, 1]"""
        vals, apply_res = self._values_from_like_params(like_params)
        return (vals - self._pmin) / (self._pmax - self._pmin), apply_res

    def _values_from_like_params(self, like_params):
        """-> (parameter values [ndim], apply_resolution)"""
        vals = self._held.copy()
        if isinstance(like_params, np.ndarray) or (
            isinstance(like_params, (list, tuple)) and len(like_params) and np.isscalar(like_params[0])
        ):
            arr = np.asarray(like_params, dtype=float)
            if arr.shape != vals.shape:
                raise ValueError("a flat like_params array must hold the %d free parameters" % len(vals))
            given = dict(zip(self._names, arr))
        elif isinstance(like_params, dict):
            given = {str(k): float(v) for k, v in like_params.items()}
        else:
            given = {str(par.name): float(par.value) for par in like_params}
        for name, v in given.items():
            if name in self._names:
                vals[self._names.index(name)] = v
            else:
                raise ValueError("like_params holds %r, which is not a free parameter of the compiled likelihood" % name)
        # a family comes whole or not at all (base_igm.py:304-308, base_contaminants.py:215-245)
        for fname, fam in self._like.spec["families"].items():
            idx = np.asarray(fam["param_idx"], dtype=int)
            free = [self._names[i] for i in idx[idx >= 0]]
            n_given = sum(1 for n in free if n in given)
            if 0 < n_given < len(free):
                raise ValueError("number of params mismatch for: " + fname)
        if np.isnan(vals).any():
            missing = [n for n, v in zip(self._names, vals) if np.isnan(v)]
            raise ValueError("no value for the free parameters %s" % missing)
        apply_res = any(n.startswith("R_coeff") for n in given)
        return vals, apply_res

    def _z_index(self, zs):
        """the reference matches redshifts exactly (np.argwhere(zs == z), lya_theory.py:445-498)"""
        iz = []
        for z in np.atleast_1d(np.asarray(zs, dtype=float)):
            hit = np.nonzero(np.abs(self.zs - z) < 1e-8)[0]
            if len(hit) != 1:
                raise ValueError("z = %r is not one of the redshifts of the compiled likelihood" % z)
            iz.append(int(hit[0]))
        return iz

    def _context(self, iz, k_kms, apply_res, remove=None, keep_data=False):
        from .likelihood import B200Likelihood

        rkey = None if not remove else tuple(sorted((str(k), float(v)) for k, v in remove.items()))
        key = (tuple(iz), bool(apply_res), None if k_kms is None else tuple(np.ascontiguousarray(k, dtype=float).tobytes() for k in k_kms), rkey, bool(keep_data))
        sub = self._cache.get(key)
        if sub is None:
            sub = B200Likelihood(grid_spec(self._like.spec, iz, k_kms, apply_res, remove, keep_data), device=self._like.device.index, _with_theory=False)
            self._cache[key] = sub
            while len(self._cache) > self.MAX_CACHED:
                _, old = self._cache.popitem(last=False)
                old.close()
        else:
            self._cache.move_to_end(key)
        # derived contexts follow the emulator arithmetic of the likelihood they belong to
        mode = getattr(self._like, "emulator_precision", "fp64")
        if sub.emulator_precision != mode:
            sub.set_emulator_precision(mode)
        return sub

    def close(self):
        for sub in self._cache.values():
            sub.close()
        self._cache.clear()

    def invalidate(self):
        """drop the derived contexts: the ones built with keep_data carry the data vector and the inverse
        covariances of the likelihood at the time they were made (set_data / set_icov call this)"""
        self.close()

    # ------------------------------------------------------------------
    def get_p1d_kms_batch(self, zs, k_kms, values, apply_hull=True, apply_resolution=True, return_emu_params=False):
        """P[W, sum_z len(k_z)] for cube points ``values[W, ndim]`` on the grid (zs, k_kms); no cube test.
        NaN rows where the reference returns None."""
        iz = self._z_index(zs)
        k_list = self._k_list(k_kms, len(iz))
        sub = self._context(iz, k_list, apply_resolution)
        return sub.get_p1d_kms_batch(values, apply_hull=apply_hull, return_emu_params=return_emu_params, cube_test=False)

    @staticmethod
    def _k_list(k_kms, nz):
        """lya_theory.py:672-685: a list of per-z arrays, or one array for a single redshift"""
        if nz == 1:
            k = np.asarray(k_kms, dtype=float)
            if k.ndim == 2 and k.shape[0] == 1:
                k = k[0]
            if k.ndim == 1:
                return [k]
        if len(k_kms) != nz:
            raise ValueError("k_kms must hold one array per redshift")
        return [np.asarray(k, dtype=float) for k in k_kms]

    def get_p1d_kms(self, zs, k_kms, like_params=[], return_covar=False, return_blob=True, return_emu_params=False,
                    apply_hull=True, hires=False, remove=None, return_contaminants=False):
        """lya_theory.py:610-814"""
        if return_covar or hires or return_contaminants:
            raise NotImplementedError("return_covar, hires and return_contaminants are not on the device path")
        vals, apply_res = self._values_from_like_params(like_params)
        x = (vals - self._pmin) / (self._pmax - self._pmin)
        iz = self._z_index(zs)
        k_list = self._k_list(k_kms, len(iz))
        sub = self._context(iz, k_list, apply_res, remove)
        p, emu = sub.get_p1d_kms_batch(x[None, :], apply_hull=apply_hull, return_emu_params=True, cube_test=False)
        if np.isnan(p[0]).all():
            return None
        out = [[p[0, sub.zoff_d[i] : sub.zoff_d[i + 1]].copy() for i in range(len(iz))]]
        if return_blob:
            out.append(tuple(sub.blobs_batch(x[None, :])[0]))
        if return_emu_params:
            out.append({n: emu[0, :, j].copy() for j, n in enumerate(self.emu_params)})
        return out[0] if len(out) == 1 else out
