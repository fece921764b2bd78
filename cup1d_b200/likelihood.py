"""Drop-in for the hot path of ``cup1d.likelihood.likelihood.Likelihood``.

``B200Likelihood(like)`` compiles a reference-API likelihood (or a model
spec) into a device context and exposes

* the reference's scalar methods with the reference's names, argument
  meaning and return conventions — ``get_p1d_kms``, ``get_chi2``,
  ``get_log_like``, ``compute_log_prob``, ``log_prob``,
  ``log_prob_and_blobs``, ``minus_log_prob`` (likelihood.py:722-1072) — so
  ``Fitter`` / ``emcee`` / ``scipy.optimize`` hooks keep working, and
* batched entry points over ``values[W, ndim]``:
  ``log_prob_batch``, ``log_prob_and_blobs_batch``, ``get_log_like_batch``,
  ``get_chi2_batch``, ``get_p1d_kms_batch`` and ``log_prob_and_blobs_vectorized``
  (the ``[W, 7]`` array ``emcee.EnsembleSampler(..., vectorize=True)`` wants).

CUDA tensors in -> CUDA tensors out, asynchronously on the current torch
stream.  NumPy in -> NumPy out through the host-buffer C-ABI call.  There is
no CPU implementation behind this class.
"""

from __future__ import annotations

import os

import numpy as np
import torch

from . import _abi
from .compile import compile_spec
from .spec import spec_from_likelihood, spec_sizes
from .theory import B200Theory

BLOB_NAMES = ("Delta2_star", "n_star", "alpha_star", "f_star", "g_star", "H0")


class B200Likelihood(object):
    def __init__(self, like, device=None, emulator_precision="auto", _with_theory=True):
        """like: a reference ``Likelihood`` (any object with its attributes) or
        a model spec dict.  emulator_precision (NN emulators): "i8" = tcgen05 tensor
        cores, int8 digit slices with exact int32 accumulation and float64 recombination
        (log-likelihood within 1e-4 of a float64 evaluation of the weights); "fp64" = the
        fp64 tensor cores (DMMA), bit-level parity; "3xtf32" = tcgen05 in 3xTF32
        (float32-class accuracy, misses the log-likelihood tolerance); "auto" (default)
        = "i8" where the network shape allows it, else "fp64"."""
        if not torch.cuda.is_available():
            raise RuntimeError("cup1d_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.spec = like if isinstance(like, dict) else spec_from_likelihood(like)
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.nz, self.ndim, self.nd, self.nf = spec_sizes(self.spec)
        cfg, fields = compile_spec(self.spec)
        self.ctx = _abi.Context(cfg, fields, device=self.device.index)
        self.n_emu = cfg["n_emu"]
        self.has_full = bool(cfg["has_full"])
        p = self.spec["params"]
        self.free_param_names = list(p["names"])
        self.pmin = np.asarray(p["min"], dtype=float)
        self.pmax = np.asarray(p["max"], dtype=float)
        self.min_log_like = float(p["min_log_like"])
        self.Gauss_priors = p["gauss_sigma_cube"]
        self.fid_cube = (np.asarray(p["value"], dtype=float) - self.pmin) / (self.pmax - self.pmin)
        self.zs = np.asarray(self.spec["zs"], dtype=float)
        self.zoff_d = fields["ZOFF_D"].astype(int)
        self._zmask_key = None
        # Theory.get_p1d_kms on other (z, k) grids (derived contexts, built on demand)
        self.theory = B200Theory(self) if _with_theory else None
        self.set_emulator_precision(emulator_precision)
        self._set_logdet()

    def _set_logdet(self):
        """constants for ignore_log_det_cov=False (likelihood.py:950, 962)"""
        d = self.spec["data"]
        self._logdet_z = np.array([np.log(np.abs(1 / np.linalg.det(c))) for c in d["icov"]])
        self._logdet_full = None
        if d["full_icov"] is not None:
            sign, ld = np.linalg.slogdet(d["full_icov"])
            self._logdet_full = -ld

    # ------------------------------------------------------------------
    def set_emulator_precision(self, mode):
        if mode not in ("auto", "fp64", "i8", "3xtf32"):
            raise ValueError("emulator_precision must be 'auto', 'i8', 'fp64' or '3xtf32'")
        paths = self.ctx.emulator_paths()
        if mode == "auto":
            mode = "i8" if (paths & _abi.FLAG_EMU_I8) else "fp64"
        if mode in ("3xtf32", "i8") and self.spec["emulator"]["kind"] != "nn":
            raise ValueError("the tcgen05 paths are for NN emulators")
        if mode == "i8" and not (paths & _abi.FLAG_EMU_I8):
            raise ValueError("the int8 digit-slice kernel does not support this network shape (a layer wider than 192 "
                             "must be followed by one of at most 64); use 'fp64'")
        self.emulator_precision = mode
        self._base_flags = {"fp64": 0, "3xtf32": _abi.FLAG_EMU_TC, "i8": _abi.FLAG_EMU_I8}[mode]
        # the dense chi2 follows the emulator: with the digit-slice emulator (log-likelihood within 1e-4) the quadratic
        # form runs on tcgen05 too (five digit planes: 5e-11 relative), with "fp64" everything stays on the fp64 pipe
        self.set_chi2_precision("i8" if mode == "i8" and (paths & _abi.FLAG_CHI2_I8) else "fp64")

    def set_chi2_precision(self, mode):
        """arithmetic of the dense chi2: "i8" (tcgen05 digit slices, up to 704 data points) or "fp64" (DMMA);
        C1D_CHI2=fp64|i8 in the environment overrides the choice that follows the emulator precision"""
        mode = os.environ.get("C1D_CHI2", mode)
        if mode not in ("i8", "fp64"):
            raise ValueError("chi2 precision must be 'i8' or 'fp64'")
        if mode == "i8" and not (self.ctx.emulator_paths() & _abi.FLAG_CHI2_I8):
            if "C1D_CHI2" in os.environ:
                mode = "fp64"  # no dense covariance (or more than 704 points): nothing to switch
            else:
                raise ValueError("the int8 digit-slice chi2 needs a dense inverse covariance of at most 704 data points")
        self.chi2_precision = mode
        self._base_flags = (self._base_flags & ~_abi.FLAG_CHI2_I8) | (_abi.FLAG_CHI2_I8 if mode == "i8" else 0)

    def set_data(self, full_Pk_kms):
        """Replace the data vector (mock generation)"""
        v = np.ascontiguousarray(full_Pk_kms, dtype=float).reshape(-1)
        if v.size != self.nd:
            raise ValueError("data vector length")
        self.ctx.upload("DATA", v)
        o = 0
        for iz, k in enumerate(self.spec["data"]["k_kms"]):
            self.spec["data"]["Pk_kms"][iz] = v[o : o + len(k)].copy()
            o += len(k)
        if self.spec["data"]["full_icov"] is not None:
            self.spec["data"]["full_Pk_kms"] = v.copy()
        if self.theory is not None:
            self.theory.invalidate()  # derived contexts that carry the data vector and the metric

    def set_icov(self, icov_Pk_kms, full_icov_Pk_kms=None):
        """Replace the inverse covariances stage 4 keeps on the device: the
        per-z blocks ``icov_Pk_kms`` and, for a full-covariance context, ``full_icov_Pk_kms`` —
        the attributes ``Likelihood.set_icov`` fills (likelihood.py:385-399, 496-503), e.g. from
        ``cup1d_b200.covariance.build_covariance``.  Shapes must match the context."""
        d = self.spec["data"]
        blocks = [np.ascontiguousarray(c, dtype=float) for c in icov_Pk_kms]
        if len(blocks) != self.nz or any(b.shape != (len(k), len(k)) for b, k in zip(blocks, d["k_kms"])):
            raise ValueError("per-z inverse covariance blocks do not match the data vector")
        if self.has_full != (full_icov_Pk_kms is not None):
            raise ValueError("context was built %s a full inverse covariance" % ("with" if self.has_full else "without"))
        self.ctx.upload("ICOV_BLOCKS", np.concatenate([b.reshape(-1) for b in blocks]))
        d["icov"] = blocks
        if self.has_full:
            full = np.ascontiguousarray(full_icov_Pk_kms, dtype=float)
            if full.shape != (self.nd, self.nd):
                raise ValueError("full inverse covariance shape")
            # d^T A d only sees the symmetric part of A; the chi2 kernel visits one triangle and doubles it
            self.ctx.upload("ICOV_FULL", 0.5 * (full + full.T))
            d["full_icov"] = full
        self._set_logdet()
        if self.theory is not None:
            self.theory.invalidate()

    def get_blobs_dtype(self):
        """lya_theory.py:500-512"""
        return [(n, float) for n in BLOB_NAMES]

    def parameters_from_sampling_point(self, values):
        """likelihood.py:561-574 as a flat array of parameter values"""
        values = np.asarray(values, dtype=float)
        return self.pmin + values * (self.pmax - self.pmin)

    def sampling_point_from_parameters(self):
        """likelihood.py:552-559"""
        return self.fid_cube.copy()

    # ------------------------------------------------------------------
    def _flags(self, zmask=None, apply_hull=True, cube_test=True):
        flags = self._base_flags
        if not apply_hull:
            flags |= _abi.FLAG_NO_HULL
        if not cube_test:
            flags |= _abi.FLAG_NO_CUBE
        if zmask is not None:
            zmask = np.atleast_1d(np.asarray(zmask, dtype=float))
            sel = np.array([int(np.any(np.abs(zmask - z) < 1e-3)) for z in self.zs], dtype=np.int32)
            key = sel.tobytes()
            if key != self._zmask_key:
                self.ctx.upload("ZMASK", sel)
                self._zmask_key = key
            flags |= _abi.FLAG_ZMASK
        return flags

    def _as_device(self, values):
        """-> (x [W, ndim] float64 contiguous CUDA tensor, was_numpy, was_1d)"""
        was_numpy = not torch.is_tensor(values)
        x = torch.as_tensor(values, dtype=torch.float64) if was_numpy else values
        was_1d = x.dim() == 1
        if was_1d:
            x = x.unsqueeze(0)
        if x.dim() != 2 or x.shape[1] != self.ndim:
            raise ValueError("values must have shape [W, %d]" % self.ndim)
        x = x.to(device=self.device, dtype=torch.float64).contiguous()
        return x, was_numpy, was_1d

    def _evaluate(self, values, want_z=False, want_chi2=False, zmask=None, apply_hull=True, cube_test=True):
        x, was_numpy, _ = self._as_device(values)
        W = x.shape[0]
        with torch.cuda.device(self.device):
            lnp = torch.empty(W, dtype=torch.float64, device=self.device)
            blobs = torch.empty((W, 6), dtype=torch.float64, device=self.device)
            lnl_z = torch.empty((W, self.nz), dtype=torch.float64, device=self.device) if want_z else None
            chi2 = torch.empty(W, dtype=torch.float64, device=self.device) if want_chi2 else None
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.ctx.loglike_batch(
                x.data_ptr(), W, lnp.data_ptr(),
                lnl_z.data_ptr() if want_z else None, blobs.data_ptr(),
                chi2.data_ptr() if want_chi2 else None,
                self._flags(zmask, apply_hull, cube_test), stream,
            )
        return lnp, blobs, lnl_z, chi2, was_numpy

    @staticmethod
    def _out(t, was_numpy):
        return t.cpu().numpy() if was_numpy else t

    # ---- batched entry points -----------------------------------------
    def _logdet_total(self, zmask=None):
        """log det C entering the log-likelihood with ignore_log_det_cov=False (likelihood.py:950, 962): the dense
        matrix when it is the metric of the call, else the sum over the (selected) redshift bins"""
        if self.has_full and zmask is None:
            return float(self._logdet_full)
        sel = np.ones(self.nz, dtype=bool)
        if zmask is not None:
            sel = np.array([bool(np.any(np.abs(np.atleast_1d(zmask) - z) < 1e-3)) for z in self.zs])
        return float(np.sum(self._logdet_z[sel]))

    def _add_logdet(self, lnp, zmask):
        """log_like - 0.5 log det C for the rows that have a likelihood; rejected rows stay at min_log_like (+ prior),
        as regulate_log_like leaves them (likelihood.py:974-980, 1019)"""
        ld = 0.5 * self._logdet_total(zmask)
        if torch.is_tensor(lnp):
            return torch.where(lnp > 0.01 * self.min_log_like, lnp - ld, lnp)
        return np.where(lnp > 0.01 * self.min_log_like, lnp - ld, lnp)

    def log_prob_batch(self, values, zmask=None, ignore_log_det_cov=True):
        """lnP[W] = Likelihood.log_prob for every row (likelihood.py:1026)"""
        if zmask is None and not torch.is_tensor(values):
            lnp = self.log_prob_and_blobs_vectorized(values)[:, 0].copy()
        else:
            lnp, _, _, _, was_numpy = self._evaluate(values, zmask=zmask)
            lnp = self._out(lnp, was_numpy)
        return lnp if ignore_log_det_cov else self._add_logdet(lnp, zmask)

    def log_prob_and_blobs_batch(self, values, zmask=None, ignore_log_det_cov=True):
        """(lnP[W], blobs[W, 6]) = Likelihood.log_prob_and_blobs per row (:1036)"""
        if zmask is None and not torch.is_tensor(values):
            out = self.log_prob_and_blobs_vectorized(values)
            lnp, blobs = out[:, 0].copy(), out[:, 1:].copy()
        else:
            lnp, blobs, _, _, was_numpy = self._evaluate(values, zmask=zmask)
            lnp, blobs = self._out(lnp, was_numpy), self._out(blobs, was_numpy)
        return (lnp if ignore_log_det_cov else self._add_logdet(lnp, zmask)), blobs

    def log_prob_and_blobs_vectorized(self, values, out=None):
        """[W, 7] rows (lnP, *blob): what ``emcee.EnsembleSampler(...,
        vectorize=True, blobs_dtype=get_blobs_dtype())`` expects.  NumPy in,
        NumPy out, through the host-buffer C-ABI call (copies included)."""
        if torch.is_tensor(values):
            lnp, blobs, _, _, _ = self._evaluate(values)
            return torch.cat([lnp[:, None], blobs], dim=1)
        x = np.ascontiguousarray(values, dtype=np.float64)
        if x.ndim == 1:
            x = x[None, :]
        if x.shape[1] != self.ndim:
            raise ValueError("values must have shape [W, %d]" % self.ndim)
        if out is None:
            out = np.empty((x.shape[0], 7), dtype=np.float64)
        self.ctx.loglike_batch_host(x, out, self._base_flags)
        return out

    def get_log_like_batch(self, values, ignore_log_det_cov=True, zmask=None):
        """(lnL[W], lnL_z[W, Nz]): -inf rows where the reference returns its
        null output (likelihood.py:822-972)"""
        lnp, blobs, lnl_z, chi2, was_numpy = self._evaluate(values, want_z=True, want_chi2=True, zmask=zmask)
        lnl = -0.5 * chi2
        if not ignore_log_det_cov:
            ldz = torch.as_tensor(self._logdet_z, device=self.device)
            if zmask is not None:
                sel = torch.as_tensor([bool(np.any(np.abs(np.atleast_1d(zmask) - z) < 1e-3)) for z in self.zs], device=self.device)
                ldz = ldz * sel
            ok = torch.isfinite(lnl)
            lnl_z = torch.where(ok[:, None], lnl_z - 0.5 * ldz[None, :], lnl_z)
            if self.has_full and zmask is None:
                lnl = torch.where(ok, lnl - 0.5 * self._logdet_full, lnl)
            else:
                lnl = torch.where(ok, lnl_z.sum(dim=1), lnl)
        return self._out(lnl, was_numpy), self._out(lnl_z, was_numpy)

    def get_chi2_batch(self, values, zmask=None):
        """chi2[W] (likelihood.py:787-798); +inf where the reference gives -2*(-inf)"""
        _, _, _, chi2, was_numpy = self._evaluate(values, want_chi2=True, zmask=zmask)
        return self._out(chi2, was_numpy)

    def blobs_batch(self, values):
        """blobs[W, 6] (Theory.get_blob_fixed_background, lya_theory.py:536-584) without the unit-cube
        test; zeros where the star-prior box or the hulls reject the point"""
        _, blobs, _, _, was_numpy = self._evaluate(values, cube_test=False)
        return self._out(blobs, was_numpy)

    def get_p1d_kms_batch(self, values, apply_hull=True, return_emu_params=False, cube_test=True):
        """P[W, N_d]: the rebinned model concatenated over z in the order of
        ``data.full_Pk_kms`` (likelihood.py:957); NaN rows where the reference
        returns None (and, with ``cube_test``, outside the unit cube).  With
        return_emu_params also emu[W, Nz, n_emu]."""
        x, was_numpy, _ = self._as_device(values)
        W = x.shape[0]
        with torch.cuda.device(self.device):
            p = torch.empty((W, self.nd), dtype=torch.float64, device=self.device)
            emu = torch.empty((W, self.nz, self.n_emu), dtype=torch.float64, device=self.device) if return_emu_params else None
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.ctx.p1d_batch(x.data_ptr(), W, p.data_ptr(), emu.data_ptr() if return_emu_params else None,
                               self._flags(None, apply_hull, cube_test), stream)
        if return_emu_params:
            return self._out(p, was_numpy), self._out(emu, was_numpy)
        return self._out(p, was_numpy)

    # ---- on-device ensemble sampler --------------------
    def run_sampler_device(self, p0, nsteps, nburn=0, thin=1, seed=0, a=2.0, zmask=None, to_numpy=True, n_ensembles=1, first_ensemble=0):
        """Stretch-move MCMC entirely on the GPU (c1d_sampler_run_ens): ``p0[W, ndim]``
        initial walkers, ``nburn`` discarded steps, then ``nsteps`` steps of
        which every ``thin``-th is kept.  ``n_ensembles`` independent ensembles of W / n_ensembles walkers (an even
        number) advance together - the reference's production layout is one ensemble per MPI rank (fitter.py:216-272);
        rows e*wn .. (e+1)*wn - 1 of ``p0`` are ensemble e and the chain keeps that order (concatenation on the walker
        axis, fitter.py:262-268).  ``first_ensemble`` is the global number of ensemble 0 (random-number key: an
        ensemble gives the same chain wherever it runs).  Returns the same products as
        Fitter.run_sampler keeps (fitter.py:238-246): dict(chain[S, W, ndim],
        lnprob[S, W], blobs[S, W] structured, acceptance_fraction[W])."""
        x, _, _ = self._as_device(np.array(p0, dtype=float) if not torch.is_tensor(p0) else p0.clone())
        x = x.clone()
        W = x.shape[0]
        if W % n_ensembles or (W // n_ensembles) % 2:
            raise ValueError("the device sampler needs an even number of walkers per ensemble")
        nkeep = nsteps // thin
        flags = self._flags(zmask, True)
        with torch.cuda.device(self.device):
            dev = self.device
            lnp = torch.empty(W, dtype=torch.float64, device=dev)
            blobs = torch.empty((W, 6), dtype=torch.float64, device=dev)
            chain = torch.empty((nkeep, W, self.ndim), dtype=torch.float64, device=dev)
            lchain = torch.empty((nkeep, W), dtype=torch.float64, device=dev)
            bchain = torch.empty((nkeep, W, 6), dtype=torch.float64, device=dev)
            nacc = torch.zeros(W, dtype=torch.int64, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            if nburn > 0:
                self.ctx.sampler_run(x.data_ptr(), lnp.data_ptr(), blobs.data_ptr(), W, nburn, 0, seed, a, True, 1,
                                     None, None, None, None, flags, stream, n_ens=n_ensembles, ens0=first_ensemble)
            nacc.zero_()
            self.ctx.sampler_run(x.data_ptr(), lnp.data_ptr(), blobs.data_ptr(), W, nsteps, nburn, seed, a, nburn == 0, thin,
                                 chain.data_ptr(), lchain.data_ptr(), bchain.data_ptr(), nacc.data_ptr(), flags, stream,
                                 n_ens=n_ensembles, ens0=first_ensemble)
        out = {"chain": chain, "lnprob": lchain, "blobs": bchain, "acceptance_fraction": nacc.double() / max(1, nsteps),
               "last_state": x, "last_lnprob": lnp}
        if to_numpy:
            out = {k: v.cpu().numpy() for k, v in out.items()}
            b = np.zeros(out["blobs"].shape[:-1], dtype=self.get_blobs_dtype())
            for i, nme in enumerate(BLOB_NAMES):
                b[nme] = out["blobs"][..., i]
            out["blobs"] = b
        return out

    # ---- scalar API with the reference's conventions ------------------
    def _eval_target(self, values):
        """(cube point, likelihood object to evaluate it on, cube test).  ``values=None`` means
        ``like_params=[]`` in the reference (likelihood.py:749-753): every model at the coefficients it holds,
        no resolution term and no unit-cube test.  On this context that is the same evaluation whenever the
        held R_coeff coefficients are zero (the term is then exactly 1); otherwise a derived context without
        the resolution term, with the same data and metric, serves the call."""
        if values is not None:
            return np.asarray(values, dtype=float), self, True
        x, _ = self.theory.cube_from_like_params([])
        names = self.free_param_names
        held_res = [self.theory._held[i] for i, n in enumerate(names) if n.startswith("R_coeff")]
        if self.spec["cont"]["apply_resolution"] and np.any(np.asarray(held_res) != 0.0):
            return x, self.theory._context(list(range(self.nz)), None, False, None, keep_data=True), False
        return x, self, False

    def get_p1d_kms(self, zs=None, _k_kms=None, values=None, return_covar=False, return_blob=False,
                    return_emu_params=False, apply_hull=True, remove=None):
        """likelihood.py:722-785.  As in the reference: ``values=None`` evaluates the models at the
        coefficients they hold, without the resolution term (``like_params=[]``, lya_theory.py:712-722);
        with rebinning the k grid is always the rebinning grid of the nearest data redshift and
        ``_k_kms`` is ignored (:741-747); there is no unit-cube test here.  ``None`` when the priors reject
        the point, otherwise ``[p1ds(, blob)(, emu_call)]``."""
        if return_covar:
            raise NotImplementedError("return_covar is not on the device path")
        rebinned = self.spec["rebin"] is not None
        if zs is None:
            iz = list(range(self.nz))
        else:
            iz = [int(np.argmin(np.abs(z - self.zs))) for z in np.atleast_1d(np.asarray(zs, dtype=float))]
        custom_k = (_k_kms is not None) and not rebinned
        if values is None:
            x, apply_res = self.theory.cube_from_like_params([])
        else:
            x, apply_res = np.asarray(values, dtype=float), bool(self.spec["cont"]["apply_resolution"])
        if values is not None and not custom_k and not remove and iz == list(range(self.nz)):
            sub = self
        else:
            k_list = self.theory._k_list(_k_kms, len(iz)) if custom_k else None
            sub = self.theory._context(iz, k_list, apply_res, remove)
        p, emu = sub.get_p1d_kms_batch(x[None, :], apply_hull=apply_hull, return_emu_params=True, cube_test=False)
        if np.isnan(p[0]).all():
            return None
        out = [[p[0, sub.zoff_d[i] : sub.zoff_d[i + 1]].copy() for i in range(len(iz))]]
        if return_blob:
            out.append(tuple(sub.blobs_batch(x[None, :])[0]))
        if return_emu_params:
            names = self.spec["emulator"]["emu_params"]
            out.append({n: emu[0, :, j].copy() for j, n in enumerate(names)})
        return out

    def get_log_like(self, values=None, ignore_log_det_cov=True, return_blob=False, zmask=None):
        """likelihood.py:822-972: [log_like, log_like_all[1, Nz](, blob)]"""
        v, tgt, cube_test = self._eval_target(values)
        lnp, blobs, lnl_z, chi2, _ = tgt._evaluate(v[None, :], want_z=True, want_chi2=True, zmask=zmask, cube_test=cube_test)
        # NumPy scalars like the reference's (its Fitter calls chi2.copy(), fitter.py:432)
        ll, llz = np.float64(-0.5 * chi2[0].item()), lnl_z[0].cpu().numpy()
        if not ignore_log_det_cov and np.isfinite(ll):
            # log det C per redshift and of the dense matrix (likelihood.py:950, 962)
            sel = np.ones(self.nz, dtype=bool)
            if zmask is not None:
                sel = np.array([bool(np.any(np.abs(np.atleast_1d(zmask) - z) < 1e-3)) for z in self.zs])
            llz = llz - 0.5 * self._logdet_z * sel
            ll = np.float64(ll - 0.5 * self._logdet_full) if (self.has_full and zmask is None) else np.float64(np.sum(llz))
        if not np.isfinite(ll):
            out = [-np.inf, -np.inf]
            if return_blob:
                out.append((0, 0, 0, 0, 0, 0))
            return out
        out = [ll, llz[None, :].copy()]
        if return_blob:
            out.append(tuple(blobs[0].cpu().numpy()))
        return out

    def get_chi2(self, values=None, return_all=False, zmask=None):
        """likelihood.py:787-798"""
        log_like, log_like_all = self.get_log_like(values, ignore_log_det_cov=True, zmask=zmask)
        if return_all:
            return -2.0 * log_like, -2.0 * log_like_all
        return -2.0 * log_like

    def regulate_log_like(self, log_like):
        """likelihood.py:974-980"""
        if (log_like is None) or np.isnan(log_like):
            return self.min_log_like
        return max(self.min_log_like, log_like)

    def get_log_prior(self, values):
        """likelihood.py:1049-1064"""
        values = np.asarray(values, dtype=float)
        if max(values) > 1 or min(values) < 0:
            return self.min_log_like
        if self.Gauss_priors is None:
            return 0
        return -np.sum((self.fid_cube - values) ** 2 / (2 * np.asarray(self.Gauss_priors) ** 2))

    def compute_log_prob(self, values, return_blob=False, ignore_log_det_cov=True, zmask=None):
        """likelihood.py:982-1024; with ignore_log_det_cov=False the log det C term of get_log_like (:950, :962)
        enters before the regulation (:1019), through slogdet (the reference's det overflows for large matrices)"""
        v = np.asarray(values, dtype=float)
        if zmask is None:
            row = self.log_prob_and_blobs_vectorized(v[None, :])[0]
            lnp, blob = float(row[0]), tuple(row[1:])
        else:
            lnp_t, blobs_t = self.log_prob_and_blobs_batch(v[None, :], zmask=zmask)
            lnp, blob = float(lnp_t[0]), tuple(blobs_t[0])
        if not ignore_log_det_cov:
            lnp = float(self._add_logdet(np.array([lnp]), zmask)[0])
        if return_blob:
            return lnp, blob
        return lnp

    def log_prob(self, values, ignore_log_det_cov=True, zmask=None):
        """likelihood.py:1026-1034"""
        return self.compute_log_prob(values, False, ignore_log_det_cov, zmask)

    def log_prob_and_blobs(self, values, ignore_log_det_cov=True, zmask=None):
        """likelihood.py:1036-1047"""
        lnprob, blob = self.compute_log_prob(values, True, ignore_log_det_cov, zmask)
        return (lnprob, *blob)

    def minus_log_prob(self, values, zmask=None, ind_fix=None, pfix=None):
        """likelihood.py:1066-1072 (in-place fixing of values included)"""
        if ind_fix is not None:
            values[ind_fix] = pfix
        return -1.0 * self.log_prob(values, zmask=zmask)

    # ------------------------------------------------------------------
    def close(self):
        if self.theory is not None:
            self.theory.close()
        self.ctx.close()
