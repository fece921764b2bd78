"""ctypes binding of libcup1d_b200.so (include/cup1d_b200.h).

There is no CPU fallback: if the shared library is missing or a CUDA device
is not usable, loading / context creation raises.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import compile as _cmp

ABI_VERSION = 8
LIB_NAME = "libcup1d_b200.so"
LIB_PATH = os.environ.get("C1D_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)  # C1D_LIB: a variant build (kernel tuning experiments)

FLAG_ZMASK = 1
FLAG_NO_HULL = 2
FLAG_EMU_TC = 4
FLAG_NO_CUBE = 8
FLAG_EMU_I8 = 16
FLAG_CHI2_I8 = 32

# every symbol include/cup1d_b200.h declares
SYMBOLS = (
    "c1d_abi_version", "c1d_ctx_create", "c1d_ctx_upload", "c1d_ctx_finalize",
    "c1d_loglike_batch", "c1d_p1d_batch", "c1d_emulator_only", "c1d_loglike_batch_host",
    "c1d_sampler_run", "c1d_sampler_run_ens", "c1d_sampler_half_step", "c1d_philox_u01", "c1d_emulator_paths", "c1d_launch_count", "c1d_graph_replays", "c1d_set_timing", "c1d_get_timing", "c1d_tc_selftest", "c1d_ctx_destroy", "c1d_last_error",
)


class Config(C.Structure):
    """c1d_config (include/cup1d_b200.h) — field order and types must match"""

    _fields_ = [
        ("abi_version", C.c_int32),
        ("nz", C.c_int32), ("ndim", C.c_int32), ("n_emu", C.c_int32), ("nd", C.c_int32), ("nf", C.c_int32),
        ("ell_width", C.c_int32), ("rebin_width", C.c_int32), ("emu_kind", C.c_int32), ("nc", C.c_int32),
        ("n_layers", C.c_int32), ("layer_dims", C.c_int32 * (_cmp.MAX_LAYERS + 1)),
        ("gp_ntrain", C.c_int32), ("gp_nnorm", C.c_int32), ("gp_imf", C.c_int32),
        ("metal_model", C.c_int32), ("apply_resolution", C.c_int32), ("has_full", C.c_int32),
        ("has_gauss", C.c_int32), ("n_hull_facets", C.c_int32),
        ("idx_As", C.c_int32), ("idx_ns", C.c_int32), ("idx_nrun", C.c_int32),
        ("emu_code", C.c_int32 * _cmp.MAX_EMU),
        ("device", C.c_int32),
        ("uniform_k", C.c_int32),
        ("has_agn", C.c_int32),
        ("gp_nm", C.c_int32), ("gp_nkn", C.c_int32),
        ("hcd_gate", C.c_int32),
        ("reserved_i", C.c_int32 * 2),
        ("As_fid", C.c_double), ("ns_fid", C.c_double), ("nrun_fid", C.c_double),
        ("L_emu", C.c_double), ("L_star", C.c_double),
        ("blob_fid", C.c_double * 6), ("star_lo", C.c_double * 3), ("star_hi", C.c_double * 3),
        ("min_log_like", C.c_double), ("hull_tol", C.c_double), ("nn_slope", C.c_double), ("gp_sigma2", C.c_double),
        ("rb3", C.c_double), ("ra3_over_rb3", C.c_double),
        ("c0_o1", C.c_double), ("c0_o2", C.c_double), ("c0_o3c", C.c_double),
        ("reserved_d", C.c_double * 8),
    ]


_lib = None


def load_library(path=None):
    """Load (once) and type the shared library; raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). cup1d_b200 has no CPU fallback." % path
        )
    lib = C.CDLL(path)
    vp, i64, u32, dp = C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p
    lib.c1d_abi_version.restype = C.c_int
    lib.c1d_ctx_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.c1d_ctx_upload.argtypes = [vp, C.c_int, vp, C.c_size_t]
    lib.c1d_ctx_finalize.argtypes = [vp]
    lib.c1d_loglike_batch.argtypes = [vp, dp, i64, dp, dp, dp, dp, u32, vp]
    lib.c1d_p1d_batch.argtypes = [vp, dp, i64, dp, dp, u32, vp]
    lib.c1d_emulator_only.argtypes = [vp, dp, i64, dp, u32, vp]
    lib.c1d_loglike_batch_host.argtypes = [vp, dp, i64, dp, u32]
    lib.c1d_sampler_run.argtypes = [vp, dp, dp, dp, i64, i64, i64, C.c_uint64, C.c_double, C.c_int, C.c_int, dp, dp, dp, dp, u32, vp]
    lib.c1d_sampler_run.restype = C.c_int
    lib.c1d_sampler_run_ens.argtypes = [vp, dp, dp, dp, i64, i64, i64, i64, i64, C.c_uint64, C.c_double, C.c_int, C.c_int, dp, dp, dp, dp, u32, vp]
    lib.c1d_sampler_run_ens.restype = C.c_int
    lib.c1d_sampler_half_step.argtypes = [vp, dp, dp, dp, i64, i64, i64, i64, C.c_int, C.c_uint64, C.c_double, i64, i64, dp, u32, vp]
    lib.c1d_sampler_half_step.restype = C.c_int
    lib.c1d_philox_u01.argtypes = [C.c_uint64, i64, C.c_int, i64, C.c_int, C.POINTER(C.c_double)]
    lib.c1d_philox_u01.restype = C.c_int
    lib.c1d_emulator_paths.argtypes = [vp]
    lib.c1d_emulator_paths.restype = C.c_int
    lib.c1d_launch_count.argtypes = [vp]
    lib.c1d_launch_count.restype = i64
    lib.c1d_graph_replays.argtypes = [vp]
    lib.c1d_graph_replays.restype = i64
    lib.c1d_set_timing.argtypes = [vp, C.c_int]
    lib.c1d_get_timing.argtypes = [vp, C.POINTER(C.c_double)]
    lib.c1d_tc_selftest.argtypes = [C.c_int, vp, vp, vp, C.c_int, C.c_int]
    lib.c1d_tc_selftest.restype = C.c_int
    lib.c1d_ctx_destroy.argtypes = [vp]
    lib.c1d_last_error.argtypes = [vp]
    lib.c1d_last_error.restype = C.c_char_p
    for fn in ("c1d_ctx_create", "c1d_ctx_upload", "c1d_ctx_finalize", "c1d_loglike_batch", "c1d_p1d_batch",
               "c1d_emulator_only", "c1d_loglike_batch_host", "c1d_set_timing", "c1d_get_timing", "c1d_ctx_destroy"):
        getattr(lib, fn).restype = C.c_int
    if lib.c1d_abi_version() != ABI_VERSION:
        raise RuntimeError("libcup1d_b200.so ABI %d != binding ABI %d: rebuild" % (lib.c1d_abi_version(), ABI_VERSION))
    if path == LIB_PATH:
        _lib = lib
    return lib


class C1DError(RuntimeError):
    pass


class Context(object):
    """Owner of one device context (one per GPU / per process rank)."""

    def __init__(self, cfg, fields, device=0):
        self.lib = load_library()
        self.handle = C.c_void_p()
        c = Config()
        c.abi_version = ABI_VERSION
        c.device = int(device)
        for k, v in cfg.items():
            if isinstance(v, (list, tuple)):
                arr = getattr(c, k)
                for i, x in enumerate(v):
                    arr[i] = x
            else:
                setattr(c, k, v)
        self.cfg = dict(cfg)
        self.device = int(device)
        rc = self.lib.c1d_ctx_create(C.byref(c), C.byref(self.handle))
        self._check(rc, "c1d_ctx_create")
        for name, arr in fields.items():
            self.upload(name, arr)
        self._check(self.lib.c1d_ctx_finalize(self.handle), "c1d_ctx_finalize")

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.c1d_last_error(self.handle) if self.handle else b"(no context)"
            raise C1DError("%s failed (%d): %s" % (what, rc, (msg or b"").decode()))

    def upload(self, name, arr):
        dt = np.int32 if name in _cmp.INT_FIELDS else np.float64
        a = np.ascontiguousarray(arr, dtype=dt)
        rc = self.lib.c1d_ctx_upload(self.handle, _cmp.FIELD_ID[name], a.ctypes.data if a.size else None, a.nbytes)
        self._check(rc, "c1d_ctx_upload(%s)" % name)

    def loglike_batch(self, x_ptr, W, lnp_ptr, lnl_z_ptr=None, blobs_ptr=None, chi2_ptr=None, flags=0, stream=None):
        rc = self.lib.c1d_loglike_batch(self.handle, x_ptr, W, lnp_ptr, lnl_z_ptr, blobs_ptr, chi2_ptr, flags, stream)
        self._check(rc, "c1d_loglike_batch")

    def p1d_batch(self, x_ptr, W, p_ptr, emu_ptr=None, flags=0, stream=None):
        self._check(self.lib.c1d_p1d_batch(self.handle, x_ptr, W, p_ptr, emu_ptr, flags, stream), "c1d_p1d_batch")

    def emulator_only(self, in_ptr, nrows, out_ptr, flags=0, stream=None):
        self._check(self.lib.c1d_emulator_only(self.handle, in_ptr, nrows, out_ptr, flags, stream), "c1d_emulator_only")

    def loglike_batch_host(self, x_host, out_host, flags=0):
        """x_host [W, ndim] float64 C-contiguous NumPy (pinned or pageable),
        out_host [W, 7] float64"""
        W = x_host.shape[0]
        rc = self.lib.c1d_loglike_batch_host(self.handle, x_host.ctypes.data, W, out_host.ctypes.data, flags)
        self._check(rc, "c1d_loglike_batch_host")

    def sampler_run(self, x_ptr, lnp_ptr, blobs_ptr, W, nsteps, step0, seed, a=2.0, init=True, thin=1,
                    chain_ptr=None, lnp_chain_ptr=None, blobs_chain_ptr=None, naccept_ptr=None, flags=0, stream=None,
                    n_ens=1, ens0=0):
        rc = self.lib.c1d_sampler_run_ens(self.handle, x_ptr, lnp_ptr, blobs_ptr, W, n_ens, ens0, nsteps, step0, seed, a, int(init), thin,
                                          chain_ptr, lnp_chain_ptr, blobs_chain_ptr, naccept_ptr, flags, stream)
        self._check(rc, "c1d_sampler_run_ens")

    def emulator_paths(self):
        """bit mask: 1 float64, FLAG_EMU_TC 3xTF32, FLAG_EMU_I8 int8 digit slices, FLAG_CHI2_I8 dense chi2 as int8 digit slices"""
        return int(self.lib.c1d_emulator_paths(self.handle))

    def sampler_half_step(self, x_ptr, lnp_ptr, blobs_ptr, W, step, half, seed, s0, ns, a=2.0, naccept_ptr=None, flags=0, stream=None,
                          n_ens=1, ens0=0):
        rc = self.lib.c1d_sampler_half_step(self.handle, x_ptr, lnp_ptr, blobs_ptr, W, n_ens, ens0, step, half, seed, a, s0, ns,
                                            naccept_ptr, flags, stream)
        self._check(rc, "c1d_sampler_half_step")

    def launch_count(self):
        return int(self.lib.c1d_launch_count(self.handle))

    def graph_replays(self):
        return int(self.lib.c1d_graph_replays(self.handle))

    def set_timing(self, on):
        self._check(self.lib.c1d_set_timing(self.handle, int(bool(on))), "c1d_set_timing")

    def get_timing(self):
        buf = (C.c_double * 8)()
        self._check(self.lib.c1d_get_timing(self.handle, buf), "c1d_get_timing")
        return list(buf)

    def close(self):
        if self.handle:
            self.lib.c1d_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
