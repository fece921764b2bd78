"""Golden vectors for the covariance builder: runs the reference's own
``Likelihood.set_icov`` (cup1d/likelihood/likelihood.py:233-507) in this container for the three
emulator-covariance modes and stores its inputs and outputs.

    python tools/make_golden_icov.py        # writes tests/golden/icov_<mode>.npz

The reference is only imported here (through tools/ref_harness.py); the fixtures travel.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from tools import ref_harness as rh  # noqa: E402
from tools.make_golden import small_nn, synth_cov, grids  # noqa: E402
from oracle.emulators import RefMLPEmulator  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def data_namespace(nz, grid, seed):
    zs = np.round(2.2 + 0.2 * np.arange(nz), 10)
    ks, kmins, kmaxs = grids(grid, nz)
    nd = sum(len(k) for k in ks)
    rng = np.random.default_rng(seed)
    p = 60.0 * np.exp(-np.concatenate(ks) / 0.02) * (1 + 0.1 * rng.normal(0, 1, nd)) + 1.0
    cov_full = synth_cov(p, seed + 1)
    stat_frac = 0.5 + 0.4 * rng.random(nd)  # a statistical part that is not a multiple of the total
    cov_stat_full = cov_full * np.sqrt(stat_frac[:, None] * stat_frac[None, :])
    Pk, cov, cstat, off = [], [], [], 0
    for k in ks:
        sl = slice(off, off + len(k))
        Pk.append(p[sl])
        cov.append(cov_full[sl, sl])
        cstat.append(cov_stat_full[sl, sl])
        off += len(k)
    return rh.make_data_namespace(zs, ks, kmins, kmaxs, Pk, cov, cstat, full_cov=cov_full, full_cov_stat=cov_stat_full)


def hook_factors(args):
    # non-trivial inflation factors (input_pipeline.py:634-655 sets them all to one)
    n = len(args.cov_factor["z"])
    rng = np.random.default_rng(7)
    for key in ("val_stat", "val_syst", "val_emu", "val_full"):
        args.cov_factor[key] = 0.8 + 0.5 * rng.random(n)


def main():
    emu = RefMLPEmulator(small_nn(5, "mpg"), "synthetic_mpg_nn", "mpg")
    for mode in ("diagonal", "block", "full"):
        data = data_namespace(5, "dr1c", 40)
        sink = io.StringIO()
        t0 = time.time()
        with contextlib.redirect_stdout(sink):
            like, args = rh.build_reference_likelihood(emu, data, rebin_k=1, emu_cov_type=mode, args_hook=hook_factors)
        dt = time.time() - t0
        from cup1d.utils.utils import get_path_repo

        path = os.path.join(get_path_repo("lace"), "data", "covariance", "l1O_cov_" + emu.emulator_label + ".npy")
        emu_cov = np.load(path, allow_pickle=True).item()
        zs = np.asarray(data.z)
        conv = np.array([like.theory.fid_cosmo["cosmo"].dkms_dMpc(z) for z in zs])
        nd = len(data.full_Pk_kms)
        out = dict(
            mode=mode, z=zs, nk=np.array([len(k) for k in data.k_kms]), k_kms=np.concatenate(data.k_kms),
            Pk_kms=np.concatenate(data.Pk_kms), full_cov=data.full_cov_Pk_kms, full_cov_stat=data.full_cov_stat_Pk_kms,
            dkms_dMpc=conv,
            factor_z=np.asarray(like.cov_factor["z"]), val_stat=np.asarray(like.cov_factor["val_stat"]),
            val_syst=np.asarray(like.cov_factor["val_syst"]), val_emu=np.asarray(like.cov_factor["val_emu"]),
            val_full=np.asarray(like.cov_factor["val_full"]),
            emu_zz_zk=emu_cov["zz_zk"], emu_k_Mpc_zk=emu_cov["k_Mpc_zk"], emu_cov_zk=emu_cov["cov_zk"],
            ref_full_cov=like.full_cov_Pk_kms, ref_full_icov=like.full_icov_Pk_kms, ref_emu_full_cov=like.emu_full_cov_Pk_kms,
        )
        for iz in range(len(zs)):
            out["ref_cov_%d" % iz] = like.cov_Pk_kms[iz]
            out["ref_icov_%d" % iz] = like.icov_Pk_kms[iz]
            out["ref_cov_emu_%d" % iz] = like.cov_emu_Pk_kms[iz]
        os.makedirs(OUT, exist_ok=True)
        fn = os.path.join(OUT, "icov_%s.npz" % mode)
        np.savez_compressed(fn, **out)
        print("%-9s N_d=%d  reference Likelihood() %.2f s  ->  %s (%.0f kB)" % (mode, nd, dt, fn, os.path.getsize(fn) / 1e3))


if __name__ == "__main__":
    main()
