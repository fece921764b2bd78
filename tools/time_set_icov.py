"""Times Likelihood.set_icov of the reference (likelihood.py:233-507) against
cup1d_b200.covariance.build_covariance on a DR1-sized data vector (11 z, 681 k) in this container.

    python tools/time_set_icov.py [mode ...]
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
from tools.make_golden_icov import hook_factors  # noqa: E402
from oracle.emulators import RefMLPEmulator  # noqa: E402
from cup1d_b200.covariance import build_covariance  # noqa: E402


def main():
    modes = sys.argv[1:] or ["block", "full"]
    emu = RefMLPEmulator(small_nn(5, "mpg"), "synthetic_mpg_nn", "mpg")
    zs = np.round(2.2 + 0.2 * np.arange(11), 10)
    ks, kmins, kmaxs = grids("dr1", 11)
    nd = sum(len(k) for k in ks)
    rng = np.random.default_rng(3)
    p = 60.0 * np.exp(-np.concatenate(ks) / 0.02) + 1.0
    cov_full = synth_cov(p, 4)
    Pk, cov, off = [], [], 0
    for k in ks:
        sl = slice(off, off + len(k))
        Pk.append(p[sl]); cov.append(cov_full[sl, sl]); off += len(k)
    data = rh.make_data_namespace(zs, ks, kmins, kmaxs, Pk, cov, [0.8 * c for c in cov], full_cov=cov_full, full_cov_stat=0.8 * cov_full)
    for mode in modes:
        with contextlib.redirect_stdout(io.StringIO()):
            like, _ = rh.build_reference_likelihood(emu, data, rebin_k=8, emu_cov_type=mode, args_hook=hook_factors)
            t0 = time.time(); like.set_icov(); t_ref = time.time() - t0
        from cup1d.utils.utils import get_path_repo
        emu_cov = np.load(os.path.join(get_path_repo("lace"), "data", "covariance", "l1O_cov_" + emu.emulator_label + ".npy"), allow_pickle=True).item()
        t0 = time.time()
        res = build_covariance(like.data, like.cov_factor, emu_cov, mode, like.theory.fid_cosmo["cosmo"].dkms_dMpc)
        t_new = time.time() - t0
        same = np.array_equal(res["full_cov_Pk_kms"], like.full_cov_Pk_kms)
        print("mode %-8s N_d=%d  reference set_icov %.2f s   build_covariance %.3f s   (x%.0f)  full cov bit-identical: %s"
              % (mode, nd, t_ref, t_new, t_ref / t_new, same))


if __name__ == "__main__":
    main()
