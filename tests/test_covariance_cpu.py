"""Covariance builder against the reference's own Likelihood.set_icov
(cup1d/likelihood/likelihood.py:233-507), run in the build container by
tools/make_golden_icov.py for the three emulator-covariance modes."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from cup1d_b200.covariance import build_covariance

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(mode):
    g = np.load(os.path.join(GOLD, "icov_%s.npz" % mode), allow_pickle=False)
    nk = g["nk"]
    off = np.concatenate([[0], np.cumsum(nk)])
    sl = [slice(off[i], off[i + 1]) for i in range(len(nk))]
    data = SimpleNamespace(
        z=g["z"], k_kms=[g["k_kms"][s] for s in sl], Pk_kms=[g["Pk_kms"][s] for s in sl], Pksmooth_kms=None,
        cov_Pk_kms=[g["full_cov"][s, s] for s in sl], covstat_Pk_kms=[g["full_cov_stat"][s, s] for s in sl],
        full_Pk_kms=g["Pk_kms"], full_zs=np.concatenate([np.full(n, z) for n, z in zip(nk, g["z"])]), full_k_kms=g["k_kms"],
        full_cov_Pk_kms=g["full_cov"], full_cov_stat_Pk_kms=g["full_cov_stat"],
    )
    cov_factor = {"z": g["factor_z"], "val_stat": g["val_stat"], "val_syst": g["val_syst"], "val_emu": g["val_emu"], "val_full": g["val_full"]}
    emu_cov = {"zz_zk": g["emu_zz_zk"], "k_Mpc_zk": g["emu_k_Mpc_zk"], "cov_zk": g["emu_cov_zk"]}
    return g, data, cov_factor, emu_cov


@pytest.mark.parametrize("mode", ["diagonal", "block", "full"])
def test_matches_reference_set_icov(mode):
    g, data, cov_factor, emu_cov = _load(mode)
    res = build_covariance(data, cov_factor, emu_cov, mode, g["dkms_dMpc"])
    # covariances: same operations in the same order as the reference loops -> bit-exact
    for iz in range(len(data.z)):
        np.testing.assert_array_equal(res["cov_emu_Pk_kms"][iz], g["ref_cov_emu_%d" % iz])
        np.testing.assert_array_equal(res["cov_Pk_kms"][iz], g["ref_cov_%d" % iz])
        # inverses go through LAPACK on both sides: tolerance 1e-9 relative to the largest entry
        ref = g["ref_icov_%d" % iz]
        assert np.max(np.abs(res["icov_Pk_kms"][iz] - ref)) <= 1e-9 * np.max(np.abs(ref))
    ref_emu = g["ref_emu_full_cov"]
    np.testing.assert_array_equal(res["emu_full_cov_Pk_kms"], ref_emu)
    np.testing.assert_array_equal(res["full_cov_Pk_kms"], g["ref_full_cov"])
    ref = g["ref_full_icov"]
    assert np.max(np.abs(res["full_icov_Pk_kms"] - ref)) <= 1e-9 * np.max(np.abs(ref))


def test_modes_differ_and_inflation_is_used():
    """the three fixtures are not degenerate: off-diagonal emulator terms and z-dependent factors matter"""
    gd, *_ = _load("diagonal")
    gb, *_ = _load("block")
    gf, *_ = _load("full")
    assert not np.allclose(gd["ref_full_cov"], gb["ref_full_cov"])
    assert not np.allclose(gb["ref_full_cov"], gf["ref_full_cov"])
    assert np.ptp(gd["val_stat"]) > 0.1 and np.ptp(gd["val_emu"]) > 0.1


def test_callable_conversion_and_no_full_vector():
    g, data, cov_factor, emu_cov = _load("block")
    table = dict(zip(np.asarray(data.z).tolist(), g["dkms_dMpc"].tolist()))
    data.full_Pk_kms = None
    res = build_covariance(data, cov_factor, emu_cov, "block", lambda z: table[z], invert=False)
    assert res["full_cov_Pk_kms"] is None and res["icov_Pk_kms"] == []
    np.testing.assert_array_equal(res["cov_Pk_kms"][2], g["ref_cov_2"])


def test_full_mode_offset_beyond_block_raises_like_reference():
    """likelihood.py:447-452: the nearest k is searched over the whole stacked table; a table whose
    redshift blocks carry different k grids makes that offset overflow the block (IndexError)."""
    g, data, cov_factor, emu_cov = _load("full")
    emu_cov = dict(emu_cov)
    k = emu_cov["k_Mpc_zk"].copy()
    first = emu_cov["zz_zk"] == emu_cov["zz_zk"][0]
    k[first] *= 1e3  # the first block no longer contains the nearest k of any data point
    emu_cov["k_Mpc_zk"] = k
    with pytest.raises(IndexError):
        build_covariance(data, cov_factor, emu_cov, "full", g["dkms_dMpc"])
