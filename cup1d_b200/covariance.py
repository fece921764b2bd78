"""Setup-side covariance builder.

Vectorised restatement of ``Likelihood.set_icov`` (cup1d/likelihood/likelihood.py:233-507):
per-redshift and full covariance of the P1D data vector with the statistical / systematic
inflation factors of ``cov_factor`` and the emulator leave-one-out covariance added as a
``"diagonal"``, ``"block"`` or ``"full"`` term, then inverted.  The reference fills these
matrices element by element in Python (``np.argmin`` inside double loops over N_d^2 entries:
minutes for the 681-point DR1 vector); here the nearest-neighbour look-ups are computed once per
data point and the matrices are assembled with gathers and outer products.

This is one-off set-up work on the host (NumPy + LAPACK), not part of the per-walker hot path;
its output feeds ``B200Likelihood.set_icov`` / the model spec (``data.icov``, ``data.full_icov``),
i.e. the inverse covariances stage 4 keeps resident in HBM.

The look-up rules are the reference's, quirks included:

* the inflation row of ``cov_factor`` is the one whose ``z`` is nearest to the data redshift
  (first one on ties, ``np.argmin``);
* per-z blocks take the emulator-covariance rows of the nearest tabulated redshift and, within
  them, the nearest tabulated ``k`` [1/Mpc] of every data ``k`` (likelihood.py:347-377);
* the ``"full"`` term finds the nearest ``k`` over the WHOLE stacked ``k_Mpc_zk`` array and uses
  that position as an offset into the rows of the nearest redshift (likelihood.py:432-480): with
  one k-grid per redshift in the table this is the nearest k of that redshift; an offset beyond
  the block raises ``IndexError`` exactly like the reference.
"""

from __future__ import annotations

import numpy as np
from scipy.linalg import block_diag


def _nearest(table, values):
    """np.argmin(np.abs(table - v)) for every v (first index on ties)"""
    table = np.asarray(table, dtype=float)
    values = np.atleast_1d(np.asarray(values, dtype=float))
    return np.argmin(np.abs(table[None, :] - values[:, None]), axis=1)


def _z_rows(zz_zk, z):
    """rows of the stacked emulator covariance that belong to the tabulated redshift nearest to z"""
    ind0 = int(np.argmin(np.abs(zz_zk - z)))
    return np.nonzero(zz_zk == zz_zk[ind0])[0]


def build_covariance(data, cov_factor, emu_cov, emu_cov_type, dkms_dMpc, invert=True):
    """Covariances of ``Likelihood.set_icov`` for one data set.

    Parameters
    ----------
    data : object with the cup1d/p1ds output contract (base_p1d_data.py:90-168): ``z``, ``k_kms``,
        ``Pk_kms``, ``Pksmooth_kms`` (or None), ``cov_Pk_kms``, ``covstat_Pk_kms`` and, for the
        full vector, ``full_Pk_kms``, ``full_zs``, ``full_k_kms``, ``full_cov_Pk_kms``,
        ``full_cov_stat_Pk_kms`` (``full_Pk_kms`` None when absent)
    cov_factor : dict with arrays ``z``, ``val_stat``, ``val_syst``, ``val_emu``, ``val_full``
        (input_pipeline.py:634-655)
    emu_cov : dict of the LaCE file ``l1O_cov_<emulator>.npy``: ``zz_zk``, ``k_Mpc_zk``, ``cov_zk``
    emu_cov_type : "diagonal" | "block" | anything else = "full" (likelihood.py:410-484)
    dkms_dMpc : callable z -> d(km/s)/d(Mpc) of the fiducial cosmology, or array aligned with data.z

    Returns dict(cov_Pk_kms, cov_emu_Pk_kms, icov_Pk_kms  [lists per z],
                 full_cov_Pk_kms, emu_full_cov_Pk_kms, full_icov_Pk_kms  [arrays or None]).
    """
    fz = np.asarray(cov_factor["z"], dtype=float)
    v_stat = np.asarray(cov_factor["val_stat"], dtype=float)
    v_syst = np.asarray(cov_factor["val_syst"], dtype=float)
    v_emu = np.asarray(cov_factor["val_emu"], dtype=float)
    v_full = np.asarray(cov_factor["val_full"], dtype=float)
    zz_zk = np.asarray(emu_cov["zz_zk"], dtype=float)
    k_zk = np.asarray(emu_cov["k_Mpc_zk"], dtype=float)
    cov_zk = np.asarray(emu_cov["cov_zk"], dtype=float)
    zs = np.asarray(data.z, dtype=float)
    if callable(dkms_dMpc):
        conv_of_z = dkms_dMpc
    else:
        table = dict(zip(zs.tolist(), np.asarray(dkms_dMpc, dtype=float).tolist()))
        conv_of_z = table.__getitem__
    pksmooth = data.Pksmooth_kms if getattr(data, "Pksmooth_kms", None) is not None else data.Pk_kms
    diagonal = emu_cov_type == "diagonal"

    out = {"cov_Pk_kms": [], "cov_emu_Pk_kms": [], "icov_Pk_kms": [],
           "full_cov_Pk_kms": None, "emu_full_cov_Pk_kms": None, "full_icov_Pk_kms": None}
    for iz, z in enumerate(zs):
        cov_stat = np.array(data.covstat_Pk_kms[iz], dtype=float)
        cov_syst = np.asarray(data.cov_Pk_kms[iz], dtype=float) - cov_stat
        f = int(np.argmin(np.abs(fz - z)))
        cov = cov_stat * v_stat[f] ** 2 + cov_syst * v_syst[f] ** 2
        # emulator term: relative covariance of the nearest (z, k) nodes times P P'
        k_Mpc = np.asarray(data.k_kms[iz], dtype=float) * conv_of_z(float(z))
        rows = _z_rows(zz_zk, z)
        j = _nearest(k_zk[rows], k_Mpc)
        pk = np.asarray(pksmooth[iz], dtype=float)
        block = cov_zk[np.ix_(rows[j], rows[j])] * pk[:, None] * pk[None, :] * v_emu[f] ** 2
        add = np.diag(np.diag(block)) if diagonal else block
        cov = (cov + add) * v_full[f] ** 2
        out["cov_Pk_kms"].append(cov)
        out["cov_emu_Pk_kms"].append(add)
        if invert:
            out["icov_Pk_kms"].append(np.linalg.inv(cov))

    if getattr(data, "full_Pk_kms", None) is None:
        return out
    full_zs = np.asarray(data.full_zs, dtype=float)
    f = _nearest(fz, full_zs)  # inflation row of every point of the full vector
    cov_stat = np.array(data.full_cov_stat_Pk_kms, dtype=float)
    cov_syst = np.asarray(data.full_cov_Pk_kms, dtype=float) - cov_stat
    # (c * v_i) * v_j, in the reference's order of operations
    cov = cov_stat * v_stat[f][:, None] * v_stat[f][None, :] + cov_syst * v_syst[f][:, None] * v_syst[f][None, :]
    if diagonal:
        full_emu = np.concatenate([np.diag(b) for b in out["cov_emu_Pk_kms"]])
        cov[np.diag_indices_from(cov)] += full_emu
    elif emu_cov_type == "block":
        full_emu = block_diag(*out["cov_emu_Pk_kms"])
        cov = cov + full_emu
    else:
        full_k = np.asarray(data.full_k_kms, dtype=float)
        j = np.empty(len(full_zs), dtype=np.int64)
        for z in np.unique(full_zs):
            sel = full_zs == z
            rows = _z_rows(zz_zk, z)
            off = _nearest(k_zk, full_k[sel] * conv_of_z(float(z)))  # over the whole stacked table
            if np.any(off >= len(rows)):
                raise IndexError("index %d is out of bounds for axis 0 with size %d" % (int(off.max()), len(rows)))
            j[sel] = rows[off]
        pk = np.asarray(data.full_Pk_kms, dtype=float)
        full_emu = cov_zk[np.ix_(j, j)] * pk[:, None] * pk[None, :] * v_emu[f][:, None] * v_emu[f][None, :]
        cov = cov + full_emu
    cov = cov * v_full[f][:, None] * v_full[f][None, :]
    out["full_cov_Pk_kms"] = cov
    out["emu_full_cov_Pk_kms"] = full_emu
    if invert:
        out["full_icov_Pk_kms"] = np.linalg.inv(cov)
    return out


def set_icov(like, emu_cov, dkms_dMpc=None):
    """Drop-in for ``Likelihood.set_icov()`` on a reference ``Likelihood`` object: fills the same
    attributes (likelihood.py:292-309, 385-399, 496-503) from :func:`build_covariance` for the
    main data set and, when present, the extra one."""
    conv = dkms_dMpc if dkms_dMpc is not None else like.theory.fid_cosmo["cosmo"].dkms_dMpc
    res = build_covariance(like.data, like.cov_factor, emu_cov, like.emu_cov_type, conv)
    like.icov_Pk_kms, like.cov_Pk_kms, like.cov_emu_Pk_kms = res["icov_Pk_kms"], res["cov_Pk_kms"], res["cov_emu_Pk_kms"]
    like.full_icov_Pk_kms, like.full_cov_Pk_kms = res["full_icov_Pk_kms"], res["full_cov_Pk_kms"]
    if res["emu_full_cov_Pk_kms"] is not None:
        like.emu_full_cov_Pk_kms = res["emu_full_cov_Pk_kms"]
    like.extra_icov_Pk_kms, like.extra_cov_Pk_kms, like.extra_cov_emu_Pk_kms = [], [], []
    like.extra_full_icov_Pk_kms = like.extra_full_cov_Pk_kms = None
    if getattr(like, "extra_data", None) is not None:
        ext = build_covariance(like.extra_data, like.cov_factor, emu_cov, like.emu_cov_type, conv)
        like.extra_icov_Pk_kms, like.extra_cov_Pk_kms, like.extra_cov_emu_Pk_kms = ext["icov_Pk_kms"], ext["cov_Pk_kms"], ext["cov_emu_Pk_kms"]
        like.extra_full_icov_Pk_kms, like.extra_full_cov_Pk_kms = ext["full_icov_Pk_kms"], ext["full_cov_Pk_kms"]
        if ext["emu_full_cov_Pk_kms"] is not None:
            like.extra_emu_full_cov_Pk_kms = ext["emu_full_cov_Pk_kms"]
    return res
