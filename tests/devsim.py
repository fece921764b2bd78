"""TEST HELPER — NumPy mirror of the device kernels' arithmetic, driven by the
*compiled* context tables (cup1d_b200/compile.py).  It lets the CPU test
suite check the context compiler (slot maps, interpolation weights, k tables,
banded rebinning, nulling rules) against the oracle without a GPU.  It is not
used by the product.
"""

import numpy as np

from cup1d_b200 import compile as cc


def simulate(cfg, F, x, zmask=None, apply_hull=True):
    """Returns dict(lnp[W], blobs[W,6], p1d[W,nd] (NaN when rejected), lnl_z[W,nz], chi2[W], emu[W,nz,n_emu])"""
    x = np.atleast_2d(np.asarray(x, dtype=float))
    W = x.shape[0]
    nz, ndim, nd, nf, n_emu = cfg["nz"], cfg["ndim"], cfg["nd"], cfg["nf"], cfg["n_emu"]
    ell, RW, nc = cfg["ell_width"], cfg["rebin_width"], cfg["nc"]
    zoff_f, zoff_d = F["ZOFF_F"], F["ZOFF_D"]
    tab = F["TAB"].reshape(cc.NCOL, nf)
    zbase = F["ZBASE"].reshape(cc.NQ, nz)
    zwgt = F["ZWGT"].reshape(cc.NQ, nz, ell)
    zidx = F["ZIDX"].reshape(cc.NQ, nz, ell)
    use_full = bool(cfg["has_full"]) and zmask is None
    if zmask is not None:
        zsel = np.asarray(zmask, dtype=bool)
    else:
        zsel = np.ones(nz, dtype=bool)
    out = {
        "lnp": np.zeros(W), "blobs": np.zeros((W, 6)), "p1d": np.full((W, nd), np.nan),
        "lnl_z": np.zeros((W, nz)), "chi2": np.zeros(W), "emu": np.zeros((W, nz, n_emu)),
    }
    icov_blocks = []
    o = 0
    for z in range(nz):
        n = zoff_d[z + 1] - zoff_d[z]
        icov_blocks.append(F["ICOV_BLOCKS"][o : o + n * n].reshape(n, n))
        o += n * n
    for w in range(W):
        xw = x[w]
        if (xw > 1).any() or (xw < 0).any():
            out["lnp"][w] = cfg["min_log_like"]
            out["blobs"][w] = np.nan
            out["lnl_z"][w] = -np.inf
            out["chi2"][w] = np.inf
            continue
        lp = -np.sum((F["PRIOR_X0"] - xw) ** 2 * F["PRIOR_W"]) if cfg["has_gauss"] else 0.0
        pv = F["PMIN"] + xw * F["PSCALE"]
        ratio_As = pv[cfg["idx_As"]] / cfg["As_fid"] if cfg["idx_As"] >= 0 else 1.0
        d_ns = pv[cfg["idx_ns"]] - cfg["ns_fid"] if cfg["idx_ns"] >= 0 else 0.0
        d_nrun = pv[cfg["idx_nrun"]] - cfg["nrun_fid"] if cfg["idx_nrun"] >= 0 else 0.0
        Le, Ls = cfg["L_emu"], cfg["L_star"]
        e_emu = np.exp(np.log(ratio_As) + (d_ns + 0.5 * d_nrun * Le) * Le)
        bf = cfg["blob_fid"]
        blob = np.array([bf[0] * np.exp(np.log(ratio_As) + (d_ns + 0.5 * d_nrun * Ls) * Ls),
                         bf[1] + d_ns + d_nrun * Ls, bf[2] + d_nrun, bf[3], bf[4], bf[5]])
        rej = False
        for i in range(3):
            if blob[i] > cfg["star_hi"][i] or blob[i] < cfg["star_lo"][i]:
                rej = True
        model = np.zeros(nd)
        chi2z = np.zeros(nz)
        for z in range(nz):
            if rej:
                break
            q = np.zeros(cc.NQ)
            for iq in range(cc.NQ):
                lin = zbase[iq, z]
                for e in range(ell):
                    if zidx[iq, z, e] >= 0:
                        lin += zwgt[iq, z, e] * pv[zidx[iq, z, e]]
                v = np.exp(lin) if F["ZOTYPE"][iq] else lin
                q[iq] = 0.0 if lin <= F["ZNULL"][iq] else v
            M = F["MZ"][z]
            fid = F["IGMFID"].reshape(4, nz)
            mF = np.exp(-q[0] * fid[0, z])
            sigT, gamma, kF = q[1] * fid[1, z], q[2] * fid[2, z], q[3] * fid[3, z]
            ev = np.zeros(n_emu)
            linp = F["LINP"].reshape(nz, 3)
            for j in range(n_emu):
                code = cfg["emu_code"][j]
                ev[j] = [linp[z, 0] * e_emu, linp[z, 1] + d_ns + d_nrun * Le, linp[z, 2] + d_nrun, mF,
                         sigT / M, gamma, kF * M, 1000.0 / (kF * M)][code]
            out["emu"][w, z] = ev
            if apply_hull and cfg["n_hull_facets"] > 0 and zsel[z]:
                abc = F["HULL_ABC"].reshape(-1, 3)
                dims = F["HULL_DIM"].reshape(-1, 2)
                val = abc[:, 0] * ev[dims[:, 0]] + abc[:, 1] * ev[dims[:, 1]] + abc[:, 2]
                if not np.all(val <= cfg["hull_tol"]):
                    rej = True
                    break
            # stage 2
            coef = emulate(cfg, F, ev)
            # stage 3
            s = slice(zoff_f[z], zoff_f[z + 1])
            k = tab[cc.COL_K, s]
            pl = np.zeros_like(k)
            for j in range(nc - 1, -1, -1):
                pl = pl * tab[cc.COL_T, s] + coef[j]
            p_emu = tab[cc.COL_S, s] * (10.0**pl if cfg["emu_kind"] == 0 else np.exp(pl))
            if cfg["emu_kind"] == 1 and cfg.get("gp_nm", 0) > 0:
                # k-table norm of the GP emulator: row of this (walker, z) from its mean flux, then the packed
                # (interval, weight) column of the fine grid
                nodes, T = F["GP_NORM_MF"], F["GP_NORM_TAB"].reshape(cfg["gp_nm"], cfg["gp_nkn"])
                j = int(np.clip(np.searchsorted(nodes, mF, side="right") - 1, 0, len(nodes) - 2))
                wgt = float(np.clip((mF - nodes[j]) / (nodes[j + 1] - nodes[j]), 0.0, 1.0))
                row = T[j] + wgt * (T[j + 1] - T[j])
                v = tab[cc.COL_SPARE, s]
                ci = v.astype(int)
                tt = v - ci
                p_emu = p_emu * (row[ci] + tt * (row[np.minimum(ci + 1, cfg["gp_nkn"] - 1)] - row[ci]))
            om = 1 - mF
            if cfg["metal_model"] == 0:
                a3 = q[4] / om
                a2 = cfg["rb3"] * q[6] / om
                r = cfg["ra3_over_rb3"] * q[8]
                gmm = q[9]
                m0 = 1 + (a3 * a3 * cfg["c0_o1"] + a2 * a2 * (r * r * cfg["c0_o2"] + cfg["c0_o3c"]))
                G3 = 2 - 2 / (1 + np.exp(-q[5] * k))
                G2 = 2 - 2 / (1 + np.exp(-q[7] * k))
                mult = m0 + (2 * a3 * G3 * tab[cc.COL_T1, s] + G2 * (2 * a2 * r * tab[cc.COL_T2A, s] + 2 * a2 * tab[cc.COL_T2BC, s])) + (
                    2 * a3 * a2 * gmm * r * tab[cc.COL_T3A, s] + 2 * a3 * a2 * gmm * tab[cc.COL_T3BC, s])
            else:
                mult = 1 + q[4] / om * np.exp(q[5] * k) + q[6] / om * tab[cc.COL_T1, s] * np.exp(-q[7] * k)
            if cfg.get("has_agn"):
                mult = mult * (1 + (q[18] if q[19] > 0 else 0.0) * tab[cc.COL_AGN, s])
            if cfg.get("hcd_gate"):
                q = q.copy()
                q[12] = q[12] if q[13] > 0 else 0.0
                q[13] = 0.0
            hcd = (1 + q[16]) + q[12] * tab[cc.COL_D1, s] + q[13] * tab[cc.COL_D2, s] + q[14] * tab[cc.COL_D3, s] + q[15] * tab[cc.COL_D4, s]
            add = q[10] ** 2 * tab[cc.COL_KT, s] * np.exp(-q[11] ** 2 * k * k)
            res = 1 + (q[17] if cfg["apply_resolution"] else 0.0) * tab[cc.COL_Q, s]
            pf = (hcd * mult * p_emu + add) * res
            ds = slice(zoff_d[z], zoff_d[z + 1])
            st = F["RB_START"][ds]
            wg = F["RB_WGT"].reshape(nd, RW)[ds]
            m = np.zeros(len(st))
            for t in range(RW):
                ii = st + t
                okk = ii < len(pf)
                m[okk] += wg[okk, t] * pf[ii[okk]]
            model[ds] = m
            dz = F["DATA"][ds] - m
            chi2z[z] = dz @ icov_blocks[z] @ dz
        if rej:
            out["lnp"][w] = cfg["min_log_like"] + lp
            out["blobs"][w] = 0.0
            out["lnl_z"][w] = -np.inf
            out["chi2"][w] = np.inf
            continue
        out["p1d"][w] = model
        if use_full:
            d = F["DATA"] - model
            chi2 = d @ F["ICOV_FULL"].reshape(nd, nd) @ d
        else:
            chi2 = np.sum(chi2z[zsel])
        ll = -0.5 * chi2
        if np.isnan(ll):
            out["lnp"][w] = cfg["min_log_like"] + lp
            out["blobs"][w] = 0.0
            out["lnl_z"][w] = -np.inf
            out["chi2"][w] = np.inf
            continue
        out["lnp"][w] = max(cfg["min_log_like"], ll) + lp
        out["blobs"][w] = blob
        out["lnl_z"][w] = np.where(zsel, -0.5 * chi2z, 0.0)
        out["chi2"][w] = chi2
    return out


def emulate(cfg, F, ev):
    nc, n_emu = cfg["nc"], cfg["n_emu"]
    if cfg["emu_kind"] == 0:
        h = (ev - F["EMU_LO"]) / (F["EMU_HI"] - F["EMU_LO"]) - 0.5
        dims = cfg["layer_dims"]
        wo = bo = 0
        for l in range(cfg["n_layers"]):
            K, N = dims[l], dims[l + 1]
            Wt = F["NN_W"][wo : wo + K * N].reshape(K, N)
            b = F["NN_B"][bo : bo + N]
            h = h @ Wt + b
            if l < cfg["n_layers"] - 1:
                h = np.where(h > 0, h, cfg["nn_slope"] * h)
            wo += K * N
            bo += N
        return h
    u = (ev - F["EMU_LO"]) / (F["EMU_HI"] - F["EMU_LO"])
    X = F["GP_X"].reshape(-1, n_emu)
    t = (u[None, :] - X) * F["GP_INVELL"][None, :]
    kern = cfg["gp_sigma2"] * np.exp(-0.5 * np.sum(t * t, axis=1))
    y = F["GP_YMEAN"] + F["GP_YSTD"] * (kern @ F["GP_ALPHA"].reshape(-1, nc))
    mF = ev[cfg["gp_imf"]]
    lnn = 0.0
    for c in F["GP_NORM"][::-1]:
        lnn = lnn * mF + c
    y = y.copy()
    y[0] += lnn
    return y
