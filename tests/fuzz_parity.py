"""Randomised differential test: CUDA path (through the C ABI) against the oracle on synthetic
configurations drawn over every option of ``synth.build_spec`` (grid, rebinning factor, number of
redshift bins and k bins, IGM nodes, nrun, resolution, HCD_const, IC correction, Gaussian priors,
AGN, metal / HCD model variants, NN shape or GP, block or dense covariance, convex-hull and star
priors), both forms of stage 3 and every k-segment cut; then the z-masked calls, the per-z
log-likelihoods (with and without log det C) and chi2.  Needs a B200; the oracle is the checker.

    python tests/fuzz_parity.py [n_configs] [seed]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ATOL = 1e-4
RTOL_P1D = 1e-5


def draw(rng):
    kind = rng.choice(["nn", "nn", "gp"])
    suite = rng.choice(["mpg", "nyx"])
    if kind == "nn":
        ndeg = int(rng.choice([4, 5, 6]))
        emu_args = dict(kind=kind, suite=suite, nhidden=int(rng.choice([2, 3, 5])), max_neurons=int(rng.choice([24, 64, 130, 250])),
                        ndeg=ndeg, head=int(rng.choice([8, 24, 50])), seed=int(rng.integers(1, 1000)))
    else:
        emu_args = dict(kind=kind, suite=suite, n_train=int(rng.choice([33, 200, 330])), n_out=int(rng.choice([5, 6])), seed=int(rng.integers(1, 1000)))
    grid = rng.choice(["dr1", "chab"])
    opts = dict(
        grid=grid, rebin_k=int(rng.choice([1, 2, 3, 5, 8])), full_cov=bool(rng.integers(0, 2)), nz=int(rng.choice([1, 2, 5, 11, 13])),
        nz_igm=int(rng.choice([1, 2, 4, 6])), vary_nrun=bool(rng.integers(0, 2)), resolution=bool(rng.integers(0, 2)),
        hcd_const=bool(rng.integers(0, 2)), ic_correction=bool(rng.integers(0, 2)), gauss_priors=bool(rng.integers(0, 2)),
        agn=bool(rng.integers(0, 2)), seed=int(rng.integers(1, 1000)),
        nk=(None if rng.integers(0, 2) else int(rng.choice([3, 9, 40, 77]))),
    )
    variant = dict(metal_model=rng.choice(["SiMult", "SiVid"]), hcd_model=rng.choice(["rogers", "rogers", "boss", "mcdonald"]),
                   hull=bool(rng.integers(0, 2)), star=bool(rng.integers(0, 2)), nzmask=int(rng.integers(1, 4)))
    return emu_args, opts, variant


def synth_priors(spec, ora, x, variant, rng):
    """2-D convex hulls of the emulator calls of half the walkers (hull.py:121-165 builds them from the
    training set) and a star-prior box around the blobs of half the walkers (lya_theory.py:647-654), so
    that both priors reject a share of the batch"""
    from scipy.spatial import ConvexHull

    pv = [spec["params"]["min"] + v * (spec["params"]["max"] - spec["params"]["min"]) for v in x if (v >= 0).all() and (v <= 1).all()]
    pv = pv[: max(8, len(pv) // 2)]
    if variant["hull"]:
        params = list(spec["emulator"]["emu_params"])
        calls = [ora.theory.emulator_calls(np.arange(len(spec["zs"])), p)[0] for p in pv]
        pts = np.concatenate([np.stack([c[k] for k in params], axis=1) for c in calls])
        hulls = []
        for d0 in range(len(params)):
            for d1 in range(d0 + 1, len(params)):
                p2 = pts[:, [d0, d1]]
                ctr = p2.mean(axis=0)
                p2 = ctr + 1.02 * (p2 - ctr) + 1e-9 * rng.normal(size=p2.shape) * (np.abs(ctr) + 1.0)
                try:
                    eq = ConvexHull(p2).equations
                except Exception:  # noqa: BLE001 - a degenerate projection (a fixed emulator input): no hull for it
                    continue
                hulls.append({"dim0": d0, "dim1": d1, "A": eq[:, :2].copy(), "b": eq[:, 2].copy()})
        spec["hull"] = {"params": params, "tol": 1e-12, "hulls": hulls}
    if variant["star"]:
        blobs = np.array([ora.theory.blob(p) for p in pv])
        sp = np.full((3, 2), np.nan)
        for i in range(3):
            if rng.integers(0, 3):
                # edges kept off the walkers' own blobs: those agree with the oracle to 1e-12, not bit for bit
                pad = 1e-9 * (abs(blobs[:, i]).max() + 1.0)
                sp[i] = [blobs[:, i].min() - pad, blobs[:, i].max() + pad]
        spec["cosmo"]["star_priors"] = sp


def main():
    import torch

    from cup1d_b200 import synth
    from cup1d_b200.likelihood import B200Likelihood
    from oracle.model import OracleLikelihood

    n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
    worst = dict(lnp=0.0, p1d=0.0, blob=0.0)
    fails = 0
    for ic in range(n_cfg):
        emu_args, opts, variant = draw(rng)
        ea = dict(emu_args)
        kind = ea.pop("kind")
        emu = synth.synth_nn_emulator(**ea) if kind == "nn" else synth.synth_gp_emulator(**ea)
        tag = "%s %s %s" % (emu_args, opts, variant)
        try:
            spec = synth.build_spec(emu, **opts)
            spec["cont"]["metal_model"] = str(variant["metal_model"])
            spec["cont"]["hcd_model"] = str(variant["hcd_model"])
            if variant["hcd_model"] == "mcdonald":
                synth.use_mcdonald_hcd(spec, n_hcd=int(rng.integers(1, 3)))
            like = B200Likelihood(spec, device=0)
            xt = synth.truth_point(spec)
            p_truth = like.get_p1d_kms_batch(xt[None, :])[0]
            like.close()
            synth.attach_mock_data(spec, p_truth, add_noise=True)
            x = np.concatenate([synth.walkers_ball(spec, 24, seed=ic), synth.walkers_sweep(spec, 40, seed=100 + ic, frac_out=0.1)])
            synth_priors(spec, OracleLikelihood(spec), x, variant, rng)
            # default arithmetic: int8 digit slices for NN emulators, float64 for GP (FUZZ_PREC=fp64 forces float64: triage)
            like = B200Likelihood(spec, device=0, emulator_precision=os.environ.get("FUZZ_PREC", "auto"))
            # int8 digit slices: model error eps <= 3e-9, d lnL <= eps sqrt(2 |lnL|) sqrt(P^T C^-1 P): below 1e-4 near the
            # fit, up to ~2e-9 |lnL| on far-from-fit walkers (|lnL| ~ 1e6..1e7)
            REL = 4e-9 if like.emulator_precision == "i8" else 1e-11
            default_precision = like.emulator_precision
            ora = OracleLikelihood(spec)
            t0 = time.perf_counter()
            exp = ora.log_prob_and_blobs_many(x)
            t_or = time.perf_counter() - t0
            bad = exp[:, 0] <= -1e99
            okrows = np.where(~bad)[0][:12]
            pref = np.stack([np.concatenate(ora.get_p1d_kms(values=x[i].copy())[0]) for i in okrows]) if len(okrows) else None
            for env in ({"C1D_P1D_LAYOUT": "warp"}, {"C1D_P1D_LAYOUT": "lanes", "C1D_P1D_PARTS": "1"}, {"C1D_P1D_LAYOUT": "lanes", "C1D_P1D_PARTS": "8"},
                        {"C1D_P1D_LAYOUT": "split", "C1D_P1D_PARTS": "1"}, {"C1D_P1D_LAYOUT": "split", "C1D_P1D_PARTS": "4"}, {}):
                for k in ("C1D_P1D_LAYOUT", "C1D_P1D_PARTS"):
                    os.environ.pop(k, None)
                os.environ.update(env)
                # the production choice of kernels depends on the batch size: also tile the batch up
                reps = 1 if env else 80
                xx = np.tile(x, (reps, 1))
                got = like.log_prob_and_blobs_vectorized(xx)[: len(x)]
                assert np.array_equal(got[bad, 0], exp[bad, 0]), "rejected rows differ"
                tol = ATOL + REL * np.abs(exp[~bad, 0])
                d = np.abs(got[~bad, 0] - exp[~bad, 0])
                assert np.all(d <= tol), "lnP differs by %g (|lnP| %.3g, tolerance %.3g, %s)" % (d.max(), np.abs(exp[~bad, 0])[np.argmax(d - tol)], tol[np.argmax(d - tol)], env)
                worst["lnp"] = max(worst["lnp"], float((d / np.maximum(1.0, 1e-7 * np.abs(exp[~bad, 0]))).max()) if d.size else 0.0)
                gb, eb = got[~bad, 1:], exp[~bad, 1:]
                np.testing.assert_allclose(gb, eb, rtol=1e-12)
                assert np.array_equal(np.isnan(got[bad, 1:]), np.isnan(exp[bad, 1:]))
                if pref is not None:
                    p = like.get_p1d_kms_batch(np.tile(x[okrows], (reps, 1)))[: len(okrows)]
                    r = float(np.max(np.abs(p / pref - 1)))
                    assert r <= RTOL_P1D, "P1D differs by %g" % r
                    worst["p1d"] = max(worst["p1d"], r)
            for k in ("C1D_P1D_LAYOUT", "C1D_P1D_PARTS"):
                os.environ.pop(k, None)
            # z-masked calls (likelihood.py:844-863: one Theory call per selected z), per-z log-likelihoods
            # with and without log det C, chi2, hull switched off
            zs = np.asarray(spec["zs"])
            zmask = rng.choice(zs, size=min(int(variant["nzmask"]), len(zs)), replace=False)
            sub = x[:20]
            for zm in (None, zmask):
                for ign in (True, False):
                    lnl, lnl_z = like.get_log_like_batch(sub, ignore_log_det_cov=ign, zmask=zm)
                    for i, v in enumerate(sub):
                        e = ora.get_log_like(v.copy(), ignore_log_det_cov=ign, zmask=zm)
                        if e[0] == np.inf:
                            # np.linalg.det of a large inverse covariance overflows in the reference's
                            # log|1/det| (likelihood.py:950, 962); the device path uses slogdet and stays finite
                            assert not ign and np.isfinite(lnl[i])
                            continue
                        if np.isinf(e[0]):
                            assert np.isinf(lnl[i]) and lnl[i] < 0, "null output differs"
                            continue
                        assert abs(lnl[i] - e[0]) <= ATOL + REL * abs(e[0]), "lnL(zmask=%s, ignore=%s) differs by %g" % (zm, ign, abs(lnl[i] - e[0]))
                        assert np.all(np.abs(lnl_z[i] - e[1][0]) <= ATOL + REL * np.abs(e[1][0])), "lnL_z differs"
                lp = like.log_prob_batch(sub, zmask=zm)
                ch = like.get_chi2_batch(sub, zmask=zm)
                for i, v in enumerate(sub):
                    e = ora.log_prob(v.copy(), zmask=zm)
                    assert (lp[i] == e) if e <= -1e99 else (abs(lp[i] - e) <= ATOL + REL * abs(e)), "log_prob(zmask=%s) differs" % (zm,)
                    e2 = ora.get_chi2(v.copy(), zmask=zm)
                    assert (np.isinf(ch[i]) and np.isinf(e2)) or abs(ch[i] - e2) <= 2 * ATOL + 2 * REL * abs(e2), "chi2 differs"
            pn = like.get_p1d_kms_batch(sub, apply_hull=False)
            for i, v in enumerate(sub):
                r = ora.get_p1d_kms(values=v.copy(), apply_hull=False)
                if r is None:
                    assert np.isnan(pn[i]).all(), "apply_hull=False: rejected row has a model"
                else:
                    assert np.max(np.abs(pn[i] / np.concatenate(r[0]) - 1)) <= RTOL_P1D, "apply_hull=False P1D differs"
            if kind == "nn" and pref is not None:
                # the tcgen05 3xTF32 emulator: held to the model-P1D gate (its log-like is fp32-class)
                like.set_emulator_precision("3xtf32")
                for reps in (1, 300):
                    p = like.get_p1d_kms_batch(np.tile(x[okrows], (reps, 1)))
                    for blk in (p[: len(okrows)], p[-len(okrows):]):
                        r = float(np.max(np.abs(blk / pref - 1)))
                        assert r <= RTOL_P1D, "3xTF32 P1D differs by %g" % r
                        worst["tc"] = max(worst.get("tc", 0.0), r)
                like.set_emulator_precision(default_precision)
            like.close()
            print("ok   %2d  nd=%4d ndim=%3d rejected=%2d oracle %.1fs  %s" % (ic, like.nd, like.ndim, int(bad.sum()), t_or, tag), flush=True)
        except Exception as e:  # noqa: BLE001 - report and go on: the point is to find every failing configuration
            fails += 1
            print("FAIL %2d  %s: %s\n     %s" % (ic, type(e).__name__, str(e)[:300], tag), flush=True)
    print("configs %d, failures %d, worst |dlnP| (absolute, or 1e-7 relative beyond 1e7) %.3g, worst P1D relative %.3g (3xTF32 emulator: %.3g)" % (n_cfg, fails, worst["lnp"], worst["p1d"], worst.get("tc", 0.0)))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
