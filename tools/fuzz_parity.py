"""Randomised differential test: CUDA path (through the C ABI) against the oracle on synthetic
configurations drawn over every option of ``synth.build_spec`` (grid, rebinning factor, number of
redshift bins and k bins, IGM nodes, nrun, resolution, HCD_const, IC correction, Gaussian priors,
AGN, metal / HCD model variants, NN shape or GP, block or dense covariance), both forms of stage 3
and every k-segment cut.  Needs a B200; the oracle is the checker.

    python tools/fuzz_parity.py [n_configs] [seed]
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
    variant = dict(metal_model=rng.choice(["SiMult", "SiVid"]), hcd_model=rng.choice(["rogers", "rogers", "boss"]))
    return emu_args, opts, variant


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
            like = B200Likelihood(spec, device=0)
            xt = synth.truth_point(spec)
            p_truth = like.get_p1d_kms_batch(xt[None, :])[0]
            like.close()
            synth.attach_mock_data(spec, p_truth, add_noise=True)
            like = B200Likelihood(spec, device=0)
            ora = OracleLikelihood(spec)
            x = np.concatenate([synth.walkers_ball(spec, 24, seed=ic), synth.walkers_sweep(spec, 40, seed=100 + ic, frac_out=0.1)])
            t0 = time.perf_counter()
            exp = ora.log_prob_and_blobs_many(x)
            t_or = time.perf_counter() - t0
            bad = exp[:, 0] <= -1e99
            okrows = np.where(~bad)[0][:12]
            pref = np.stack([np.concatenate(ora.get_p1d_kms(values=x[i].copy())[0]) for i in okrows]) if len(okrows) else None
            for env in ({"C1D_P1D_LAYOUT": "warp"}, {"C1D_P1D_LAYOUT": "lanes", "C1D_P1D_PARTS": "1"}, {"C1D_P1D_LAYOUT": "lanes", "C1D_P1D_PARTS": "8"}, {}):
                for k in ("C1D_P1D_LAYOUT", "C1D_P1D_PARTS"):
                    os.environ.pop(k, None)
                os.environ.update(env)
                # the production choice of kernels depends on the batch size: also tile the batch up
                reps = 1 if env else 80
                xx = np.tile(x, (reps, 1))
                got = like.log_prob_and_blobs_vectorized(xx)[: len(x)]
                assert np.array_equal(got[bad, 0], exp[bad, 0]), "rejected rows differ"
                tol = ATOL + 1e-11 * np.abs(exp[~bad, 0])
                d = np.abs(got[~bad, 0] - exp[~bad, 0])
                assert np.all(d <= tol), "lnP differs by %g" % d.max()
                worst["lnp"] = max(worst["lnp"], float((d / np.maximum(1.0, 1e-7 * np.abs(exp[~bad, 0]))).max()) if d.size else 0.0)
                gb, eb = got[~bad, 1:], exp[~bad, 1:]
                np.testing.assert_allclose(gb, eb, rtol=1e-12)
                assert np.array_equal(np.isnan(got[bad, 1:]), np.isnan(exp[bad, 1:]))
                if pref is not None:
                    p = like.get_p1d_kms_batch(np.tile(x[okrows], (reps, 1)))[: len(okrows)]
                    r = float(np.max(np.abs(p / pref - 1)))
                    assert r <= RTOL_P1D, "P1D differs by %g" % r
                    worst["p1d"] = max(worst["p1d"], r)
            like.close()
            print("ok   %2d  nd=%4d ndim=%3d rejected=%2d oracle %.1fs  %s" % (ic, like.nd, like.ndim, int(bad.sum()), t_or, tag), flush=True)
        except Exception as e:  # noqa: BLE001 - report and go on: the point is to find every failing configuration
            fails += 1
            print("FAIL %2d  %s: %s\n     %s" % (ic, type(e).__name__, str(e)[:300], tag), flush=True)
    print("configs %d, failures %d, worst |dlnP| (absolute, or 1e-7 relative beyond 1e7) %.3g, worst P1D relative %.3g" % (n_cfg, fails, worst["lnp"], worst["p1d"]))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
