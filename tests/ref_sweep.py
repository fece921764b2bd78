"""Build-container check (needs /root/reference): random combinations of the reference's baselines
(Args.set_baseline name variations, suites, emulator kinds, rebinning factors, redshift counts, grids,
covariance modes, hull / star / Gaussian priors, IC correction) are built as real reference Likelihood
objects through tools/ref_harness.py; on seeded walkers the oracle must reproduce the reference's
log_prob_and_blobs (bit for bit in practice) and the compiled device tables, run through the NumPy mirror
of the kernels (tests/devsim.py), must agree to rounding.

    python tests/ref_sweep.py [n_configs] [seed]
"""
import sys, io, contextlib, numpy as np, traceback
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tools import make_golden as mg
from tools import ref_harness as rh
from oracle.emulators import RefGPEmulator, RefMLPEmulator
from cup1d_b200.spec import spec_from_likelihood
from cup1d_b200.compile import compile_spec
from oracle.model import OracleLikelihood
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
import devsim

rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 77)
vars_ = [None,"metal_trad","metal_si2","metal_deco","metal_thin","DLAs","no_contaminants","less_igm","more_igm","Gaikwad21","Gaikwad21T","Turner24","LLS_nz4","no_res","zmin","zmax","HCD0","HCD_BOSS","Metals_Ma2025"]
nfail=0
for it in range(int(sys.argv[1]) if len(sys.argv)>1 else 24):
    var = vars_[rng.integers(len(vars_))]
    fit_type = "global_opt"
    if rng.integers(4) == 0:
        fit_type, var = [("global_all", None), ("at_a_time_global", None), ("at_a_time_global", "metal_thin"), ("global_igm", None)][rng.integers(4)]
    suite = ["mpg","nyx"][rng.integers(2)]
    gp = bool(rng.integers(2))
    rebin_k = int(rng.choice([1,2,3,5,8]))
    nz = int(rng.choice([3,5,7,11]))
    grid = ["chab","dr1c"][rng.integers(2)]
    kw = dict(use_hull=bool(rng.integers(2)), vary_alphas=(suite=="nyx" and bool(rng.integers(2))), ic_correction=bool(rng.integers(2)),
              prior_Gauss_rms=(0.5 if rng.integers(2) else None), use_ic=False)
    if rng.integers(3)==0: kw["use_star_priors"]={"Delta2_star":[0.2,0.6]}
    full = bool(rng.integers(2))
    name = "%s %s gp=%d rebin=%d nz=%d %s full=%d %s" % (var, suite, gp, rebin_k, nz, grid, full, {k:v for k,v in kw.items() if v})
    try:
        emu = (RefGPEmulator(mg.small_gp(100+it, suite), "synthetic_%s_gp"%suite, suite) if gp else RefMLPEmulator(mg.small_nn(100+it, suite), "synthetic_%s_nn"%suite, suite))
        zs = np.round(2.2 + 0.2 * np.arange(nz), 10)
        ks, kmins, kmaxs = mg.grids(grid, nz)
        nd = sum(len(k) for k in ks)
        def data_ns(Pk_full, cov_full):
            Pk, cov, off = [], [], 0
            for k in ks:
                Pk.append(Pk_full[off:off+len(k)]); cov.append(cov_full[off:off+len(k), off:off+len(k)]); off += len(k)
            if full:
                return rh.make_data_namespace(zs, ks, kmins, kmaxs, Pk, cov, [0.8*c for c in cov], full_cov=cov_full, full_cov_stat=0.8*cov_full)
            return rh.make_data_namespace(zs, ks, kmins, kmaxs, Pk, cov, [0.8*c for c in cov])
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            like0,_ = rh.build_reference_likelihood(emu, data_ns(np.ones(nd), np.eye(nd)*0.01), rebin_k=rebin_k, fit_type=fit_type, name_variation=var, emu_cov_type="block", **kw)
            x_fid = np.clip(like0.sampling_point_from_parameters(), 0.03, 0.97)
            r = like0.get_p1d_kms(values=x_fid, apply_hull=False)
            p_truth = np.concatenate(r[0])
            cov = mg.synth_cov(p_truth, 5)
            dvec = p_truth + np.linalg.cholesky(cov) @ np.random.default_rng(6).normal(0,1,nd)
            like,args = rh.build_reference_likelihood(emu, data_ns(dvec, cov), rebin_k=rebin_k, fit_type=fit_type, name_variation=var, emu_cov_type=("full" if full else "block"), **kw)
            x = mg.walkers(like, 5, 5, 2, 7, width=0.06)
            ref = np.array([like.log_prob_and_blobs(v.copy()) for v in x])
            zm = [zs[0], zs[-1]]
            refz = np.full(len(x), np.nan)
            if not kw.get("ic_correction"):  # the reference itself fails for zmask + ic_correction (model_contaminants.py:319)
                refz = np.array([like.log_prob(v.copy(), zmask=zm) for v in x])
            refp = [like.get_p1d_kms(values=v.copy()) if ((v <= 1).all() and (v >= 0).all()) else None for v in x]
            refll = [like.get_log_like(v.copy(), return_blob=True) for v in x]
        spec = spec_from_likelihood(like)
        ora = OracleLikelihood(spec)
        got = ora.log_prob_and_blobs_many(x)
        ok = ref[:,0] > -1e99
        err = np.max(np.abs(got[ok,0]-ref[ok,0])/np.maximum(1,np.abs(ref[ok,0]))) if ok.any() else 0
        same_rej = np.array_equal(got[~ok,0],ref[~ok,0])
        # the compiled tables through the NumPy mirror of the kernels
        cfg, F = compile_spec(spec)
        sim = devsim.simulate(cfg, F, x)
        err2 = np.max(np.abs(sim["lnp"][ok]-ref[ok,0])/np.maximum(1,np.abs(ref[ok,0]))) if ok.any() else 0
        same2 = np.array_equal(sim["lnp"][~ok], ref[~ok,0])
        # z-masked log-posterior, model P1D and per-z log-likelihoods
        if not np.isnan(refz).any():
            gz = np.array([ora.log_prob(v.copy(), zmask=zm) for v in x])
            sz = devsim.simulate(cfg, F, x, zmask=np.array([int(np.any(np.abs(np.array(zm) - z) < 1e-3)) for z in zs]))["lnp"]
            okz = refz > -1e99
            err = max(err, np.max(np.abs(gz[okz]-refz[okz])/np.maximum(1,np.abs(refz[okz]))) if okz.any() else 0)
            err2 = max(err2, np.max(np.abs(sz[okz]-refz[okz])/np.maximum(1,np.abs(refz[okz]))) if okz.any() else 0)
            same_rej &= np.array_equal(gz[~okz], refz[~okz]); same2 &= np.array_equal(sz[~okz], refz[~okz])
        for i, v in enumerate(x):
            if refp[i] is None:
                continue
            pr = np.concatenate(refp[i][0])
            po = np.concatenate(ora.get_p1d_kms(values=v.copy())[0])
            err = max(err, np.max(np.abs(po/pr-1))); err2 = max(err2, np.max(np.abs(sim["p1d"][i]/pr-1)))
            if np.isfinite(refll[i][0]):
                lo = ora.get_log_like(v.copy(), return_blob=True)
                err = max(err, np.max(np.abs(np.asarray(lo[1]).reshape(-1)[:nz] - np.asarray(refll[i][1]).reshape(-1)[:nz]) / np.maximum(1, np.abs(np.asarray(refll[i][1]).reshape(-1)[:nz]))))
                err2 = max(err2, np.max(np.abs(sim["lnl_z"][i] - np.asarray(refll[i][1]).reshape(-1)[:nz]) / np.maximum(1, np.abs(np.asarray(refll[i][1]).reshape(-1)[:nz]))))
        flag = "ok  " if (err < 1e-10 and same_rej and err2 < 1e-9 and same2) else "BAD "
        if flag != "ok  ": nfail += 1
        print("%s %-100s ndim=%3d ok=%2d oracle %.1e mirror %.1e rej %s %s" % (flag, name, x.shape[1], ok.sum(), err, err2, same_rej, same2), flush=True)
    except Exception as e:
        nfail += 1
        tb = traceback.format_exc().strip().splitlines()
        print("FAIL %-100s %s: %s" % (name, type(e).__name__, str(e)[:150]), flush=True)
        print("     ", " | ".join(tb[-3:])[:300], flush=True)
print("failures", nfail)
