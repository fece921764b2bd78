"""Samplers on the GPU: the device stretch move replayed step by step on the
host with the same Philox numbers, and the host-driven vectorised sampler."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _u01(lib, seed, step, half, walker, which):
    buf = (C.c_double * 2)()
    assert lib.c1d_philox_u01(seed, step, half, walker, which, buf) == 0
    return buf[0], buf[1]


def test_device_sampler_matches_host_replay(golden):
    from cup1d_b200 import _abi
    from cup1d_b200.likelihood import B200Likelihood
    from cup1d_b200.sampler import get_initial_walkers

    spec, ref = golden("c1_chab_nn")
    like = B200Likelihood(spec, device=0)
    lib = _abi.load_library()
    ndim, W, nsteps, seed, a = like.ndim, 120, 6, 1234, 2.0
    x0 = np.clip(ref["x"][0], 0.05, 0.95)
    p0 = get_initial_walkers(ndim, W, pini=x0, rng=np.random.default_rng(5))
    out = like.run_sampler_device(p0, nsteps, thin=2, seed=seed, a=a)
    assert out["chain"].shape == (3, W, ndim) and out["lnprob"].shape == (3, W)
    assert out["blobs"].dtype.names[0] == "Delta2_star"

    # host replay: same proposals, same uniforms, likelihood through the batched entry point
    x = p0.copy()
    rows = like.log_prob_and_blobs_vectorized(x)
    lnp, blobs = rows[:, 0].copy(), rows[:, 1:].copy()
    nh = W // 2
    nacc = np.zeros(W)
    kept = []
    for step in range(nsteps):
        for half in range(2):
            q = np.zeros((nh, ndim))
            fac = np.zeros(nh)
            for s in range(nh):
                u1, u2 = _u01(lib, seed, step, half, s, 0)
                zz = ((a - 1.0) * u1 + 1.0) ** 2 / a
                j = min(int(u2 * nh), nh - 1)
                xs, xc = x[half * nh + s], x[(1 - half) * nh + j]
                q[s] = xc - (xc - xs) * zz
                fac[s] = (ndim - 1.0) * np.log(zz)
            rq = like.log_prob_and_blobs_vectorized(q)
            for s in range(nh):
                u3, _ = _u01(lib, seed, step, half, s, 1)
                w = half * nh + s
                if np.log(u3) < fac[s] + rq[s, 0] - lnp[w]:
                    x[w], lnp[w], blobs[w] = q[s], rq[s, 0], rq[s, 1:]
                    nacc[w] += 1
        if (step + 1) % 2 == 0:
            kept.append((x.copy(), lnp.copy()))
    for i, (xx, ll) in enumerate(kept):
        np.testing.assert_allclose(out["chain"][i], xx, rtol=0, atol=1e-12)
        np.testing.assert_allclose(out["lnprob"][i], ll, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(out["acceptance_fraction"], nacc / nsteps)
    assert nacc.sum() > 0
    # the stored log-probabilities are those of the stored positions
    relnp = like.log_prob_batch(out["last_state"])
    np.testing.assert_allclose(relnp, out["last_lnprob"], rtol=1e-12, atol=1e-9)
    like.close()


def test_host_driven_sampler_runs_on_the_batched_likelihood(golden):
    from cup1d_b200.likelihood import B200Likelihood
    from cup1d_b200.sampler import run_sampler

    spec, ref = golden("sivid_pivot")  # 15 parameters
    like = B200Likelihood(spec, device=0)
    x0 = np.clip(ref["x"][0], 0.05, 0.95)
    res = run_sampler(like, pini=x0, nwalkers=40, nburn=20, nsteps=60, thin=3, seed=2)
    assert res["chain"].shape == (20, 40, like.ndim)
    assert res["lnprob"].shape == (20, 40) and res["blobs"].shape == (20, 40)
    assert np.all(np.isfinite(res["lnprob"]))
    # chains move uphill from the initial ball on average and stay inside the cube
    assert res["lnprob"][-1].mean() >= res["lnprob"][0].mean() - 50
    assert np.all((res["chain"] >= 0) & (res["chain"] <= 1))
    # stored lnprob is consistent with a fresh evaluation
    np.testing.assert_allclose(like.log_prob_batch(res["chain"][-1]), res["lnprob"][-1], rtol=1e-12, atol=1e-9)
    like.close()


def test_many_ensembles_in_one_call_equal_the_separate_runs(golden, monkeypatch):
    """n_ensembles independent ensembles advanced together (the reference's production layout, one ensemble per MPI
    rank: fitter.py:216-272) give, bit for bit, the chains each ensemble gives alone under its global number - so
    ensembles can be dealt to GPUs with no communication and concatenated on the walker axis (fitter.py:262-268)"""
    from cup1d_b200.likelihood import B200Likelihood
    from cup1d_b200.sampler import get_initial_walkers

    monkeypatch.setenv("C1D_P1D_LAYOUT", "warp")  # one form of stage 3 whatever the batch size (the forms agree to rounding only)
    spec, ref = golden("sivid_pivot")
    like = B200Likelihood(spec, device=0)
    ndim, wn, E, nsteps, seed = like.ndim, 2 * like.ndim + 2, 3, 7, 99
    x0 = np.clip(ref["x"][0], 0.05, 0.95)
    p0 = get_initial_walkers(ndim, wn * E, pini=x0, rng=np.random.default_rng(8))
    joint = like.run_sampler_device(p0, nsteps, nburn=2, thin=1, seed=seed, n_ensembles=E)
    assert joint["chain"].shape == (nsteps, wn * E, ndim)
    for e in range(E):
        alone = like.run_sampler_device(p0[e * wn : (e + 1) * wn], nsteps, nburn=2, thin=1, seed=seed, n_ensembles=1, first_ensemble=e)
        sl = slice(e * wn, (e + 1) * wn)
        np.testing.assert_array_equal(joint["chain"][:, sl], alone["chain"])
        np.testing.assert_array_equal(joint["lnprob"][:, sl], alone["lnprob"])
        np.testing.assert_array_equal(joint["acceptance_fraction"][sl], alone["acceptance_fraction"])
    # ensembles do not talk: walkers of ensemble 0 are the same whatever ensemble 1 starts from
    p1 = p0.copy()
    p1[wn : 2 * wn] = np.clip(p1[wn : 2 * wn] + 0.01, 0, 1)
    other = like.run_sampler_device(p1, nsteps, nburn=2, thin=1, seed=seed, n_ensembles=E)
    np.testing.assert_array_equal(other["chain"][:, :wn], joint["chain"][:, :wn])
    assert not np.array_equal(other["chain"][:, wn : 2 * wn], joint["chain"][:, wn : 2 * wn])
    like.close()


def test_half_step_in_shares_equals_the_whole_half_step(golden, monkeypatch):
    """c1d_sampler_half_step over sub-ranges of the active half (what the ranks of a spanning ensemble do, one share
    each) reproduces the whole half-step bit for bit"""
    import torch

    from cup1d_b200.likelihood import B200Likelihood
    from cup1d_b200.sampler import get_initial_walkers

    monkeypatch.setenv("C1D_P1D_LAYOUT", "warp")
    spec, ref = golden("sivid_pivot")
    like = B200Likelihood(spec, device=0)
    ndim, W, seed = like.ndim, 64, 5
    x0 = np.clip(ref["x"][0], 0.05, 0.95)
    p0 = get_initial_walkers(ndim, W, pini=x0, rng=np.random.default_rng(3))
    whole = like.run_sampler_device(p0, 3, seed=seed, to_numpy=False)
    x = torch.as_tensor(p0, device="cuda:0").clone()
    rows = like.log_prob_and_blobs_vectorized(x)
    lnp, blobs = rows[:, 0].contiguous(), rows[:, 1:].contiguous()
    nacc = torch.zeros(W, dtype=torch.int64, device="cuda:0")
    nh = W // 2
    shares = [(0, 5), (5, 16), (21, 11)]
    for step in range(3):
        for half in (0, 1):
            for s0, ns in shares:
                like.ctx.sampler_half_step(x.data_ptr(), lnp.data_ptr(), blobs.data_ptr(), W, step, half, seed, s0, ns, naccept_ptr=nacc.data_ptr(),
                                           flags=like._flags(None, True))
        torch.cuda.synchronize()
        assert torch.equal(x, whole["chain"][step])
        assert torch.equal(lnp, whole["lnprob"][step])
    assert sum(n for _, n in shares) == nh
    like.close()
