"""Batched Nelder-Mead and the minimiser driver.

``Fitter.run_minimizer`` (cup1d/likelihood/fitter.py:288-464) calls
``scipy.optimize.minimize(method="Nelder-Mead", bounds=(0, 1)^n)`` once per start: 25 perturbed
starts in its burn-in (:342-380), then repeated polishing runs (:386-441).  Every scipy run
evaluates the likelihood one point at a time.  Here the independent starts advance in lockstep
and each step evaluates the points they need in ONE batched likelihood call
(``B200Likelihood.log_prob_batch``): at most three calls per simplex iteration whatever the
number of starts.

``nelder_mead_batch`` restates scipy's algorithm (scipy/optimize/_optimize.py,
``_minimize_neldermead``, non-adaptive coefficients rho=1, chi=2, psi=0.5, sigma=0.5, the default
initial simplex, bound clipping, the ``maxiter``/``maxfev`` accounting) decision for decision, so
that each start returns what ``scipy.optimize.minimize`` returns for it when the objective gives
the same values (tests/test_minimizer_cpu.py checks x, fun, nit, nfev and status against scipy).
"""

from __future__ import annotations

import numpy as np

RHO, CHI, PSI, SIGMA = 1.0, 2.0, 0.5, 0.5
NONZDELT, ZDELT = 0.05, 0.00025


def initial_simplex(x0, lower, upper):
    """scipy's default simplex around x0 (5 % steps, 0.00025 for zeros), reflected into the bounds"""
    x0 = np.clip(np.asarray(x0, dtype=float), lower, upper)
    n = len(x0)
    sim = np.empty((n + 1, n))
    sim[0] = x0
    for k in range(n):
        y = x0.copy()
        y[k] = (1 + NONZDELT) * y[k] if y[k] != 0 else ZDELT
        sim[k + 1] = y
    sim = np.where(sim > upper, 2 * upper - sim, sim)
    return np.clip(sim, lower, upper)


def _sort(sim, fsim):
    for s in range(len(fsim)):
        ind = np.argsort(fsim[s])  # same call as scipy: identical order on ties
        sim[s] = sim[s][ind]
        fsim[s] = fsim[s][ind]


def nelder_mead_batch(fun_batch, x0, lower=0.0, upper=1.0, fatol=1e-4, xatol=1e-4, maxiter=None, maxfev=None):
    """S independent bounded Nelder-Mead searches advanced in lockstep.

    fun_batch : callable X[M, n] -> f[M]   (rows are independent points, any M >= 1)
    x0        : [S, n] starting points
    Returns dict(x[S, n], fun[S], nit[S], nfev[S], status[S] (0 ok, 1 maxfev, 2 maxiter),
                 success[S], n_batches, n_points) — per start the fields of scipy's OptimizeResult.
    """
    x0 = np.atleast_2d(np.asarray(x0, dtype=float))
    S, n = x0.shape
    lower = np.broadcast_to(np.asarray(lower, dtype=float), (n,))
    upper = np.broadcast_to(np.asarray(upper, dtype=float), (n,))
    if (lower > upper).any():
        raise ValueError("Nelder Mead - one of the lower bounds is greater than an upper bound.")
    if maxiter is None and maxfev is None:
        maxiter = maxfev = n * 200
    elif maxiter is None:
        maxiter = n * 200 if maxfev == np.inf else np.inf
    elif maxfev is None:
        maxfev = n * 200 if maxiter == np.inf else np.inf
    stats = {"n_batches": 0, "n_points": 0}

    def evaluate(points):
        stats["n_batches"] += 1
        stats["n_points"] += len(points)
        return np.asarray(fun_batch(np.ascontiguousarray(points)), dtype=float).reshape(len(points))

    sim = np.stack([initial_simplex(x0[s], lower, upper) for s in range(S)])  # [S, n+1, n]
    fsim = np.full((S, n + 1), np.inf)
    fcalls = np.zeros(S, dtype=np.int64)
    n_init = int(min(n + 1, maxfev))  # the call that would exceed maxfev raises inside scipy's wrapper
    if n_init > 0:
        fsim[:, :n_init] = evaluate(sim[:, :n_init].reshape(S * n_init, n)).reshape(S, n_init)
        fcalls += n_init
    _sort(sim, fsim)
    iters = np.ones(S, dtype=np.int64)
    running = np.ones(S, dtype=bool)

    while True:
        running &= (fcalls < maxfev) & (iters < maxiter)
        for s in np.nonzero(running)[0]:
            if np.max(np.abs(sim[s, 1:] - sim[s, 0])) <= xatol and np.max(np.abs(fsim[s, 0] - fsim[s, 1:])) <= fatol:
                running[s] = False
        act = np.nonzero(running)[0]
        if len(act) == 0:
            break
        # ---- reflection (every running search) ----
        xbar = np.add.reduce(sim[act, :-1], 1) / n
        worst = sim[act, -1]
        xr = np.clip((1 + RHO) * xbar - RHO * worst, lower, upper)
        fxr = evaluate(xr)
        fcalls[act] += 1
        # ---- second point: expansion / outside contraction / inside contraction ----
        f0, fn1, fn = fsim[act, 0], fsim[act, -2], fsim[act, -1]
        expand = fxr < f0
        accept_r = ~expand & (fxr < fn1)
        outside = ~expand & ~accept_r & (fxr < fn)
        inside = ~expand & ~accept_r & ~outside
        x2 = np.empty_like(xr)
        x2[expand] = (1 + RHO * CHI) * xbar[expand] - RHO * CHI * worst[expand]
        x2[outside] = (1 + PSI * RHO) * xbar[outside] - PSI * RHO * worst[outside]
        x2[inside] = (1 - PSI) * xbar[inside] + PSI * worst[inside]
        need2 = expand | outside | inside
        x2 = np.clip(x2, lower, upper)
        # a search that has used up maxfev raises here in scipy: the iteration is abandoned
        broke = need2 & (fcalls[act] >= maxfev)
        do2 = need2 & ~broke
        f2 = np.full(len(act), np.inf)
        if do2.any():
            f2[do2] = evaluate(x2[do2])
            fcalls[act[do2]] += 1
        shrink = np.zeros(len(act), dtype=bool)
        for i, s in enumerate(act):
            if broke[i]:
                continue
            if expand[i]:
                if f2[i] < fxr[i]:
                    sim[s, -1], fsim[s, -1] = x2[i], f2[i]
                else:
                    sim[s, -1], fsim[s, -1] = xr[i], fxr[i]
            elif accept_r[i]:
                sim[s, -1], fsim[s, -1] = xr[i], fxr[i]
            elif outside[i]:
                if f2[i] <= fxr[i]:
                    sim[s, -1], fsim[s, -1] = x2[i], f2[i]
                else:
                    shrink[i] = True
            else:
                if f2[i] < fsim[s, -1]:
                    sim[s, -1], fsim[s, -1] = x2[i], f2[i]
                else:
                    shrink[i] = True
        # ---- shrink towards the best vertex ----
        if shrink.any():
            pts, owner = [], []
            for i in np.nonzero(shrink)[0]:
                s = act[i]
                budget = maxfev - fcalls[s]
                for j in range(1, n + 1):
                    sim[s, j] = np.clip(sim[s, 0] + SIGMA * (sim[s, j] - sim[s, 0]), lower, upper)
                    if j > budget:  # scipy moves the vertex, then its wrapper raises
                        broke[i] = True
                        break
                    pts.append(sim[s, j].copy())
                    owner.append((s, j))
            if pts:
                fp = evaluate(np.array(pts))
                for (s, j), v in zip(owner, fp):
                    fsim[s, j] = v
                    fcalls[s] += 1
        iters[act[~broke]] += 1
        for s in act:
            ind = np.argsort(fsim[s])
            sim[s] = sim[s][ind]
            fsim[s] = fsim[s][ind]

    status = np.where(fcalls >= maxfev, 1, np.where(iters >= maxiter, 2, 0))
    return {
        "x": sim[:, 0].copy(), "fun": fsim.min(axis=1), "nit": iters, "nfev": fcalls, "status": status,
        "success": status == 0, "final_simplex": (sim, fsim), "n_batches": stats["n_batches"], "n_points": stats["n_points"],
    }


def latin_hypercube(n, samples, rng):
    """pyDOE2.lhs(n, samples) (the reference's burn-in design, fitter.py:349): one point per
    stratum and dimension, strata permuted independently per dimension"""
    u = rng.random((samples, n))
    cut = np.linspace(0, 1, samples + 1)
    pts = cut[:samples, None] + u * (cut[1:, None] - cut[:samples, None])
    out = np.empty_like(pts)
    for j in range(n):
        out[:, j] = pts[rng.permutation(samples), j]
    return out


def run_minimizer(like, p0=None, burn_in=False, zmask=None, ind_fix=None, neval=1000, chi2_tol=0.1,
                  nsamples=25, sig=0.25, seed=0, verbose=False):
    """``Fitter.run_minimizer`` (fitter.py:288-464) on a ``B200Likelihood``: optional multi-start
    burn-in (25 Latin-hypercube perturbed starts, all advanced together on the GPU), then
    Nelder-Mead polishing runs until scipy's success flag or two runs without improvement.

    ind_fix : indices of parameters held at their ``p0`` values (the reference's ``mask_pars``)
    Returns dict(mle_cube, chi2, lnprob, n_evals, n_batches, burn_in=[fun of every start],
    burn_in_evals, burn_in_batches).
    """
    npars = like.ndim
    mle_cube = np.full(npars, 0.5) if p0 is None else np.array(p0, dtype=float)
    pfix = None if ind_fix is None else mle_cube[np.asarray(ind_fix)].copy()
    counts = {"evals": 0, "batches": 0}

    def fun_batch(x):
        x = np.array(x, dtype=float)
        if ind_fix is not None:
            x[:, ind_fix] = pfix
        counts["evals"] += len(x)
        counts["batches"] += 1
        return -1.0 * np.asarray(like.log_prob_batch(x, zmask=zmask), dtype=float)

    opts = dict(lower=0.0, upper=1.0, fatol=chi2_tol, xatol=1e-6, maxiter=neval, maxfev=neval)
    burn = None
    if burn_in:
        rng = np.random.default_rng(seed)
        arr_p0 = latin_hypercube(npars, nsamples, rng)
        pini = mle_cube[None, :] + (arr_p0 - 0.5) * sig / (np.arange(nsamples)[:, None] + 1)
        pini[pini <= 0] = 0.05
        pini[pini >= 1] = 0.95
        res = nelder_mead_batch(fun_batch, pini, **opts)
        burn = res["fun"].copy()
        counts["burn_evals"], counts["burn_batches"] = counts["evals"], counts["batches"]
        best = int(np.argmin(res["fun"]))  # first of the smallest, like the strict `<` of fitter.py:375
        mle_cube = res["x"][best].copy()
        if verbose:
            print("burn-in: best start %d, -lnP %.4f" % (best, burn[best]), flush=True)
    chi2 = float(like.get_chi2(mle_cube, zmask=zmask))
    rep = 0
    while True:
        res = nelder_mead_batch(fun_batch, mle_cube[None, :], **opts)
        x = res["x"][0]
        _chi2 = float(like.get_chi2(x, zmask=zmask))
        diff_chi = _chi2 - chi2
        if verbose:
            print("NM run: chi2 %.4f -> %.4f (success %s)" % (chi2, _chi2, bool(res["success"][0])), flush=True)
        if res["success"][0]:
            break
        if -diff_chi > chi2_tol:
            chi2, mle_cube, rep = _chi2, x.copy(), 0
        elif diff_chi < 0:
            chi2, mle_cube, rep = _chi2, x.copy(), rep + 1
        else:
            rep += 1
        if rep >= 2:
            break
    mle_cube = x.copy()
    chi2 = float(like.get_chi2(mle_cube, zmask=zmask))
    return {"mle_cube": mle_cube, "chi2": chi2, "lnprob": float(like.log_prob(mle_cube.copy(), zmask=zmask)),
            "n_evals": counts["evals"], "n_batches": counts["batches"], "burn_in": burn,
            "burn_in_evals": counts.get("burn_evals", 0), "burn_in_batches": counts.get("burn_batches", 0)}
