"""Walker sharding across the GPUs of one box.

Each walker's log-posterior depends only on its own parameter vector and on
read-only constants, so the ensemble is split into contiguous row blocks, one
per rank (one process per GPU, ``torchrun``), every rank evaluates its block
with its own context, and the per-walker results are exchanged with exactly
one all-gather per evaluation (NCCL over NVLink on GPUs; gloo in the CPU
tests).  There is no other collective on the path — the reference's analogue
is fully independent ensembles per MPI rank (fitter.py:216-272).
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(W, world_size, rank):
    """Contiguous split of W rows: every rank gets ceil(W / world) rows except
    possibly the last ones (which may be shorter or empty)."""
    per = (W + world_size - 1) // world_size
    lo = min(rank * per, W)
    hi = min(lo + per, W)
    return lo, hi, per


class ShardedEvaluator(object):
    """evaluate: callable x_local[W_loc, ndim] -> rows[W_loc, ncol] (tensor on
    the rank's device); ``__call__`` returns the gathered [W, ncol] on every rank."""

    def __init__(self, evaluate, ncol=7, group=None):
        self.evaluate = evaluate
        self.ncol = ncol
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def local_rows(self, W):
        return shard_bounds(W, self.world, self.rank)

    def __call__(self, x_all):
        """x_all: the full [W, ndim] ensemble, identical on every rank"""
        W = x_all.shape[0]
        lo, hi, per = self.local_rows(W)
        rows = self.evaluate(x_all[lo:hi])
        if self.world == 1:
            return rows
        # pad the short shards so that one all_gather_into_tensor suffices
        buf = torch.zeros((per, self.ncol), dtype=rows.dtype, device=rows.device)
        buf[: hi - lo] = rows
        out = torch.empty((self.world * per, self.ncol), dtype=rows.dtype, device=rows.device)
        dist.all_gather_into_tensor(out, buf, group=self.group)
        return out[:W]

    def gather_local(self, rows_local):
        """all-gather of equally sized per-rank blocks (weak-scaling benchmark)"""
        if self.world == 1:
            return rows_local
        out = torch.empty((self.world * rows_local.shape[0], rows_local.shape[1]), dtype=rows_local.dtype, device=rows_local.device)
        dist.all_gather_into_tensor(out, rows_local.contiguous(), group=self.group)
        return out


def run_sampler_sharded(like, p0, nsteps, n_ensembles=1, nburn=0, thin=1, seed=0, a=2.0, zmask=None, gather=True, group=None):
    """The device stretch-move sampler over the GPUs of one box (SURVEY.md 8 e / f2).  ``p0[W, ndim]`` holds
    ``n_ensembles`` independent ensembles of W / n_ensembles walkers (rows in ensemble order), identical on every rank.

    * ``n_ensembles`` a multiple of the world size - the reference's production layout, one emcee ensemble per MPI rank
      (fitter.py:216-272; 128 ranks x 107 walkers, launch_data_sampler.sh:69): the ensembles are dealt to the ranks in
      contiguous blocks, every rank advances its own with ``B200Likelihood.run_sampler_device`` and NOTHING is
      exchanged while sampling; the chains are concatenated on the walker axis at the end (fitter.py:262-268) with one
      all-gather per product when ``gather`` (else every rank returns its own block).
    * one ensemble spanning the ranks (``n_ensembles == 1``, world > 1): every rank keeps the whole state, updates its
      share of the active half (``c1d_sampler_half_step``) and the ranks exchange their updated rows with ONE all-gather
      per half-step; the random numbers are keyed on the walker, so the chain equals the single-GPU chain.

    Returns dict(chain[S, W, ndim], lnprob[S, W], blobs[S, W, 6], acceptance_fraction[W]) of device tensors."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = like.device
    x_all = torch.as_tensor(p0, dtype=torch.float64).to(dev).contiguous()
    W, ndim = x_all.shape
    if W % n_ensembles or (W // n_ensembles) % 2:
        raise ValueError("an even number of walkers per ensemble is needed")
    wn = W // n_ensembles

    def cat_walkers(t):
        """all-gather of per-rank blocks [S, W_loc, ...] along the walker axis"""
        if world == 1 or not gather:
            return t
        loc = t.transpose(0, 1).contiguous()  # [W_loc, S, ...]
        out = torch.empty((world * loc.shape[0],) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
        dist.all_gather_into_tensor(out, loc, group=group)
        return out.transpose(0, 1).contiguous()

    if n_ensembles % world == 0:
        e_loc = n_ensembles // world
        lo = rank * e_loc * wn
        res = like.run_sampler_device(x_all[lo : lo + e_loc * wn], nsteps, nburn=nburn, thin=thin, seed=seed, a=a, zmask=zmask,
                                      to_numpy=False, n_ensembles=e_loc, first_ensemble=rank * e_loc)
        out = {"chain": cat_walkers(res["chain"]), "lnprob": cat_walkers(res["lnprob"]), "blobs": cat_walkers(res["blobs"])}
        acc = res["acceptance_fraction"]
        if world > 1 and gather:
            full = torch.empty(world * acc.shape[0], dtype=acc.dtype, device=acc.device)
            dist.all_gather_into_tensor(full, acc.contiguous(), group=group)
            acc = full
        out["acceptance_fraction"] = acc
        return out
    if n_ensembles != 1:
        raise ValueError("n_ensembles must be a multiple of the number of ranks, or 1 (one ensemble spanning the ranks)")
    nh = W // 2
    if nh % world:
        raise ValueError("a spanning ensemble needs W / 2 divisible by the number of ranks")
    ns = nh // world
    s0 = rank * ns
    flags = like._flags(zmask, True)
    nkeep = nsteps // thin
    with torch.cuda.device(dev):
        x = x_all.clone()
        lnp = torch.empty(W, dtype=torch.float64, device=dev)
        blobs = torch.empty((W, 6), dtype=torch.float64, device=dev)
        nacc = torch.zeros(W, dtype=torch.int64, device=dev)
        chain = torch.empty((nkeep, W, ndim), dtype=torch.float64, device=dev)
        lchain = torch.empty((nkeep, W), dtype=torch.float64, device=dev)
        bchain = torch.empty((nkeep, W, 6), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        # initial log-posteriors: every rank its share of the rows, one all-gather
        rows = ShardedEvaluator(lambda xx: like.log_prob_and_blobs_vectorized(xx), group=group)(x)
        lnp.copy_(rows[:, 0])
        blobs.copy_(rows[:, 1:])
        pack = torch.empty((ns, ndim + 8), dtype=torch.float64, device=dev)
        allp = torch.empty((world * ns, ndim + 8), dtype=torch.float64, device=dev)
        stored = 0
        for it in range(nburn + nsteps):
            if it == nburn:
                nacc.zero_()
            for half in (0, 1):
                like.ctx.sampler_half_step(x.data_ptr(), lnp.data_ptr(), blobs.data_ptr(), W, it, half, seed, s0, ns, a=a,
                                           naccept_ptr=nacc.data_ptr(), flags=flags, stream=stream)
                r0 = half * nh + s0
                pack[:, :ndim] = x[r0 : r0 + ns]
                pack[:, ndim] = lnp[r0 : r0 + ns]
                pack[:, ndim + 1 : ndim + 7] = blobs[r0 : r0 + ns]
                pack[:, ndim + 7] = nacc[r0 : r0 + ns].double()
                dist.all_gather_into_tensor(allp, pack, group=group)  # the ONE exchange of the half-step
                h0 = half * nh
                x[h0 : h0 + nh] = allp[:, :ndim]
                lnp[h0 : h0 + nh] = allp[:, ndim]
                blobs[h0 : h0 + nh] = allp[:, ndim + 1 : ndim + 7]
                nacc[h0 : h0 + nh] = allp[:, ndim + 7].long()
            if it >= nburn and (it - nburn + 1) % thin == 0:
                chain[stored].copy_(x)
                lchain[stored].copy_(lnp)
                bchain[stored].copy_(blobs)
                stored += 1
    return {"chain": chain, "lnprob": lchain, "blobs": bchain, "acceptance_fraction": nacc.double() / max(1, nsteps)}
