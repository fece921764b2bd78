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
