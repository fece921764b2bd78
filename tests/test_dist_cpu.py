"""Walker sharding + single all-gather (cup1d_b200/dist.py) on the CPU with the
gloo backend, world size 2: the N>1 plumbing of bench.py / production use."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cup1d_b200.dist import ShardedEvaluator, shard_bounds


def test_shard_bounds_cover_everything():
    for W in (1, 2, 7, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi, per = shard_bounds(W, world, r)
                assert 0 <= lo <= hi <= W and hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(W))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, ndim, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    x = torch.rand((W, ndim), dtype=torch.float64)  # same ensemble on every rank

    def evaluate(xl):  # stand-in for the device evaluation: rows depend only on their own walker
        out = torch.empty((xl.shape[0], 7), dtype=torch.float64)
        out[:, 0] = -0.5 * (xl**2).sum(dim=1)
        out[:, 1:] = xl[:, :6] * 3.0
        return out

    ev = ShardedEvaluator(evaluate)
    full = ev(x)
    ref = evaluate(x)
    ok = bool(torch.equal(full, ref))
    lo, hi, _ = ev.local_rows(W)
    g = ev.gather_local(evaluate(x[:5]))  # equally sized blocks (weak-scaling benchmark path)
    ok = ok and g.shape == (5 * world, 7) and bool(torch.equal(g[:5], g[5:10]))
    q.put((rank, ok, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluator_gloo_world2():
    world, W, ndim = 2, 37, 9  # ragged: 19 + 18 rows
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, ndim, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 19, 19, 37)
