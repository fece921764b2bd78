"""C5: likelihood-throughput sweep, 2^10 .. 2^20 walkers per step, at the world size it is
launched with (python tools/sweep.py, or torchrun ... tools/sweep.py).  Prints one JSON line per
(walkers, emulator precision) on rank 0; device time from CUDA events, max over ranks."""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from cup1d_b200 import synth
    from cup1d_b200.dist import ShardedEvaluator
    from cup1d_b200.workloads import make_likelihood

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    spec, like = make_likelihood("c4", device=local)
    gather = ShardedEvaluator(None)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    pmax = int(os.environ.get("SWEEP_MAX_LOG2", "20"))
    for lg in range(10, pmax + 1):
        Wtot = 2**lg
        W = Wtot // world  # strong scaling of one step of Wtot walkers
        if W < 1:
            continue
        x = torch.as_tensor(synth.walkers_sweep(spec, W, seed=16 + rank), device=dev)
        for prec in ("i8", "fp64"):  # the default arithmetic of NN emulators and the all-float64 path
            like.set_emulator_precision(prec)
            steps = 5 if lg >= 16 else 20
            for _ in range(3):
                gather.gather_local(like.log_prob_and_blobs_vectorized(x))
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            tot = 0.0
            for _ in range(steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gather.gather_local(like.log_prob_and_blobs_vectorized(x))
                e1.record()
                e1.synchronize()
                tot += e0.elapsed_time(e1)
            t = torch.tensor([tot], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                ms = float(t.item()) / steps
                print(json.dumps({"walkers_total": Wtot, "n_gpus": world, "emulator_precision": prec,
                                  "ms_per_step": ms, "evals_per_s": Wtot / (ms * 1e-3)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    like.close()


if __name__ == "__main__":
    main()
