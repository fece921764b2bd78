"""Multi-GPU checks on real NCCL ranks (run with torchrun on 2+ GPUs of one box; prints one JSON line):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py

1. the gathered [W, 7] rows of a sharded evaluation equal rank 0's own evaluation of the whole batch, bit for bit;
2. the device sampler with the ensembles dealt to the ranks (no communication while sampling) equals the single-GPU run;
3. one ensemble spanning the ranks (one all-gather per half-step) equals the single-GPU chain.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["C1D_P1D_LAYOUT"] = "warp"  # one form of stage 3 at every batch size: bit-identical rows
    from cup1d_b200 import synth
    from cup1d_b200.dist import ShardedEvaluator, run_sampler_sharded
    from cup1d_b200.sampler import get_initial_walkers
    from cup1d_b200.workloads import make_likelihood

    spec, like = make_likelihood("c1", device=local)
    dev = like.device
    out = {"world": world}
    # 1. sharded evaluation
    x = torch.as_tensor(synth.walkers_sweep(spec, 4096 + 8 * world, seed=3), device=dev)
    rows = ShardedEvaluator(lambda xx: like.log_prob_and_blobs_vectorized(xx))(x)
    own = like.log_prob_and_blobs_vectorized(x)
    out["gather_bit_identical"] = bool(torch.equal(torch.nan_to_num(rows, nan=-7.0), torch.nan_to_num(own, nan=-7.0)))
    # 2. ensembles dealt to the ranks
    ndim = like.ndim
    wn, E = 2 * ndim + 2, 2 * world
    p0 = get_initial_walkers(ndim, wn * E, pini=np.clip(synth.truth_point(spec), 0.05, 0.95), rng=np.random.default_rng(8))
    sharded = run_sampler_sharded(like, p0, 5, n_ensembles=E, nburn=1, seed=11)
    single = like.run_sampler_device(p0, 5, nburn=1, seed=11, n_ensembles=E, to_numpy=False)
    out["ensembles_sharded_equal_single_gpu"] = bool(torch.equal(sharded["chain"], single["chain"]) and torch.equal(sharded["lnprob"], single["lnprob"]))
    # 3. one ensemble spanning the ranks
    W1 = 32 * world
    p1 = get_initial_walkers(ndim, W1, pini=np.clip(synth.truth_point(spec), 0.05, 0.95), rng=np.random.default_rng(9))
    span = run_sampler_sharded(like, p1, 4, n_ensembles=1, nburn=1, seed=12)
    one = like.run_sampler_device(p1, 4, nburn=1, seed=12, to_numpy=False)
    out["spanning_ensemble_equals_single_gpu"] = bool(torch.equal(span["chain"], one["chain"]) and torch.equal(span["lnprob"], one["lnprob"])
                                                      and torch.equal(span["acceptance_fraction"], one["acceptance_fraction"]))
    flags = torch.tensor([int(out[k]) for k in ("gather_bit_identical", "ensembles_sharded_equal_single_gpu", "spanning_ensemble_equals_single_gpu")], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out["all_ranks_agree"] = bool(flags.min().item() == 1)
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    like.close()
    sys.exit(0 if out["all_ranks_agree"] else 1)


if __name__ == "__main__":
    main()
