"""Dense chi2 stage alone on C4: float64 (DMMA) against tcgen05 int8 digit slices, per batch size (run on the GPU box).
usage: python tools/time_chi2_i8.py"""
import os, sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from cup1d_b200 import synth  # noqa: E402
from cup1d_b200.workloads import make_likelihood  # noqa: E402

spec, like = make_likelihood("c4", device=0)
like.ctx.set_timing(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for W in (65536, 16384, 8192, 1024, 64):
    x = torch.as_tensor(synth.walkers_sweep(spec, W, seed=16, frac_out=0.01), device="cuda:0")
    row = {}
    for mode in ("fp64", "i8"):
        like.set_chi2_precision(mode)
        for _ in range(3):
            like.log_prob_batch(x)
        ts, tot = [], []
        for _ in range(7):
            flush.zero_()
            like.log_prob_batch(x)
            torch.cuda.synchronize()
            t = like.ctx.get_timing()
            ts.append(t[3])
            tot.append(sum(t[:5]))
        row[mode] = (round(float(np.median(ts)), 4), round(float(np.median(tot)), 4))
    print("W %6d  chi2 stage / whole step [ms]:  fp64 %s   i8 %s" % (W, row["fp64"], row["i8"]), flush=True)
like.close()
