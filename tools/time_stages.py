"""per-stage CUDA-event times of the C4 step (run on the GPU box): python tools/time_stages.py [walkers ...]"""
import os, sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from cup1d_b200 import synth  # noqa: E402
from cup1d_b200.workloads import make_likelihood  # noqa: E402

wl = os.environ.get("WORKLOAD", "c4")
spec, like = make_likelihood(wl, device=0)
like.ctx.set_timing(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for W in [int(a) for a in sys.argv[1:]] or [65536, 8192]:
    x = torch.as_tensor(synth.walkers_sweep(spec, W, seed=16, frac_out=0.01), device="cuda:0")
    for _ in range(3):
        like.log_prob_batch(x)
    rows = []
    for _ in range(9):
        flush.zero_()
        like.log_prob_batch(x)
        torch.cuda.synchronize()
        rows.append(like.ctx.get_timing()[:5])
    m = np.median(np.array(rows), axis=0)
    print("%s W %6d  stages ms: %s  sum %.4f" % (wl, W, " ".join("%.4f" % v for v in m), m.sum()), flush=True)
like.close()
