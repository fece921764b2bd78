# dense chi2 stage time against the walker-group size of the grid order (run on the GPU box)
import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
from cup1d_b200 import synth
from cup1d_b200.workloads import make_likelihood
spec, like = make_likelihood("c4", device=0)
like.ctx.set_timing(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for W in (65536, 16384, 4096):
    x = torch.as_tensor(synth.walkers_sweep(spec, W, seed=16, frac_out=0.01), device="cuda:0")
    row = {}
    for grp in (8, 16, 32, 64, 128, 100000):
        os.environ["C1D_CHI2_GROUP"] = str(grp)
        for _ in range(3): like.log_prob_batch(x)
        ts = []
        for _ in range(7):
            flush.zero_(); like.log_prob_batch(x); torch.cuda.synchronize(); ts.append(like.ctx.get_timing()[3])
        row[grp] = round(float(np.median(ts)), 4)
    print(W, row)
