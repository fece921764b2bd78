"""a few C4 evaluations at one batch size with the digit-slice chi2 (profiling target)"""
import os, sys
import torch
sys.path.insert(0, os.getcwd())
from cup1d_b200 import synth
from cup1d_b200.workloads import make_likelihood
W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
spec, like = make_likelihood("c4", device=0)
x = torch.as_tensor(synth.walkers_sweep(spec, W, seed=16, frac_out=0.01), device="cuda:0")
for _ in range(4):
    like.log_prob_batch(x)
torch.cuda.synchronize()
like.close()
