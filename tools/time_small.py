# small-batch latency: per-call time and per-stage device times (run on the GPU box)
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.getcwd())
from cup1d_b200 import synth
from cup1d_b200.workloads import make_likelihood
for wl, Ws in (("c1", (64, 256)), ("c4", (64, 1024, 4096)), ("c3", (1024,))):
    spec, like = make_likelihood(wl, device=0)
    for W in Ws:
        x = torch.as_tensor(synth.walkers_sweep(spec, W, seed=16, frac_out=0.01), device="cuda:0")
        like.ctx.set_timing(0)
        for _ in range(20): like.log_prob_batch(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        N = 200
        t0 = time.perf_counter(); e0.record()
        for _ in range(N): out = like.log_prob_batch(x)
        e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        like.ctx.set_timing(1)
        for _ in range(5): like.log_prob_batch(x)
        torch.cuda.synchronize()
        like.log_prob_batch(x); torch.cuda.synchronize()
        st = like.ctx.get_timing()[:5]
        print("%s W=%5d  %.1f us/call (events) %.1f us/call (wall)  stages us: %s  prec=%s" % (wl, W, 1e3 * e0.elapsed_time(e1) / N, 1e6 * (t1 - t0) / N, [round(1e3 * v, 1) for v in st], like.emulator_precision))
    like.close()
