# stage-3 timing at C4 / 65536 for the forms (run on the GPU box)
import os, sys, json, torch, numpy as np
sys.path.insert(0, os.getcwd())
from cup1d_b200 import synth
from cup1d_b200.workloads import make_likelihood
spec, like = make_likelihood("c4", device=0)
for W in (65536, 16384, 8192):
    x = torch.as_tensor(synth.walkers_sweep(spec, W, seed=16, frac_out=0.01), device="cuda:0")
    res = {}
    for layout in ("lanes", "split"):
        os.environ["C1D_P1D_LAYOUT"] = layout
        like.ctx.set_timing(1)
        for _ in range(3): out = like.log_prob_batch(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            out = like.log_prob_batch(x); torch.cuda.synchronize()
            ts.append(like.ctx.get_timing()[2])
        res[layout] = (float(np.median(ts)), out.clone())
    a, b = res["lanes"][1], res["split"][1]
    fin = torch.isfinite(a) & (a > -1e99)
    print(W, {k: v[0] for k, v in res.items()}, "max rel diff", float(((a - b)[fin].abs() / a[fin].abs().clamp(min=1)).max()), "equal rejected", bool(torch.equal(a[~fin], b[~fin])))
