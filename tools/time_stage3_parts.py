# stage-3 time of the split kernel against the number of k-segments (run on the GPU box)
import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
from cup1d_b200 import synth
from cup1d_b200.workloads import make_likelihood
spec, like = make_likelihood(sys.argv[1] if len(sys.argv) > 1 else "c4", device=0)
like.ctx.set_timing(1)
for W in (65536, 32768, 16384, 8192, 4096, 2048, 1024):
    x = torch.as_tensor(synth.walkers_sweep(spec, W, seed=16, frac_out=0.01), device="cuda:0")
    row = {}
    for layout in ("warp", "lanes", "split", "auto"):
        os.environ["C1D_P1D_LAYOUT"] = layout
        if layout == "auto": del os.environ["C1D_P1D_LAYOUT"]
        for parts in ((0,) if layout in ("warp", "auto") else (1, 2, 4, 8)):
            os.environ["C1D_P1D_PARTS"] = str(parts)
            if not parts: del os.environ["C1D_P1D_PARTS"]
            for _ in range(3): like.log_prob_batch(x)
            ts = []
            for _ in range(7):
                like.log_prob_batch(x); torch.cuda.synchronize(); ts.append(like.ctx.get_timing()[2])
            row[(layout, parts)] = round(float(np.median(ts)), 4)
    print(W, row)
