"""Stage 2 alone on the C4 emulator: float64 (DMMA), 3xTF32 and int8 digit slices, CUDA-event times.
usage: python tools/i8_time.py [walkers]"""
import os, sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cup1d_b200 import _abi  # noqa: E402
from cup1d_b200.workloads import make_likelihood  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
spec, like = make_likelihood("c4", device=0, emulator_precision="fp64")
nrows = W * like.nz
rng = np.random.default_rng(1)
lo, hi = np.asarray(spec["emulator"]["xmin"]), np.asarray(spec["emulator"]["xmax"])
x = np.zeros((nrows, 8))
x[:, : len(lo)] = lo + (hi - lo) * rng.uniform(0.1, 0.9, (nrows, len(lo)))
xd = torch.as_tensor(x, device="cuda:0")
out = torch.zeros((nrows, 8), dtype=torch.float64, device="cuda:0")
ref = None
for name, fl in (("fp64", 0), ("3xtf32", _abi.FLAG_EMU_TC), ("i8", _abi.FLAG_EMU_I8)):
    for _ in range(3):
        like.ctx.emulator_only(xd.data_ptr(), nrows, out.data_ptr(), fl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        like.ctx.emulator_only(xd.data_ptr(), nrows, out.data_ptr(), fl)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    c = out[:, : like.ctx.cfg["nc"]].cpu().numpy()
    if ref is None:
        ref = c
    print("%-7s %8.3f ms  (%d rows; %.1f TFLOP/s algorithmic)  max|coef - fp64| %.2e"
          % (name, ms, nrows, 2 * 94860 * nrows / ms / 1e9, np.max(np.abs(c - ref))))
like.close()
