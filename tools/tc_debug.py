import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from cup1d_b200.workloads import make_likelihood
from cup1d_b200 import _abi
spec, like = make_likelihood("c4", device=0)
nrows = 128*148*4
x = torch.rand((nrows,8), dtype=torch.float64, device='cuda')*0.5+0.25
out = torch.zeros((nrows,8), dtype=torch.float64, device='cuda')
for i in range(2):
    like.ctx.emulator_only(x.data_ptr(), nrows, out.data_ptr(), _abi.FLAG_EMU_TC)
torch.cuda.synchronize()
