"""|dlnL| of the int8 digit-slice emulator against |lnL| on the C4 workload (which walkers carry the tail)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cup1d_b200 import synth
from cup1d_b200.workloads import make_likelihood
W = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
spec, like = make_likelihood("c4", device=0, emulator_precision="fp64")
x = synth.walkers_sweep(spec, W, seed=16)
l64 = like.log_prob_batch(x)
c64 = like.get_chi2_batch(x)
like.set_emulator_precision("i8")
l8 = like.log_prob_batch(x)
ok = l64 > -1e90
dl = np.abs(l8 - l64)[ok]
a = np.abs(l64[ok])
print("walkers", ok.sum(), "|lnL| quantiles", np.quantile(a, [0, 0.01, 0.5, 0.99, 1]))
for lo, hi in ((0, 1e3), (1e3, 1e4), (1e4, 1e5), (1e5, 1e6), (1e6, 1e7), (1e7, 1e99)):
    m = (a >= lo) & (a < hi)
    if m.any():
        print("|lnL| in [%g, %g): n %6d  |dlnL| median %.2e  max %.2e   max dlnL/|lnL| %.2e" % (lo, hi, m.sum(), np.median(dl[m]), dl[m].max(), (dl[m] / a[m]).max()))
for r in (0, 1e-10, 3e-10, 1e-9):
    print("rel floor %g: max dl / (1e-4 + r |lnL|) = %.3f" % (r, (dl / (1e-4 + r * a)).max()))
like.close()
