"""Issue-rate probe of tcgen05.mma kind::tf32 (M=128, K=8): cycles per MMA vs N for the
SS / TS operand forms.  Run on a B200: python tools/tc_issue_probe.py"""
import os, sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from cup1d_b200 import _abi

lib = _abi.load_library()
rng = np.random.default_rng(0)
for mode, name in ((100, "SS"), (101, "TS"), (102, "SS,SS,TS")):
    for N in (16, 32, 64, 128, 192, 256):
        A = rng.standard_normal((128, 32)).astype(np.float32)
        B = rng.standard_normal((N, 32)).astype(np.float32)
        D = np.zeros((128, N), dtype=np.float32)
        rc = lib.c1d_tc_selftest(0, A.ctypes.data, B.ctypes.data, D.ctypes.data, N, mode)
        print(f"{name:9s} N={N:3d} rc={rc} cycles/MMA total {D[0,0]:7.1f} issue {D[0,1]:7.1f}")
