"""The arithmetic of the tcgen05 dense chi2 (cup1d_b200/csrc/c1d_chi2_i8.cu) restated in NumPy with exact integer
sums (tools/chi2_ozaki_sim.py): five balanced base-256 digit planes per operand, digit pairs of level <= 4,
float64 recombination.  Its distance from an extended-precision d^T C^-1 d (likelihood.py:956-960) on the golden
configurations that carry a dense inverse covariance is the kernel's design tolerance (the GPU tests hold the kernel
itself to 2e-10 relative against the float64 kernel)."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools.chi2_ozaki_sim import residual_rows, sliced_quadratic  # noqa: E402


@pytest.mark.parametrize("name", ["dr1_full_nn", "hcd_const", "nyx_gp_priors"])
def test_five_digit_planes_hold_the_design_tolerance(golden, name):
    spec, extra = golden(name)
    x = np.asarray(extra["x"])[:12]
    d = residual_rows(spec, x)
    assert len(d) >= 6
    A = np.asarray(spec["data"]["full_icov"])
    A = 0.5 * (A + A.T)
    dl = d.astype(np.longdouble)
    ref = np.sum((dl @ A.astype(np.longdouble)) * dl, axis=1).astype(np.float64)
    got = sliced_quadratic(d, A, 5, 5, 5)
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 2e-10
    # four planes (the emulator's count) would not do for the quadratic form: the fifth is what buys the margin
    coarse = sliced_quadratic(d, A, 4, 4, 4)
    assert np.max(np.abs(coarse - ref) / np.abs(ref)) > np.max(np.abs(got - ref) / np.abs(ref))
