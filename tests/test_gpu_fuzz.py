"""Randomised differential test of the CUDA path against the oracle over the options of
synth.build_spec (tests/fuzz_parity.py): grids, rebinning factors, 1-13 redshift bins, 3-77 k bins,
IGM nodes, nrun, resolution, HCD_const, IC correction, Gaussian priors, AGN, SiMult / SiVid,
Rogers / BOSS, NN shapes and GP sizes, block or dense covariance; both forms of stage 3 and the
k-segment cuts.  Tolerances: lnP 1e-4 absolute, P1D 1e-5 relative, blobs 1e-12 relative."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [2026, 7])
def test_random_configurations_match_the_oracle(seed, monkeypatch):
    import fuzz_parity

    monkeypatch.setattr(sys, "argv", ["fuzz_parity.py", "10", str(seed)])
    try:
        assert fuzz_parity.main() == 0
    finally:
        for k in ("C1D_P1D_LAYOUT", "C1D_P1D_PARTS"):
            os.environ.pop(k, None)
