"""The context compiler (cup1d_b200/compile.py) checked on the CPU: a NumPy
mirror of the kernels' arithmetic (tests/devsim.py) evaluates the compiled
tables and must reproduce the reference outputs stored in the goldens."""

import numpy as np
import pytest

from conftest import GOLDEN_NAMES
from cup1d_b200.compile import compile_spec
from devsim import simulate

RTOL_P1D = 1e-5
ATOL_LOGLIKE = 1e-4


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_compiled_tables_reproduce_reference(golden, name):
    spec, ref = golden(name)
    cfg, F = compile_spec(spec)
    x = ref["x"]
    sim = simulate(cfg, F, x)
    exp = ref["lnprob_blobs"]
    for i in range(len(x)):
        if exp[i, 0] <= -1e99:
            assert sim["lnp"][i] == exp[i, 0]
            np.testing.assert_array_equal(np.isnan(sim["blobs"][i]), np.isnan(exp[i, 1:]))
            assert np.isnan(sim["p1d"][i]).all() == np.isnan(ref["p1d"][i]).all()
        else:
            assert abs(sim["lnp"][i] - exp[i, 0]) <= ATOL_LOGLIKE + 1e-11 * abs(exp[i, 0]), (i, sim["lnp"][i], exp[i, 0])
            np.testing.assert_allclose(sim["blobs"][i], exp[i, 1:], rtol=1e-12)
            assert np.max(np.abs(sim["p1d"][i] / ref["p1d"][i] - 1)) <= 1e-11
            np.testing.assert_allclose(sim["emu"][i], ref["emu_call"][i], rtol=1e-12)
            np.testing.assert_allclose(sim["lnl_z"][i], ref["loglike_all"][i], rtol=1e-9, atol=ATOL_LOGLIKE)
            assert sim["chi2"][i] == pytest.approx(ref["chi2"][i], rel=1e-10, abs=ATOL_LOGLIKE)
    # zmask evaluation (per-z chi2 of the selected bins, hull tested on those bins only)
    if not np.isnan(ref["lnprob_zmask"]).any():
        zsel = np.array([np.any(np.abs(ref["zmask"] - z) < 1e-3) for z in spec["zs"]])
        simz = simulate(cfg, F, x, zmask=zsel)
        for i in range(len(x)):
            e = ref["lnprob_zmask"][i]
            assert abs(simz["lnp"][i] - e) <= ATOL_LOGLIKE + 1e-11 * abs(e)
