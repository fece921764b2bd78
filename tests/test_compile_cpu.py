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


def test_synthetic_c3_with_agn_compiles_to_the_oracle():
    """full nuisance set incl. the optional AGN term: compiled tables (NumPy mirror) vs oracle"""
    from cup1d_b200 import synth
    from oracle.model import OracleLikelihood

    emu = synth.synth_nn_emulator(seed=13, suite="nyx", ndeg=6, nhidden=3, max_neurons=24, head=12)
    spec = synth.build_spec(emu, grid="dr1", rebin_k=4, full_cov=True, vary_nrun=True, nz_igm=6, hcd_const=True,
                            gauss_priors=True, agn=True, seed=14, nz=5, nk=20)
    ora = OracleLikelihood(spec)
    xt = synth.truth_point(spec)
    synth.attach_mock_data(spec, np.concatenate(ora.get_p1d_kms(values=xt)[0]))
    ora = OracleLikelihood(spec)
    cfg, F = compile_spec(spec)
    assert cfg["has_agn"] == 1 and cfg["uniform_k"] == 1
    x = np.concatenate([synth.walkers_ball(spec, 4, seed=0), synth.walkers_sweep(spec, 12, seed=3, frac_out=0.1)])
    # switch the AGN term off for some walkers through its null rule (constant coefficient <= -5.5)
    ia = spec["params"]["names"].index("ln_AGN_0")
    x[5, ia] = 0.0
    sim = simulate(cfg, F, x)
    exp = ora.log_prob_and_blobs_many(x)
    ok = exp[:, 0] > -1e99
    assert np.array_equal(sim["lnp"][~ok], exp[~ok, 0])
    assert np.max(np.abs(sim["lnp"][ok] - exp[ok, 0]) / np.maximum(1.0, np.abs(exp[ok, 0]))) < 1e-11
    # the AGN term matters: removing it changes the model
    spec2 = dict(spec)
    spec2["cont"] = dict(spec["cont"])
    spec2["cont"]["agn"] = None
    p_with = np.concatenate(ora.get_p1d_kms(values=x[0])[0])
    p_without = np.concatenate(OracleLikelihood(spec2).get_p1d_kms(values=x[0])[0])
    assert np.max(np.abs(p_with / p_without - 1)) > 1e-4


def test_repeated_nodes_follow_np_interp():
    """a one-redshift data set gives interp_lin families with repeated nodes: the compiled weights
    must reproduce what np.interp (the reference's routine, base_igm.py:253-254) returns for them"""
    import numpy as np

    from cup1d_b200.compile import family_z_weights
    from oracle.model import family_value

    zs = np.array([2.2])
    for nodes in ([2.2, 2.2], [2.2, 2.2, 2.2, 2.2], [2.0, 2.2, 2.2, 3.0]):
        n = len(nodes)
        fam = {"ztype": "interp_lin", "otype": "const", "znodes": np.array(nodes), "coeffs": np.arange(1.0, n + 1.0),
               "param_idx": np.full(n, -1), "z0": 3.0, "null": np.nan, "zmax": np.nan}
        Wz, forced = family_z_weights(fam, zs)
        assert not forced.any()
        np.testing.assert_allclose(Wz @ fam["coeffs"], family_value(fam, zs, np.zeros(0)), rtol=0, atol=1e-15)


def test_derived_grid_spec_reproduces_reference_theory():
    """the derived spec behind B200Theory.get_p1d_kms (redshift subset in any order, custom k grid, no
    rebinning), lowered by the compiler and run through the NumPy mirror of the kernels, against the
    reference's own Theory.get_p1d_kms outputs"""
    import os

    import numpy as np

    import devsim
    from conftest import GOLDEN_DIR
    from cup1d_b200.compile import compile_spec
    from cup1d_b200.spec import load_spec
    from cup1d_b200.theory import grid_spec

    spec, ref = load_spec(os.path.join(GOLDEN_DIR, "theory_kat_chab.npz"), with_extra=True)
    off = np.concatenate([[0], np.cumsum(ref["kc_len"])])
    kc = [ref["kc"][off[i] : off[i + 1]] for i in range(len(ref["kc_len"]))]
    cfg, F = compile_spec(grid_spec(spec, ref["izs"], kc, True))
    x = ref["x"][:8]  # the in-cube rows (the mirror keeps the cube test)
    out = devsim.simulate(cfg, F, x)
    assert np.max(np.abs(out["p1d"] / ref["p_full"][:8] - 1)) < 1e-10


def test_theory_like_params_mapping(golden):
    """B200Theory's translation of like_params (lya_theory.py:610-814 conventions) needs no device: absent
    parameters keep the coefficient their model holds, the resolution term needs an R_coeff parameter in
    the list, a family comes whole or not at all"""
    from types import SimpleNamespace

    import numpy as np
    import pytest

    from cup1d_b200.theory import B200Theory

    spec, _ = golden("c1_chab_nn")
    th = B200Theory(SimpleNamespace(spec=spec, device=None))
    names = list(spec["params"]["names"])
    pmin, pmax = np.asarray(spec["params"]["min"]), np.asarray(spec["params"]["max"])
    x0, res0 = th.cube_from_like_params([])
    assert not res0 and x0.shape == (len(names),)
    held = pmin + x0 * (pmax - pmin)
    for fam in spec["families"].values():
        idx = np.asarray(fam["param_idx"])
        for j in np.nonzero(idx >= 0)[0]:
            assert held[idx[j]] == pytest.approx(float(np.asarray(fam["coeffs"])[j]), rel=1e-14, abs=1e-300)
    assert held[names.index("As")] == pytest.approx(spec["cosmo"]["As_fid"], rel=1e-14)
    # every free parameter, three spellings
    vals = pmin + 0.37 * (pmax - pmin)
    xa, ra = th.cube_from_like_params(vals)
    xb, rb = th.cube_from_like_params(dict(zip(names, vals)))
    xc, rc = th.cube_from_like_params([SimpleNamespace(name=n, value=v) for n, v in zip(names, vals)])
    np.testing.assert_allclose(xa, 0.37, rtol=1e-12)
    np.testing.assert_array_equal(xa, xb)
    np.testing.assert_array_equal(xa, xc)
    assert ra and rb and rc  # R_coeff parameters are free in this baseline
    # only the resolution family: resolution on, everything else held
    rn = [n for n in names if n.startswith("R_coeff")]
    xr, rr = th.cube_from_like_params({n: 0.01 for n in rn})
    assert rr and all(xr[names.index(n)] == x0[names.index(n)] for n in names if n not in rn)
    with pytest.raises(ValueError, match="number of params mismatch"):
        th.cube_from_like_params({rn[0]: 0.01})
    with pytest.raises(ValueError, match="not a free parameter"):
        th.cube_from_like_params({"no_such_parameter": 1.0})


def test_remove_switches_through_a_derived_spec(golden):
    """``remove=`` (si_mult.py:185-206, si_add.py:156-159) is served by a derived spec with the metal switch
    tables overridden: compiled and run through the NumPy mirror, it must match the oracle's remove path"""
    import numpy as np

    import devsim
    from cup1d_b200.compile import compile_spec
    from cup1d_b200.theory import grid_spec
    from oracle.model import OracleLikelihood

    spec, ref = golden("c1_chab_nn")
    ora = OracleLikelihood(spec)
    x = ref["x"][:6]
    nz = len(spec["zs"])
    for remove in ({"SiIII_Lya": 0}, {"SiIIa_Lya": 0, "SiIIb_Lya": 0}, {"SiIIb_SiIIa": 0}, {"SiIII_SiIIa": 0, "SiIIc_Lya": 1}):
        cfg, F = compile_spec(grid_spec(spec, np.arange(nz), None, spec["cont"]["apply_resolution"], remove))
        out = devsim.simulate(cfg, F, x)
        changed = False
        for i, v in enumerate(x):
            want = np.concatenate(ora.get_p1d_kms(values=v.copy(), remove=remove)[0])
            assert np.max(np.abs(out["p1d"][i] / want - 1)) < 1e-11
            changed |= bool(np.max(np.abs(want / ref["p1d"][i] - 1)) > 1e-6)
        assert changed  # the switch does something


@pytest.mark.parametrize("suite", ["mpg", "nyx"])
def test_gp_ktable_norm_compiles_to_the_oracle(suite):
    """GP emulator with the norm of the reference's own uses, np.interp(k_Mpc, input_norm["k_Mpc"], norm_imF(mF))
    (notebooks/figures_CM26/Fig3a.py:83-85): compiled tables (NumPy mirror of the kernels) against the oracle,
    with mean fluxes and wavenumbers beyond the table's nodes"""
    from cup1d_b200 import synth
    from oracle.model import OracleLikelihood

    emu = synth.synth_gp_emulator(seed=5, suite=suite, n_train=60, n_out=5, n_norm_mF=7, n_norm_k=9)
    # a norm grid narrower than the data: both clamps of np.interp are exercised
    emu["norm_k_Mpc"] = np.geomspace(0.2, 2.5, 9)
    emu["norm_mF"] = np.linspace(0.25, 0.7, 7)
    spec = synth.build_spec(emu, grid="dr1", rebin_k=4, full_cov=False, seed=21, nz=4, nk=18, vary_nrun=(suite == "nyx"))
    ora = OracleLikelihood(spec)
    xt = synth.truth_point(spec)
    synth.attach_mock_data(spec, np.concatenate(ora.get_p1d_kms(values=xt)[0]))
    ora = OracleLikelihood(spec)
    cfg, F = compile_spec(spec)
    assert cfg["gp_nm"] == 7 and cfg["gp_nkn"] == 10 and cfg["gp_nnorm"] == 0
    x = np.concatenate([synth.walkers_ball(spec, 4, seed=0), synth.walkers_sweep(spec, 12, seed=3, frac_out=0.1)])
    sim = simulate(cfg, F, x)
    exp = ora.log_prob_and_blobs_many(x)
    ok = exp[:, 0] > -1e99
    assert np.array_equal(sim["lnp"][~ok], exp[~ok, 0])
    assert np.max(np.abs(sim["lnp"][ok] - exp[ok, 0]) / np.maximum(1.0, np.abs(exp[ok, 0]))) < 1e-11
    for i in np.where(ok)[0][:6]:
        po = np.concatenate(ora.get_p1d_kms(values=x[i])[0])
        assert np.max(np.abs(sim["p1d"][i] / po - 1)) < 1e-12
    # the table matters: a flat norm gives another model
    emu2 = dict(emu)
    emu2["norm_table"] = np.ones_like(emu["norm_table"])
    spec2 = dict(spec)
    spec2["emulator"] = emu2
    cfg2, F2 = compile_spec(spec2)
    assert np.max(np.abs(simulate(cfg2, F2, x[:1])["p1d"][0] / sim["p1d"][0] - 1)) > 1e-3
