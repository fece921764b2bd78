"""Known-answer tests: the reference's published figure data
(data/zenodo/fig_5.npy, fig_6.npy, copied verbatim as .npz) against the
oracle's contaminant templates.  Inputs are those of
notebooks/figures_CM26/Fig5.py:104-186 and Fig6.py:40-66."""

import os

import numpy as np

from conftest import GOLDEN_DIR
from cup1d_b200.synth import ROGERS, silicon_constants
from oracle import model as om


def _pivot(vals, null):
    return {"ztype": "pivot", "otype": "exp", "znodes": np.zeros(0), "coeffs": np.array(vals, dtype=float),
            "param_idx": np.full(len(vals), -1), "z0": 3.0, "null": null, "zmax": 10.0}


def test_fig5_simult_siadd():
    ref = np.load(os.path.join(GOLDEN_DIR, "zenodo_fig_5.npz"))
    k = np.linspace(1e-3, 0.04, 1000)
    np.testing.assert_array_equal(ref["x"], k)
    si_mult, si_add = silicon_constants()
    z = np.array([3.0])
    mF = np.array([0.75])
    coeffs = {"f_Lya_SiIII": [0, -4.17], "s_Lya_SiIII": [0, 4.90], "f_Lya_SiII": [0, -3.66],
              "s_Lya_SiII": [0, 5.65], "f_SiIIa_SiIII": [0, 0.84], "f_SiIIb_SiIII": [0, 0.59]}
    vals = {n: om.nulled_value(_pivot(c, -20.0), z, None) for n, c in coeffs.items()}
    remove = [
        {"SiIII_Lya": 1, "SiIIa_Lya": 0, "SiIIb_Lya": 0, "SiIII_SiIIa": 0, "SiIII_SiIIb": 0},
        {"SiIII_Lya": 0, "SiIIa_Lya": 1, "SiIIb_Lya": 1, "SiIII_SiIIa": 0, "SiIII_SiIIb": 0},
        {"SiIII_Lya": 0, "SiIIa_Lya": 0, "SiIIb_Lya": 0, "SiIII_SiIIa": 1, "SiIII_SiIIb": 1},
    ]
    for ii in range(3):
        got = om.si_mult_contamination(si_mult, vals, [k], mF, remove=remove[ii])[0]
        assert np.max(np.abs(got - ref["y%d" % ii])) <= 4e-16
    coeffs = {"f_SiIIa_SiIIb": [0, 0.40], "s_SiIIa_SiIIb": [0, 4.43]}
    vals = {n: om.nulled_value(_pivot(c, -20.0), z, None) for n, c in coeffs.items()}
    got = om.si_add_contamination(si_add, vals, [k], remove={})[0]
    assert np.max(np.abs(got / ref["y3"] - 1)) <= 1e-15


def test_fig6_hcd_rogers():
    ref = np.load(os.path.join(GOLDEN_DIR, "zenodo_fig_6.npz"))
    k = ref["x"]  # np.logspace(-3, log10(0.04), 100) as stored by the reference
    np.testing.assert_allclose(k, np.logspace(-3, np.log10(0.04), 100), rtol=1e-15)
    z = np.array([3.0])
    amps = [0.1, 3e-2, 1e-2, 1e-2]
    for ii in range(4):
        vals = {}
        for it in range(4):
            c = [0, np.log(amps[ii])] if it == ii else [0, -11.5]
            vals["HCD_damp%d" % (it + 1)] = om.nulled_value(_pivot(c, -21.5), z, None)
        fam_c = _pivot([0, 0], np.nan)
        fam_c["otype"] = "const"
        fam_c["zmax"] = np.nan
        vals["HCD_const"] = om.nulled_value(fam_c, z, None)
        got = om.hcd_rogers_contamination(ROGERS, vals, z, [k])[0]
        assert np.max(np.abs(got - ref["y%d" % ii])) <= 4e-16


def test_agn_template_against_reference_model():
    """the reference's own AGN_Model.get_contamination (tools/make_golden.py agn_kat)"""
    from cup1d_b200.synth import AGN_TABLE

    ref = np.load(os.path.join(GOLDEN_DIR, "agn_kat.npz"))
    np.testing.assert_array_equal(ref["agn_z"], AGN_TABLE[:, 0])
    np.testing.assert_array_equal(ref["agn_expansion"], AGN_TABLE[:, 1:])
    agn = {"z": ref["agn_z"], "expansion": ref["agn_expansion"]}
    k = ref["k"]
    for case, exp in zip(ref["cases"], ref["out"]):
        z, n = case[0], int(case[1])
        fam = {"ztype": "pivot", "otype": "exp", "znodes": np.zeros(0), "coeffs": case[2 : 2 + n],
               "param_idx": np.full(n, -1), "z0": float(ref["z0"]), "null": float(ref["null"]), "zmax": np.nan}
        got = om.agn_contamination(agn, fam, np.array([z]), [k], None)[0]
        assert np.max(np.abs(got - exp)) <= 1e-15


def test_hcd_mcdonald2005_against_reference_model():
    """the reference's own HCD_Model_McDonald2005.get_contamination (tools/make_golden.py mcdonald_kat): nulled and
    live amplitudes, one and two coefficients, coefficients from like_params"""
    ref = np.load(os.path.join(GOLDEN_DIR, "mcdonald_kat.npz"))
    k = ref["k"]
    n_null = 0
    for case, exp in zip(ref["cases"], ref["out"]):
        z, n = case[0], int(case[1])
        fam = {"ztype": "pivot", "otype": "exp", "znodes": np.zeros(0), "coeffs": case[2 : 2 + n],
               "param_idx": np.full(n, -1), "z0": float(ref["z0"]), "null": float(ref["null"]), "zmax": np.nan}
        got = om.hcd_mcdonald_contamination(fam, np.array([z]), [k], None)[0]
        assert np.max(np.abs(got - exp)) <= 4e-16 * np.max(np.abs(exp))
        n_null += int(np.all(exp == 1.0))
    assert 0 < n_null < len(ref["cases"])


def test_hcd_mcdonald2005_compiled_tables():
    """a synthetic configuration switched to HCD_Model_McDonald2005: the compiled tables (through the NumPy mirror of
    the kernels) against the oracle, walkers on both sides of the null value"""
    from cup1d_b200 import compile as cc
    from cup1d_b200 import synth
    from oracle.model import OracleLikelihood
    import devsim

    emu = synth.synth_nn_emulator(seed=5, nhidden=2, max_neurons=16, head=8)
    for n_hcd in (1, 2):
        spec = synth.build_spec(emu, grid="chab", rebin_k=2, full_cov=False, nz=4, nz_igm=2, seed=9)
        synth.use_mcdonald_hcd(spec, n_hcd=n_hcd)
        ora = OracleLikelihood(spec)
        p_truth = np.concatenate(ora.get_p1d_kms(values=synth.truth_point(spec))[0])
        synth.attach_mock_data(spec, p_truth)
        ora = OracleLikelihood(spec)
        x = synth.walkers_ball(spec, 12, seed=3)
        i0 = spec["params"]["names"].index("HCD_damp1_0")
        x[:4, i0] = 0.05  # ln_A_damp_0 = -6.5: below the null value, HCD term = 1
        exp = ora.log_prob_and_blobs_many(x)
        cfg, F = cc.compile_spec(spec)
        got = devsim.simulate(cfg, F, x)
        ok = exp[:, 0] > -1e99
        assert ok.sum() >= 8
        np.testing.assert_allclose(got["lnp"][ok], exp[ok, 0], rtol=1e-10, atol=1e-8)
        # the nulled walkers differ from the live ones: the term is in the model
        x2 = x[:1].copy()
        x2[0, i0] = 0.8
        assert abs(ora.log_prob_and_blobs_many(x2)[0, 0] - exp[0, 0]) > 1e-3
