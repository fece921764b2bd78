"""Theory.get_p1d_kms (lya_theory.py:610-814) and Likelihood.get_p1d_kms (likelihood.py:722-785) off the
likelihood's own grid, against outputs of the reference itself (tests/golden/theory_kat_*.npz, made by
tools/make_golden.py:make_theory_kat): redshift subsets in any order, custom k grids, full / partial /
empty like_params, points outside the unit cube, values=None."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

RTOL_P1D = 1e-5


def _split(flat, lens):
    off = np.concatenate([[0], np.cumsum(lens)])
    return [flat[off[i] : off[i + 1]] for i in range(len(lens))]


@pytest.mark.parametrize("name", ["theory_kat_chab", "theory_kat_nyx"])
def test_theory_get_p1d_kms_matches_the_reference(name):
    from cup1d_b200.likelihood import B200Likelihood
    from cup1d_b200.spec import load_spec

    spec, ref = load_spec(os.path.join(GOLDEN_DIR, name + ".npz"), with_extra=True)
    like = B200Likelihood(spec, device=0, emulator_precision="fp64")
    th = like.theory
    x = ref["x"]
    zsub = np.asarray(spec["zs"])[ref["izs"]]
    kc = _split(ref["kc"], ref["kc_len"])
    names = list(spec["params"]["names"])
    keep = [str(n) for n in ref["partial_names"]]
    n_rej = 0
    for i, v in enumerate(x):
        vals = like.parameters_from_sampling_point(v)
        # (a) every free parameter, as LikelihoodParameter-like objects / a flat array / a dict
        lp = [SimpleNamespace(name=n, value=float(t)) for n, t in zip(names, vals)]
        r = th.get_p1d_kms(zsub, kc, like_params=lp, return_emu_params=True)
        if np.isnan(ref["p_full"][i]).all():
            assert r is None
            assert th.get_p1d_kms(zsub, kc, like_params=vals) is None
            n_rej += 1
        else:
            p, blob, emu = r
            assert [len(a) for a in p] == [len(k) for k in kc]
            assert np.max(np.abs(np.concatenate(p) / ref["p_full"][i] - 1)) <= RTOL_P1D
            np.testing.assert_allclose(blob, ref["blob_full"][i], rtol=1e-12)
            for j, nm in enumerate(spec["emulator"]["emu_params"]):
                np.testing.assert_allclose(emu[str(nm)], ref["emu_full"][i, :, j], rtol=1e-12)
            r2 = th.get_p1d_kms(zsub, kc, like_params=vals, return_blob=False)
            np.testing.assert_array_equal(np.concatenate(r2), np.concatenate(p))
            r3 = th.get_p1d_kms(zsub, kc, like_params=dict(zip(names, vals)), return_blob=False)
            np.testing.assert_array_equal(np.concatenate(r3), np.concatenate(p))
        # (b) whole families left out keep their models' coefficients; no R_coeff in the list: no resolution term
        for key, sel in (("p_partial", keep), ("p_res_only", [n for n in names if n.startswith("R_coeff")])):
            r = th.get_p1d_kms(zsub, kc, like_params=[q for q in lp if q.name in sel], return_blob=False)
            if np.isnan(ref[key][i]).all():
                assert r is None
            else:
                assert np.max(np.abs(np.concatenate(r) / ref[key][i] - 1)) <= RTOL_P1D
        # (c) Likelihood.get_p1d_kms on a redshift subset (custom k only without rebinning, as in the reference)
        r = like.get_p1d_kms(zs=zsub, _k_kms=(kc if spec["rebin"] is None else None), values=v)
        if np.isnan(ref["p_like_sub"][i]).all():
            assert r is None
        else:
            assert np.max(np.abs(np.concatenate(r[0]) / ref["p_like_sub"][i] - 1)) <= RTOL_P1D
    assert n_rej == int(np.isnan(ref["p_full"]).all(axis=1).sum())
    # (d) like_params=[] and values=None: the models' own coefficients
    r = th.get_p1d_kms(zsub, kc, like_params=[], return_blob=False)
    assert np.max(np.abs(np.concatenate(r) / ref["p_empty"] - 1)) <= RTOL_P1D
    r = like.get_p1d_kms()
    assert np.max(np.abs(np.concatenate(r[0]) / ref["p_like_none"] - 1)) <= RTOL_P1D
    # one redshift, one bare k array (lya_theory.py:679-685)
    r1 = th.get_p1d_kms(zsub[0], kc[0], like_params=[], return_blob=False)
    np.testing.assert_array_equal(r1[0], r[0] if False else th.get_p1d_kms(zsub[:1], [kc[0]], like_params=[], return_blob=False)[0])
    # (e) the reference's errors: a family given in part, a redshift the likelihood does not have
    fam = next(f for f, d in spec["families"].items() if (np.asarray(d["param_idx"]) >= 0).sum() > 1)
    first = names[int(np.asarray(spec["families"][fam]["param_idx"])[np.asarray(spec["families"][fam]["param_idx"]) >= 0][0])]
    with pytest.raises(ValueError, match="number of params mismatch"):
        th.get_p1d_kms(zsub, kc, like_params={first: 0.1})
    with pytest.raises(ValueError):
        th.get_p1d_kms([7.7], [kc[0]], like_params=[])
    like.close()


def test_values_none_follows_the_reference_rules(golden):
    """values=None = like_params=[] (likelihood.py:749-753): the models' own coefficients, no resolution term,
    no cube test - also when the resolution model holds non-zero coefficients (derived context with the same
    data and metric); the scalar log-det path equals the batched one"""
    import copy

    from cup1d_b200.likelihood import B200Likelihood
    from oracle.model import OracleLikelihood

    spec, ref = golden("c1_chab_nn")
    spec = copy.deepcopy(spec)
    spec["families"]["R_coeff"]["coeffs"] = np.asarray(spec["families"]["R_coeff"]["coeffs"], dtype=float) + 0.03
    like = B200Likelihood(spec, device=0, emulator_precision="fp64")
    spec0 = copy.deepcopy(spec)
    spec0["cont"]["apply_resolution"] = False
    ora0 = OracleLikelihood(spec0)
    for ign in (True, False):
        got = like.get_log_like(None, ignore_log_det_cov=ign, return_blob=True)
        want = ora0.get_log_like(None, ignore_log_det_cov=ign, return_blob=True)
        assert abs(got[0] - want[0]) <= 1e-4 + 1e-11 * abs(want[0])
        np.testing.assert_allclose(got[1][0], want[1][0], rtol=1e-10, atol=1e-4)
        np.testing.assert_allclose(got[2], want[2], rtol=1e-12)
    assert abs(like.get_chi2() - ora0.get_chi2()) <= 2e-4 + 1e-11 * abs(ora0.get_chi2())
    p = np.concatenate(like.get_p1d_kms()[0])
    assert np.max(np.abs(p / np.concatenate(ora0.get_p1d_kms()[0]) - 1)) <= RTOL_P1D
    # scalar log-det path == batched
    ora = OracleLikelihood(spec)
    zm = [spec["zs"][1], spec["zs"][3]]
    for v in ref["x"][:5]:
        for zmask in (None, zm):
            a = like.get_log_like(v.copy(), ignore_log_det_cov=False, zmask=zmask)
            b = like.get_log_like_batch(v[None, :], ignore_log_det_cov=False, zmask=zmask)
            e = ora.get_log_like(v.copy(), ignore_log_det_cov=False, zmask=zmask)
            if np.isinf(e[0]):
                assert np.isinf(a[0])
                continue
            assert a[0] == pytest.approx(float(b[0][0]), rel=1e-14)
            np.testing.assert_allclose(a[1][0], b[1][0], rtol=1e-14)
            assert abs(a[0] - e[0]) <= 1e-4 + 1e-11 * abs(e[0])
    like.close()


def test_set_data_and_set_icov_reach_the_derived_contexts(golden):
    """values=None with a non-zero held R_coeff is served by a derived context that carries the data vector and
    the metric: set_data / set_icov on the likelihood must invalidate it (get_chi2() follows the new data)"""
    import copy

    from cup1d_b200.likelihood import B200Likelihood
    from oracle.model import OracleLikelihood

    spec, ref = golden("c1_chab_nn")
    spec = copy.deepcopy(spec)
    spec["families"]["R_coeff"]["coeffs"] = np.asarray(spec["families"]["R_coeff"]["coeffs"], dtype=float) + 0.03
    like = B200Likelihood(spec, device=0, emulator_precision="fp64")
    before = like.get_chi2()
    new = np.concatenate(spec["data"]["Pk_kms"]) * 1.07
    like.set_data(new)
    spec1 = copy.deepcopy(spec)
    spec1["cont"]["apply_resolution"] = False
    o = 0
    for iz, k in enumerate(spec1["data"]["k_kms"]):
        spec1["data"]["Pk_kms"][iz] = new[o : o + len(k)].copy()
        o += len(k)
    want = OracleLikelihood(spec1).get_chi2()
    after = like.get_chi2()
    assert abs(after - want) <= 2e-4 + 1e-11 * abs(want)
    assert abs(after - before) > 1.0
    icov2 = [1.3 * np.asarray(c) for c in spec1["data"]["icov"]]
    like.set_icov(icov2)
    spec1["data"]["icov"] = icov2
    want2 = OracleLikelihood(spec1).get_chi2()
    assert abs(like.get_chi2() - want2) <= 2e-4 + 1e-11 * abs(want2)
    like.close()
