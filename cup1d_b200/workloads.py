"""Named synthetic workloads = the configurations of BASELINE.json
(SURVEY.md §8d C1-C4), with their seeds."""

from __future__ import annotations

from . import synth

WORKLOADS = {
    # C1 sam_sim-style mock: Cabayol23+-shaped NN emulator, Chabanier k-grid, block-diagonal covariance, 64 walkers
    "c1": dict(desc="C1 sam_sim-style NN mock, 11 z x 35 k (Chabanier grid), block-diag cov, rebin 1", walkers=64),
    # C2 GP emulator, 4096 walkers on one GPU
    "c2": dict(desc="C2 GP emulator (1320 training points), DR1 grid 11 z / 681 k, rebin 8, dense 681^2 icov", walkers=4096),
    # C3 nyx-like full nuisance set, 65k walkers over 8 GPUs
    "c3": dict(desc="C3 nyx-like NN (7 inputs, ndeg 6), As/ns/nrun + 6 IGM nodes + HCD_const + AGN, Gaussian priors, DR1 grid, dense icov", walkers=8192),
    # C4 DESI-DR1-like: NN emulator, rebin 8 (5448 fine k), dense 681^2 inverse covariance, As/ns/nrun
    "c4": dict(desc="C4 DESI-DR1-like mock: Cabayol23+-shaped NN (6-10-70-130-190-250-50-6), 11 z, 681 k, rebin 8 (5448 fine k), dense 681^2 icov, 54 params (As, ns, nrun, IGM, Si, HCD, resolution)", walkers=65536),
}


def build_workload_spec(name):
    """model spec of a workload without data (fill with finish_workload_spec)"""
    if name == "c1":
        emu = synth.synth_nn_emulator(seed=11)
        return synth.build_spec(emu, grid="chab", rebin_k=1, full_cov=False, seed=14)
    if name == "c2":
        emu = synth.synth_gp_emulator(seed=12, n_train=1320)
        return synth.build_spec(emu, grid="dr1", rebin_k=8, full_cov=True, seed=14)
    if name == "c3":
        emu = synth.synth_nn_emulator(seed=13, suite="nyx", ndeg=6)
        return synth.build_spec(emu, grid="dr1", rebin_k=8, full_cov=True, vary_nrun=True, nz_igm=6,
                                hcd_const=True, gauss_priors=True, agn=True, seed=14)
    if name == "c4":
        emu = synth.synth_nn_emulator(seed=11)
        return synth.build_spec(emu, grid="dr1", rebin_k=8, full_cov=True, vary_nrun=True, seed=14)
    raise ValueError("unknown workload " + name)


def finish_workload_spec(name, spec, p_truth):
    """attach the mock data vector and metric: noise-free for C1 (Mock_P1D,
    mock_data.py:96-100), one noise realisation otherwise (seed 15)"""
    return synth.attach_mock_data(spec, p_truth, seed_cov=14, seed_noise=15, add_noise=(name != "c1"))


def make_likelihood(name, device=0, emulator_precision="auto"):
    """(spec, B200Likelihood) with mock data generated on the device"""
    from .likelihood import B200Likelihood

    spec = build_workload_spec(name)
    like = B200Likelihood(spec, device=device, emulator_precision="fp64")
    p_truth = like.get_p1d_kms_batch(synth.truth_point(spec)[None, :])[0]
    like.close()
    finish_workload_spec(name, spec, p_truth)
    return spec, B200Likelihood(spec, device=device, emulator_precision=emulator_precision)
