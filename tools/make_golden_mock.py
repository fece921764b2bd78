"""Golden vectors for mock generation: the reference's own
``BaseMockP1D.get_Pk_iz_perturbed`` (cup1d/p1ds/base_p1d_mock.py:57-75) run in this container.

    python tools/make_golden_mock.py        # writes tests/golden/mock_noise.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from tools import ref_harness as rh  # noqa: E402
from tools.make_golden import synth_cov  # noqa: E402


def main():
    rh.setup()
    # base_p1d_mock imports a LaCE smoothing helper that get_Pk_iz_perturbed never calls
    import types

    for name in ("lace.utils", "lace.utils.smoothing_manager"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["lace.utils.smoothing_manager"].apply_smoothing = None
    from cup1d.p1ds.base_p1d_mock import BaseMockP1D

    rng = np.random.default_rng(11)
    nk = [13, 17, 9, 21]
    Pk = [40.0 * np.exp(-np.linspace(0.2, 2.0, n)) * (1 + 0.05 * rng.normal(0, 1, n)) for n in nk]
    cov = [synth_cov(p, 50 + i) for i, p in enumerate(Pk)]
    out = {"nk": np.array(nk), "Pk": np.concatenate(Pk)}
    for i, c in enumerate(cov):
        out["cov_%d" % i] = c
    for seed in (0, 7):
        pert = BaseMockP1D.get_Pk_iz_perturbed(None, Pk, cov, seed=seed)
        out["perturbed_seed%d" % seed] = np.concatenate(pert)
    pert3 = BaseMockP1D.get_Pk_iz_perturbed(None, Pk, cov, nsamples=3, seed=5)
    out["perturbed3_seed5"] = np.concatenate(pert3, axis=1)
    fn = os.path.join(os.path.dirname(HERE), "tests", "golden", "mock_noise.npz")
    np.savez_compressed(fn, **out)
    print("wrote", fn, os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()
