import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = [
    "c1_chab_nn",
    "dr1_full_nn",
    "nyx_gp_priors",
    "sivid_pivot",
    "hcd_boss",
    "hcd_const",
    "global_all",
    "chab2019_real",  # the reference's measured Chabanier et al. (2019) data set and covariance
    # further Args.set_baseline variations (tools/make_golden.py:VARIATIONS)
    "var_metal_trad",
    "var_no_contaminants",
    "var_gaikwad21",
    "var_turner24",
    "var_lls_nz4",
    "var_zmin",
    "var_global_igm",
    "var_at_a_time",
    "var_fix_cosmo",
]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    from cup1d_b200.spec import load_spec

    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_spec(os.path.join(GOLDEN_DIR, name + ".npz"), with_extra=True)
        return cache[name]

    return get
