"""Small end-to-end run for compute-sanitizer (memcheck): every kernel on a golden configuration."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cup1d_b200.likelihood import B200Likelihood
from cup1d_b200.spec import load_spec

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name in ("dr1_full_nn", "nyx_gp_priors"):
    spec, ref = load_spec(os.path.join(root, "tests", "golden", name + ".npz"), with_extra=True)
    like = B200Likelihood(spec, device=0)
    x = ref["x"]
    out = like.log_prob_and_blobs_vectorized(x)
    lnl, lnlz = like.get_log_like_batch(x)
    p = like.get_p1d_kms_batch(x[:8])
    z = like.log_prob_batch(x, zmask=ref["zmask"]) if name == "dr1_full_nn" else None
    if spec["emulator"]["kind"] == "nn":
        like.set_emulator_precision("3xtf32")
        out2 = like.log_prob_and_blobs_vectorized(x)
    res = like.run_sampler_device(np.clip(x[:40], 0.01, 0.99), 3)
    like.close()
    print(name, "ok", float(np.nanmax(out[:, 0])))
