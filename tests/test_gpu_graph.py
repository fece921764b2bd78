"""CUDA graphs of small-batch evaluations (c1d_loglike_batch, cup1d_b200/csrc/c1d_abi.cu): a call repeating the
previous call's shape, flags and buffers is captured and replayed; results are bit-identical to the launched
path, and uploads / workspace growth drop the captured graphs."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _eval(like, xd, lnp, blobs, flags):
    import torch

    like.ctx.loglike_batch(xd.data_ptr(), xd.shape[0], lnp.data_ptr(), None, blobs.data_ptr(), None, flags,
                           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return lnp.cpu().numpy().copy(), blobs.cpu().numpy().copy()


@pytest.mark.parametrize("wl,W", [("c1", 64), ("c4", 1024), ("c3", 500)])
def test_replay_equals_launch(wl, W):
    import torch

    from cup1d_b200 import synth
    from cup1d_b200.workloads import make_likelihood

    spec, like = make_likelihood(wl, device=0)
    flags = like._flags()
    x = synth.walkers_sweep(spec, W, seed=3, frac_out=0.05)
    xd = torch.as_tensor(x, device="cuda:0")
    lnp = torch.zeros(W, dtype=torch.float64, device="cuda:0")
    blobs = torch.zeros((W, 6), dtype=torch.float64, device="cuda:0")
    r0 = like.ctx.graph_replays()
    n0 = like.ctx.launch_count()
    a = _eval(like, xd, lnp, blobs, flags)            # launched
    per_call = like.ctx.launch_count() - n0
    outs = [_eval(like, xd, lnp, blobs, flags) for _ in range(4)]  # captured on the second occurrence, then replayed
    assert like.ctx.graph_replays() - r0 == 4
    assert like.ctx.launch_count() - n0 == 5 * per_call
    for b in outs:
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1], equal_nan=True)
    # new inputs in the same buffers: the replay reads them
    x2 = synth.walkers_sweep(spec, W, seed=4, frac_out=0.05)
    xd.copy_(torch.as_tensor(x2, device="cuda:0"))
    c = _eval(like, xd, lnp, blobs, flags)
    assert like.ctx.graph_replays() - r0 == 5
    ref = np.asarray(like.log_prob_batch(x2))        # other buffers: launched
    assert np.array_equal(c[0], ref)
    # a new data vector drops the graphs; the next calls see the new data
    d2 = np.concatenate([np.asarray(v).reshape(-1) for v in spec["data"]["Pk_kms"]]) * 1.01
    like.set_data(d2)
    e = _eval(like, xd, lnp, blobs, flags)
    assert like.ctx.graph_replays() - r0 == 5
    assert not np.array_equal(e[0], c[0])
    f = _eval(like, xd, lnp, blobs, flags)
    g = _eval(like, xd, lnp, blobs, flags)
    assert like.ctx.graph_replays() - r0 == 7
    assert np.array_equal(e[0], f[0]) and np.array_equal(e[0], g[0])
    # workspace growth (a larger batch) also drops them
    big = synth.walkers_sweep(spec, 3 * W + 17, seed=6)
    like.log_prob_batch(big)
    h = _eval(like, xd, lnp, blobs, flags)
    assert like.ctx.graph_replays() - r0 == 7
    assert np.array_equal(h[0], e[0])
    like.close()


def test_device_sampler_uses_graphs():
    """the sampler's half-steps repeat the same workspace buffers: every evaluation after the second is a replay,
    and the chain equals the one sampled with graphs off (same Philox counters)"""
    import os

    from cup1d_b200 import synth
    from cup1d_b200.workloads import make_likelihood

    spec, like = make_likelihood("c1", device=0)
    p0 = synth.walkers_sweep(spec, 64, seed=9)
    r0 = like.ctx.graph_replays()
    out = like.run_sampler_device(p0, 12, seed=5)
    assert like.ctx.graph_replays() - r0 >= 20
    like.close()
    os.environ["C1D_GRAPH"] = "0"
    try:
        import subprocess, sys, json

        code = ("import numpy as np, json, sys; sys.path.insert(0, '.');"
                "from cup1d_b200 import synth; from cup1d_b200.workloads import make_likelihood;"
                "spec, like = make_likelihood('c1', device=0); p0 = synth.walkers_sweep(spec, 64, seed=9);"
                "out = like.run_sampler_device(p0, 12, seed=5); assert like.ctx.graph_replays() == 0;"
                "print(json.dumps(np.asarray(out['lnprob']).tolist()))")
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True)
        ref = np.array(json.loads(res.stdout.strip().splitlines()[-1]))
    finally:
        del os.environ["C1D_GRAPH"]
    assert np.array_equal(np.asarray(out["lnprob"]), ref)
