"""Design study (CPU, NumPy): accuracy of the int8 digit-slice (Ozaki-style) evaluation of the
emulator MLP against the float64 oracle, on the C4 workload.

Activations and weights are cut into balanced base-256 digits relative to a per-row / per-column
scale; digit products of level i + j <= LEVELS - 1 are accumulated exactly (int32 on the tensor
cores, integers here) and recombined in float64.  Prints coefficient, model-P1D and log-likelihood
differences for several digit counts.

usage: python tools/ozaki_sim.py [n_walkers]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from cup1d_b200 import synth, workloads  # noqa: E402
from oracle.model import OracleLikelihood  # noqa: E402
from oracle.emulators import RefMLPEmulator  # noqa: E402


def digits(q, nd):
    """balanced base-256 digits of int64 q (most significant first); q = sum d_i 256^(nd-1-i)"""
    out = []
    for _ in range(nd):
        d = ((q + 128) & 255) - 128
        out.append(d)
        q = (q - d) >> 8
    assert np.all(q == 0), "top digit overflow"
    return out[::-1]


def quantise_rows(x, nd, tight=True):
    """per-row scale; returns integer q with |q| < 2^(8 nd - 1) and the scale s: x ~ q * s"""
    m = np.max(np.abs(x), axis=1, keepdims=True)
    m = np.where(m > 0, m, 1.0)
    full = 2.0 ** (8 * nd - 1) - 2.0 ** (8 * nd - 8)  # top digit stays within +-127 after carries
    if tight:
        s = m / full
    else:
        s = 2.0 ** np.ceil(np.log2(m)) / 2.0 ** (8 * nd - 2)
    q = np.rint(x / s).astype(np.int64)
    return q, s


class SlicedMLP(RefMLPEmulator):
    def __init__(self, spec, nd_a=4, nd_w=4, levels=4, exact_first=1, exact_last=0, **kw):
        super().__init__(spec, **kw)
        self.nd_a, self.nd_w, self.levels = nd_a, nd_w, levels
        self.exact_first, self.exact_last = exact_first, exact_last
        self.wq = []
        for W, b in self.layers:
            q, s = quantise_rows(W, nd_w)  # per output neuron (row of W)
            self.wq.append((digits(q, nd_w), s))

    def coefficients(self, x):
        h = (x - self.xmin) / (self.xmax - self.xmin) - 0.5
        nl = len(self.layers)
        for il, (W, b) in enumerate(self.layers):
            if il < self.exact_first or il >= nl - self.exact_last:
                y = h @ W.T + b
            else:
                q, sa = quantise_rows(h, self.nd_a)
                da = digits(q, self.nd_a)
                dw, sw = self.wq[il]
                tot = np.zeros((h.shape[0], W.shape[0]), dtype=np.float64)
                # exact integer level sums; combined in float64 like the device epilogue
                acc = np.zeros((h.shape[0], W.shape[0]), dtype=np.int64)
                top = (self.nd_a - 1) + (self.nd_w - 1)
                for lev in range(self.levels):
                    g = np.zeros_like(acc)
                    for i in range(self.nd_a):
                        j = lev - i
                        if 0 <= j < self.nd_w:
                            g += da[i] @ dw[j].T
                    assert np.max(np.abs(g)) < 2**31
                    acc += g << (8 * (self.levels - 1 - lev))
                tot = acc.astype(np.float64) * 256.0 ** (top - (self.levels - 1))
                y = tot * sa * sw.T + b
            h = np.where(y > 0, y, self.slope * y) if il < nl - 1 else y
        return h


def main():
    nw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    spec = workloads.build_workload_spec("c4")
    like0 = OracleLikelihood(spec)
    p_truth = np.concatenate(like0.get_p1d_kms(synth.truth_point(spec))[0])
    workloads.finish_workload_spec("c4", spec, p_truth)
    x = synth.walkers_sweep(spec, nw, seed=16)
    ref = OracleLikelihood(spec)
    lnp0 = np.array([ref.log_prob(v) for v in x])
    p0 = [ref.get_p1d_kms(v) for v in x]
    # emulator inputs for the coefficient comparison
    rng = np.random.default_rng(5)
    e = spec["emulator"]
    xe = e["xmin"] + (e["xmax"] - e["xmin"]) * rng.uniform(0.1, 0.9, (4096, len(e["xmin"])))
    c0 = RefMLPEmulator(e).coefficients(xe)
    for nd_a, nd_w, lev in [(4, 4, 4), (4, 4, 5), (4, 5, 4), (4, 5, 5), (5, 5, 5)]:
        emu = SlicedMLP(e, nd_a, nd_w, lev)
        c1 = emu.coefficients(xe)
        tst = OracleLikelihood(spec, emulator=emu)
        lnp1 = np.array([tst.log_prob(v) for v in x])
        ok = lnp0 > -1e90
        dl = np.abs(lnp1[ok] - lnp0[ok])
        rel = 0.0
        for a, b in zip(p0, [tst.get_p1d_kms(v) for v in x]):
            if a is None:
                continue
            rel = max(rel, np.max(np.abs(np.concatenate(b[0]) / np.concatenate(a[0]) - 1)))
        npairs = sum(1 for i in range(nd_a) for j in range(nd_w) if i + j < lev)
        print("digits a=%d w=%d levels=%d pairs=%d: max|dcoef| %.2e  max rel dP %.2e  |dlnL| median %.2e max %.2e"
              % (nd_a, nd_w, lev, npairs, np.max(np.abs(c1 - c0)), rel, np.median(dl), dl.max()))


if __name__ == "__main__":
    main()
