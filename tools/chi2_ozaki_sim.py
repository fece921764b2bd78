"""Design study (CPU, NumPy): accuracy of the int8 digit-slice evaluation of the dense
chi2 = d^T C^-1 d (likelihood.py:956-960) against float64 / extended precision.

The residual rows d (per walker) and the rows of C^-1 are cut into balanced base-256 digits relative to
per-row scales; digit products of level i + j < LEVELS are accumulated exactly (int32 on the tensor cores,
integers here) and recombined in float64; y = C^-1 d, then chi2 = sum_j y_j d_j in float64 with the
unquantised d.  Prints the error of chi2 against a longdouble evaluation for several digit counts,
on the C4 workload and on the golden configurations that carry a dense inverse covariance.

usage: python tools/chi2_ozaki_sim.py [n_walkers]
"""
import glob
import sys

import numpy as np

sys.path.insert(0, ".")
from cup1d_b200 import synth, workloads  # noqa: E402
from oracle.model import OracleLikelihood  # noqa: E402
from tools.ozaki_sim import digits, quantise_rows  # noqa: E402


def sliced_quadratic(d, A, nd_d, nd_a, levels, sym=True):
    """chi2 rows: d [W, n], A [n, n] -> [W]"""
    qd, sd = quantise_rows(d, nd_d)
    qa, sa = quantise_rows(A, nd_a)  # per row of A = per output column j of y = A d
    dd, da = digits(qd, nd_d), digits(qa, nd_a)
    acc = np.zeros((d.shape[0], A.shape[0]), dtype=object)
    top = nd_d - 1 + nd_a - 1
    y = np.zeros((d.shape[0], A.shape[0]))
    tot = np.zeros((d.shape[0], A.shape[0]), dtype=np.float64)
    # Horner in float64 over the levels (exact while the running value stays below 2^53)
    for lev in range(levels):
        g = np.zeros((d.shape[0], A.shape[0]), dtype=np.int64)
        for i in range(nd_d):
            j = lev - i
            if 0 <= j < nd_a:
                g += dd[i] @ da[j].T
        assert np.max(np.abs(g)) < 2**31, "int32 overflow at level %d" % lev
        tot = tot * 256.0 + g.astype(np.float64)
    y = tot * 256.0 ** (top - (levels - 1)) * sd * sa.T
    return np.sum(y * d, axis=1)


def report(name, d, A):
    dl = d.astype(np.longdouble)
    ref = np.sum((dl @ A.astype(np.longdouble)) * dl, axis=1)
    f64 = np.sum((d @ A) * d, axis=1)
    amp = np.sum((np.abs(d) @ np.abs(A)) * np.abs(d), axis=1) / np.abs(ref).astype(np.float64)
    print("%s: n = %d, walkers %d, chi2 median %.3g max %.3g, sum|d||A||d| / chi2 median %.3g max %.3g, float64 BLAS: max |dchi2| %.2e"
          % (name, A.shape[0], d.shape[0], np.median(f64), f64.max(), np.median(amp), amp.max(), float(np.max(np.abs(f64 - ref)))))
    for nd_d, nd_a, lev in [(4, 4, 4), (5, 5, 5), (6, 6, 6), (6, 6, 5), (7, 7, 7)]:
        got = sliced_quadratic(d, A, nd_d, nd_a, lev)
        err = np.abs(got - ref).astype(np.float64)
        npairs = sum(1 for i in range(nd_d) for j in range(nd_a) if i + j < lev)
        print("   digits d=%d A=%d levels=%d pairs=%2d: |dchi2| median %.2e max %.2e   relative max %.2e"
              % (nd_d, nd_a, lev, npairs, np.median(err), err.max(), float(np.max(err / np.abs(ref).astype(np.float64)))))


def residual_rows(spec, x):
    ref = OracleLikelihood(spec)
    rows = []
    for v in x:
        p = ref.get_p1d_kms(v)
        if p is None or p[0] is None:
            continue
        rows.append(np.asarray(spec["data"]["full_Pk_kms"]) - np.concatenate([np.asarray(q).reshape(-1) for q in p[0]]))
    return np.array(rows)


def main():
    from cup1d_b200.spec import load_spec

    nw = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    spec = workloads.build_workload_spec("c4")
    like0 = OracleLikelihood(spec)
    p_truth = np.concatenate(like0.get_p1d_kms(synth.truth_point(spec))[0])
    workloads.finish_workload_spec("c4", spec, p_truth)
    x = np.concatenate([synth.walkers_sweep(spec, nw, seed=16), synth.truth_point(spec)[None, :]])
    report("C4 (sweep + truth point)", residual_rows(spec, x), np.asarray(spec["data"]["full_icov"]))
    for f in sorted(glob.glob("tests/golden/*.npz")):
        try:
            gs, extra = load_spec(f, with_extra=True)
        except Exception:  # noqa: BLE001 - not a spec file
            continue
        if gs["data"].get("full_icov") is None:
            continue
        xs = np.asarray(extra["x"]) if "x" in extra else None
        if xs is None:
            print(f, "no sample points", list(extra))
            continue
        d = residual_rows(gs, xs[:nw])
        if len(d):
            report(f, d, np.asarray(gs["data"]["full_icov"]))


if __name__ == "__main__":
    main()
