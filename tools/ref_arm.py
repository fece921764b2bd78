"""The reference arm of bench.py: the UNMODIFIED reference ``Likelihood`` (cup1d/likelihood/likelihood.py), built
through tools/ref_harness.py from /root/reference or its vendored copy baseline/_ref, for the DESI-DR1-like
configuration the metric is quoted on: 11 redshift bins, the DR1 k grid (681 points, data/zenodo/fig_1.npy shape),
rebin_k = 8, a dense 681^2 inverse covariance, the 53-parameter ``global_opt`` baseline of ``Args.set_baseline``
and a Cabayol23+-shaped NN emulator (the same seeded weights as workload C4; LaCE itself is absent).
Its ``log_prob_and_blobs`` (likelihood.py:1036-1047) is what gets timed, one single-threaded process per core.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_LIKE = {}


def available():
    from tools import ref_harness as rh

    return rh.available()


def build_reference_c4(seed=200):
    """(reference Likelihood, walkers[W, ndim] generator) - takes ~10 s (the reference's set_icov loops)"""
    from tools import ref_harness as rh
    from tools.make_golden import grids, synth_cov, walkers
    from cup1d_b200 import synth
    from oracle.emulators import RefMLPEmulator

    nz = 11
    zs = np.round(2.2 + 0.2 * np.arange(nz), 10)
    ks, kmins, kmaxs = grids("dr1", nz)
    nd = sum(len(k) for k in ks)
    emu = RefMLPEmulator(synth.synth_nn_emulator(seed=11), "synthetic_mpg_nn", "mpg")

    def data_ns(Pk_full, cov_full):
        Pk, cov, off = [], [], 0
        for k in ks:
            Pk.append(Pk_full[off : off + len(k)])
            cov.append(cov_full[off : off + len(k), off : off + len(k)])
            off += len(k)
        return rh.make_data_namespace(zs, ks, kmins, kmaxs, Pk, cov, [0.8 * c for c in cov], full_cov=cov_full, full_cov_stat=0.8 * cov_full)

    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        like0, _ = rh.build_reference_likelihood(emu, data_ns(np.ones(nd), np.eye(nd) * 0.01), rebin_k=8, emu_cov_type="block")
        x_fid = np.clip(like0.sampling_point_from_parameters(), 0.03, 0.97)
        p_truth = np.concatenate(like0.get_p1d_kms(values=x_fid, apply_hull=False)[0])
        cov = synth_cov(p_truth, seed)
        dvec = p_truth + np.linalg.cholesky(cov) @ np.random.default_rng(seed + 1).normal(0, 1, nd)
        like, _ = rh.build_reference_likelihood(emu, data_ns(dvec, cov), rebin_k=8, emu_cov_type="block")
    return like, (lambda n, s=seed + 2: walkers(like, n, 0, 0, s))


def _worker(x):
    like = _LIKE["like"]
    t0 = time.perf_counter()
    out = np.array([like.log_prob_and_blobs(v.copy()) for v in x])
    return out, time.perf_counter() - t0


def throughput(like, x, n_per_core, cores=None):
    """evals/s of the reference's log_prob_and_blobs over `cores` single-threaded forked processes (its production
    layout: one MPI rank per core, OMP_NUM_THREADS=1, scripts/data/data_sampler.py:5)"""
    import multiprocessing as mp

    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits

        _LIKE["limit"] = threadpool_limits(limits=1)
    except Exception:
        pass
    cores = cores or os.cpu_count() or 1
    _LIKE["like"] = like
    chunks = [x[i::cores][:n_per_core] for i in range(cores)]
    chunks = [c for c in chunks if len(c)]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        pool.map(_worker, [c[:2] for c in chunks])  # warm-up
        t0 = time.perf_counter()
        res = pool.map(_worker, chunks)
        wall = time.perf_counter() - t0
    nev = sum(len(r[0]) for r in res)
    return nev / wall, len(chunks), nev, wall


if __name__ == "__main__":
    t0 = time.perf_counter()
    like, wk = build_reference_c4()
    print("built in %.1f s, ndim %d" % (time.perf_counter() - t0, len(like.free_params)))
    x = wk(8 * (os.cpu_count() or 1))
    print("evals/s %.1f on %d cores (%d evals, %.2f s)" % throughput(like, x, 8))
