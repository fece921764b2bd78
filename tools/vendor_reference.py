"""Copy the reference's Python package and the few data files its likelihood set-up reads into
baseline/_ref/ (git-ignored, not gpurun-ignored: it travels to the GPU box with the snapshot), so that
`bench.py --impl reference` can time the UNMODIFIED reference's own Likelihood.log_prob_and_blobs
(cup1d/likelihood/likelihood.py:1036-1047) on the box's host cores.  Nothing is modified; LaCE / mpi4py /
matplotlib are stubbed at import time by tools/ref_harness.py exactly as for the golden fixtures.

usage: python tools/vendor_reference.py [source = /root/reference]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
# what Args.set_baseline / Likelihood read through get_path_repo("cup1d") (utils/utils.py:100-147)
DATA = ("data/ics", "data/nuisance")


def vendor(src="/root/reference"):
    if not os.path.isdir(os.path.join(src, "cup1d")):
        return False
    os.makedirs(DEST, exist_ok=True)
    dst_pkg = os.path.join(DEST, "cup1d")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    shutil.copytree(os.path.join(src, "cup1d"), dst_pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "old"))
    for d in DATA:
        t = os.path.join(DEST, d)
        if os.path.isdir(t):
            shutil.rmtree(t)
        shutil.copytree(os.path.join(src, d), t)
    with open(os.path.join(DEST, "VENDORED_FROM"), "w") as f:
        f.write(src + "\n")
    return True


if __name__ == "__main__":
    ok = vendor(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("vendored into %s" % DEST if ok else "reference not found")
