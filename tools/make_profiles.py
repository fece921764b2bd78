"""Regenerate the profiling evidence of the shipping kernels on a B200 (run under gpurun; one call):

    python tools/make_profiles.py [tag]        # tag defaults to r2

1. the plain benchmark command must exit 0;
2. ncu launch list of the same command (gpu__time_duration.sum per launch)   -> profiles/<tag>_launches_bench_c4.csv
3. ncu --set full of one launch of each stage kernel of a timed step         -> gpurun_out/<tag>_full_c4.ncu-rep
4. condensed metric table                                                    -> profiles/<tag>_ncu_full_c4.csv
5. DRAM bytes per launch of every captured kernel                            -> profiles/<tag>_traffic.json
   (bench.py reads this file for roofline.traffic: no constants in the benchmark)
6. SASS evidence of the tcgen05 / TMA instructions in the built library      -> profiles/<tag>_sass_tcgen05.txt
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
OUT = "gpurun_out"
os.makedirs(OUT, exist_ok=True)
# only gpurun_out/ comes back from the GPU box: write there, then `cp gpurun_out/profiles/* profiles/` here
PROF = os.path.join(OUT, "profiles")
os.makedirs(PROF, exist_ok=True)
CMD = [sys.executable, "bench.py", "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-alt", "--no-sampler", "--no-verify"]
KERNELS = "k_mlp_i8|k_fused_p1d_lanes|k_p1d_split|k_chi2_i8|k_chi2_dmma|k_stage1"


def run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    return subprocess.run(cmd, **kw)


def kernel_name(s):
    """'void <unnamed>::k_x<6, 0>(DevCtx, ...)' -> 'k_x'"""
    s = re.sub(r"\(.*", "", s).replace("void ", "").replace("<unnamed>::", "")
    return re.sub(r"<.*", "", s).split("::")[-1].strip()


def main():
    with open(os.path.join(OUT, TAG + "_plain.log"), "w") as f:
        r = run(CMD, stdout=f, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        print("plain run failed: not profiling")
        return 1
    line = json.loads([l for l in open(os.path.join(OUT, TAG + "_plain.log")) if l.startswith("{")][-1])
    launches = os.path.join(PROF, TAG + "_launches_bench_c4.csv")
    run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-c", "800", "--csv", "--log-file", launches] + CMD,
        stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    rep = os.path.join(OUT, TAG + "_full_c4")
    # the first 16 matching launches belong to the set-up and the warm-up steps; the next four are one timed step
    run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:" + KERNELS, "-s", "16", "-c", "4", "-f", "-o", rep] + CMD,
        stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    raw = os.path.join(OUT, TAG + "_full_c4_raw.csv")
    with open(raw, "w") as f:
        run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], stdout=f)
    condensed = os.path.join(PROF, TAG + "_ncu_full_c4.csv")
    run([sys.executable, "tools/ncu_condense.py", raw, condensed])
    rows = list(csv.reader(open(raw)))
    hdr, data = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    traffic, times = {}, {}
    units = {h: rows[1][i] for h, i in col.items()}
    byte_scale = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

    def nbytes(r_, metric):  # ncu picks a unit per metric column
        return float(r_[col[metric]].replace(",", "")) * byte_scale[units[metric]]

    for r_ in data:
        name = kernel_name(r_[col["Kernel Name"]])
        traffic[name] = nbytes(r_, "dram__bytes_read.sum") + nbytes(r_, "dram__bytes_write.sum")
        times[name] = float(r_[col["gpu__time_duration.sum"]].replace(",", ""))
    with open(os.path.join(PROF, TAG + "_traffic.json"), "w") as f:
        json.dump({"workload": "c4", "walkers": line["config"]["walkers_per_gpu_per_step"], "emulator_precision": line["config"]["emulator_precision"],
                   "dram_bytes_per_launch": traffic, "ncu_duration": {"unit": units.get("gpu__time_duration.sum"), "per_kernel": times},
                   "source": "ncu --set full --clock-control none, one launch per kernel of a timed step of: " + " ".join(CMD[1:])}, f, indent=1)
    # share of each kernel in the launch list (cold-cache, serialised times: shares, not absolutes)
    tot, per = 0.0, {}
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    for r_ in rows[1:]:
        nm = kernel_name(r_[ik])
        v = float(r_[iv].replace(",", ""))
        per[nm] = per.get(nm, 0.0) + v
        tot += v
    with open(os.path.join(PROF, TAG + "_launch_shares.txt"), "w") as f:
        for nm, v in sorted(per.items(), key=lambda kv: -kv[1]):
            f.write("%-40s %12.0f  %5.1f %%\n" % (nm, v, 100.0 * v / tot))
    with open(os.path.join(PROF, TAG + "_sass_tcgen05.txt"), "w") as f:
        sass = subprocess.run(["cuobjdump", "-sass", "cup1d_b200/libcup1d_b200.so"], capture_output=True, text=True).stdout
        cur, counts = None, {}
        for l in sass.splitlines():
            m = re.search(r"Function : (\S+)", l)
            if m:
                cur = m.group(1)
            for mn in re.findall(r"\b(UTC[A-Z0-9]*MMA|UTCBAR|LDTM|STTM|UBLKCP|UTMALDG|DMMA)\b", l):
                counts.setdefault(cur, {}).setdefault(mn, 0)
                counts[cur][mn] += 1
        for fn, c in counts.items():
            f.write("%s: %s\n" % (fn, ", ".join("%s x%d" % kv for kv in sorted(c.items()))))
    print(open(os.path.join(PROF, TAG + "_traffic.json")).read())
    return 0


if __name__ == "__main__":
    sys.exit(main())
