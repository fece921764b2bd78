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
os.makedirs("profiles", exist_ok=True)
CMD = [sys.executable, "bench.py", "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-alt", "--no-sampler", "--no-verify"]
KERNELS = "k_mlp_i8|k_fused_p1d_lanes|k_chi2_dmma|k_stage1"


def run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    return subprocess.run(cmd, **kw)


def main():
    with open(os.path.join(OUT, TAG + "_plain.log"), "w") as f:
        r = run(CMD, stdout=f, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        print("plain run failed: not profiling")
        return 1
    line = json.loads([l for l in open(os.path.join(OUT, TAG + "_plain.log")) if l.startswith("{")][-1])
    launches = os.path.join("profiles", TAG + "_launches_bench_c4.csv")
    run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-c", "800", "--csv", "--log-file", launches] + CMD,
        stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    rep = os.path.join(OUT, TAG + "_full_c4")
    # the first 16 matching launches belong to the set-up and the warm-up steps; the next four are one timed step
    run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:" + KERNELS, "-s", "16", "-c", "4", "-f", "-o", rep] + CMD,
        stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    raw = os.path.join(OUT, TAG + "_full_c4_raw.csv")
    with open(raw, "w") as f:
        run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], stdout=f)
    condensed = os.path.join("profiles", TAG + "_ncu_full_c4.csv")
    run([sys.executable, "tools/ncu_condense.py", raw, condensed])
    rows = list(csv.reader(open(raw)))
    hdr, data = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    traffic, times = {}, {}
    for r_ in data:
        name = re.sub(r"\(.*", "", r_[col["Kernel Name"]]).replace("void ", "")
        name = re.sub(r"<.*", "", name).split("::")[-1]
        traffic[name] = float(r_[col["dram__bytes_read.sum"]].replace(",", "")) + float(r_[col["dram__bytes_write.sum"]].replace(",", ""))
        times[name] = float(r_[col["gpu__time_duration.sum"]].replace(",", ""))
    units = {h: rows[1][i] for h, i in col.items()}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
    traffic = {k: v * scale for k, v in traffic.items()}
    with open(os.path.join("profiles", TAG + "_traffic.json"), "w") as f:
        json.dump({"workload": "c4", "walkers": line["config"]["walkers_per_gpu_per_step"], "emulator_precision": line["config"]["emulator_precision"],
                   "dram_bytes_per_launch": traffic, "ncu_duration": {"unit": units.get("gpu__time_duration.sum"), "per_kernel": times},
                   "source": "ncu --set full --clock-control none, one launch per kernel of a timed step of: " + " ".join(CMD[1:])}, f, indent=1)
    # share of each kernel in the launch list (cold-cache, serialised times: shares, not absolutes)
    tot, per = 0.0, {}
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    for r_ in rows[1:]:
        nm = re.sub(r"\(.*", "", r_[ik]).replace("void ", "")
        nm = re.sub(r"<.*", "", nm).split("::")[-1]
        v = float(r_[iv].replace(",", ""))
        per[nm] = per.get(nm, 0.0) + v
        tot += v
    with open(os.path.join("profiles", TAG + "_launch_shares.txt"), "w") as f:
        for nm, v in sorted(per.items(), key=lambda kv: -kv[1]):
            f.write("%-40s %12.0f  %5.1f %%\n" % (nm, v, 100.0 * v / tot))
    with open(os.path.join("profiles", TAG + "_sass_tcgen05.txt"), "w") as f:
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
    print(open(os.path.join("profiles", TAG + "_traffic.json")).read())
    return 0


if __name__ == "__main__":
    sys.exit(main())
