"""Condense `ncu -i X.ncu-rep --page raw --csv` output into a small metric-by-kernel table.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_condense.py /tmp/raw.csv profiles/rN_ncu_<what>.csv
"""
import csv
import re
import sys

KEEP = [
    r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^lts__t_bytes\.sum$", r"^launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|cluster)",
    r"^sm__throughput\.avg\.pct", r"^sm__issue_active\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__warps_active\.avg\.pct",
    r"^sm__pipe_(fp64|tensor)[a-z_0-9]*cycles_active\.avg\.pct_of_peak_sustained_active$", r"^sm__inst_executed_pipe_(alu|fma|fp64|lsu|xu|uniform|tensor[a-z_0-9]*)\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__inst_executed\.sum$", r"^smsp__pcsamp_warps_issue_stalled_[a-z_]+$", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"^l1tex__t_sector_hit_rate\.pct$", r"^lts__t_sector_hit_rate\.pct$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__cycles_elapsed\.max$", r"^smsp__warps_eligible\.avg\.per_cycle_active$",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    names = [re.sub(r"\(.*", "", r[name_col]).replace("<unnamed>::", "").replace("void ", "") for r in data]
    with open(sys.argv[2], "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit"] + names)
        for i, h in enumerate(hdr):
            if any(re.search(p, h) for p in KEEP) and "not_issued" not in h:
                w.writerow([h, units[i]] + [r[i] for r in data])


if __name__ == "__main__":
    main()
