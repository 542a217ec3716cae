#!/usr/bin/env python
"""Turn the ncu outputs of scripts/profile.sh (gpurun_out/) into the small CSV summaries committed under profiles/.

    python scripts/summarise_profiles.py r2            # reads gpurun_out/r2_prof.ncu-rep, gpurun_out/r2_launches.csv

Runs without a GPU (ncu -i only reads the report)."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof.ncu-rep")
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.csv")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    name_col = head.index("Kernel Name")
    with open(out, "w") as f:
        f.write("kernel," + ",".join(f"{k} [{units[head.index(k)]}]" for k in KEYS if k in head) + ",dram_bytes_total [GB],dram_GBps\n")
        for r in data:
            vals = [r[head.index(k)] for k in KEYS if k in head]
            rd = float(r[head.index("dram__bytes_read.sum")]) * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[units[head.index("dram__bytes_read.sum")]]
            wr = float(r[head.index("dram__bytes_write.sum")]) * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[units[head.index("dram__bytes_write.sum")]]
            t = float(r[head.index("gpu__time_duration.sum")]) * {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}[units[head.index("gpu__time_duration.sum")]]
            f.write('"' + r[name_col].split("(")[0].replace("void ", "") + '",' + ",".join(vals) + f",{rd + wr:.3f},{(rd + wr) / t:.0f}\n")
    print(open(out).read())
    # launch list: keep kernel name + duration only
    src = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if os.path.exists(src):
        dst = os.path.join(ROOT, "profiles", f"{tag}_launches.csv")
        lines = [l for l in open(src).read().splitlines() if l.startswith('"')]
        rd = list(csv.reader(lines))
        h = rd[0]
        kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
        with open(dst, "w") as f:
            f.write("launch,kernel,gpu__time_duration.sum,unit\n")
            for k, r in enumerate(rd[1:]):
                f.write(f'{k},"{r[kn].split("(")[0].replace("void ", "")}",{r[mv]},{r[mu]}\n')
        print(f"{dst}: {len(rd) - 1} launches")


if __name__ == "__main__":
    main()
