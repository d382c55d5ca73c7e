#!/usr/bin/env python
"""profiles/r02_dram_traffic.json (read by bench.py for `roofline.traffic`) and a per-kernel summary table from an
`ncu --set full` report of bench.py:

    ncu -i gpurun_out/r02g_prof.ncu-rep --page raw --csv > /tmp/prof_raw.csv
    python tools/ncu_traffic.py /tmp/prof_raw.csv gpurun_out/r02g_prof.ncu-rep
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = {"source": f"ncu --set full --clock-control none, {os.path.basename(sys.argv[2])} (bench.py 2^24 Ft63 workload, one launch each)",
           "kernels": {}}
    table = []
    for r in rows[2:]:
        name = re.match(r"(?:void )?(\w+)", r[idx["Kernel Name"]]).group(1)
        rd = float(r[idx["dram__bytes_read.sum"]]) * SCALE[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * SCALE[units[idx["dram__bytes_write.sum"]]]
        out["kernels"][name] = {"dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
                                "time_us_under_ncu": float(r[idx["gpu__time_duration.sum"]]),
                                "warp_instructions": int(float(r[idx["smsp__inst_executed.sum"]]))}
        table.append([name] + [r[idx[m]] for m in METRICS])
    with open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print("| kernel | " + " | ".join(m.split(".")[0].replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active", "")
                                    for m in METRICS) + " |")
    print("|" + "---|" * (len(METRICS) + 1))
    for t in table:
        print("| " + " | ".join(t) + " |")


if __name__ == "__main__":
    main()
