#!/usr/bin/env python
"""Turn an .ncu-rep brought back in gpurun_out/ into the small, tracked artefacts under profiles/.

  python profiles/summarize.py gpurun_out/prof_X.ncu-rep profiles/rNN_ncu_X.txt "title" [workload nodes]

Writes the key-metric summary (text) and, when `workload nodes` are given, records the kernel's
DRAM traffic per launch in profiles/traffic.json (bench.py reads it for roofline.traffic).
Runs in the build container (ncu -i needs no GPU).
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    # a report, or its `ncu -i X --page raw --csv` export (tools/ncu_capture.sh keeps only the export of
    # the secondary captures: gpurun_out/ is capped at 64 MiB)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(
        ["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: {Path(rep).name} (ncu --set full --clock-control none --import-source on, first captured launch)\n")
        for k in KEYS:
            if k in d:
                f.write(f"{k:90s} {d[k][0]:>24s} {d[k][1]}\n")
    if len(sys.argv) >= 6:
        workload, nodes = sys.argv[4], int(sys.argv[5])
        tj = Path(__file__).resolve().parent / "traffic.json"
        t = json.loads(tj.read_text()) if tj.exists() else {}
        b = sum(float(d[k][0]) * UNIT_SCALE[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t[workload] = {"dram_bytes_per_launch": b, "nodes_per_launch": nodes, "report": Path(rep).name,
                       "kernel": d["Kernel Name"][0]}
        tj.write_text(json.dumps(t, indent=1) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
