#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into the handful of numbers DESIGN.md / bench.py cite.
usage: ncu_summary.py report.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_inst",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "launch__occupancy_limit_registers": "occ_limit_regs_blocks",
    "launch__waves_per_multiprocessor": "waves_per_sm",
}
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for i, h in enumerate(hdr):
        if h in WANT:
            d[WANT[h]] = "%s %s" % (r[i], units[i])
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                if float(r[i]) >= 0.3:
                    d.setdefault("stalls_per_issue", {})[h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = float(r[i])
            except ValueError:
                pass
    out.append(d)
txt = json.dumps(out, indent=1)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
print(txt)
