#!/usr/bin/env python
"""Key per-kernel metrics of an ncu report as JSON lines:
    python tools/ncu_summary.py report.ncu-rep"""
import csv
import json
import subprocess
import sys

WANT = {
    "Kernel Name": "kernel", "gpu__time_duration.sum": "time", "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "smsp__inst_executed.sum": "warp_inst",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__cycles_elapsed.max": "cycles", "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
}
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = {}
    for i, h in enumerate(hdr):
        if h in WANT:
            v = r[i]
            if h == "Kernel Name":
                v = v.split("(")[0]
            else:
                v = f"{v} {units[i]}".strip()
            d[WANT[h]] = v
    print(json.dumps(d))
