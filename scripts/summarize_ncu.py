"""Summarise an ncu report (.ncu-rep) into the few numbers profiles/ keeps: duration, DRAM bytes,
throughputs, occupancy, stall reasons.  Usage: python scripts/summarize_ncu.py report.ncu-rep [launch_index]"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active"]
out = {}
for k in keys:
    if k in d:
        out[k] = f"{d[k][1]} {d[k][0]}".strip()
stalls = {}
for k, (u, v) in d.items():
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        try:
            f = float(v)
        except ValueError:
            continue
        if f >= 0.05:
            stalls[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(f, 3)
out["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda x: -x[1]))
print(json.dumps(out, indent=1))
