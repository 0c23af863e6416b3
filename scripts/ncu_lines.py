"""Per-source-line instruction and stall-sample totals of one kernel from an ncu report captured with
--set full --import-source on (binary built with -lineinfo).  Usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
lines = []
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        d = dict(zip(hdr, r))
        lines.append((int(r[0]), r[1], int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0),
                      int(d["Warp Stall Sampling (Not-issued Samples)"] or 0)))
ti = sum(x[2] for x in lines) or 1
ts = sum(x[3] for x in lines) or 1
print(f"total warp instructions {ti}, samples {ts}")
print("-- by instructions")
for ln, src, ins, smp, ni in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f"{ln:5d} {100 * ins / ti:5.1f}% inst {100 * smp / ts:5.1f}% smp  {src.strip()[:110]}")
print("-- by stall samples")
for ln, src, ins, smp, ni in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{ln:5d} {100 * ins / ti:5.1f}% inst {100 * smp / ts:5.1f}% smp  {src.strip()[:110]}")
