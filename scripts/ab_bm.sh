#!/bin/bash
# A/B of the bitmap AND kernel variants on the C2 batch (scratch; results under gpurun_out/)
M="smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem"
for v in "$@"; do
  XGM_BM_VARIANT=$v python scripts/quick_perf.py 10000000 1000000 4096 > gpurun_out/r2_perf_v$v.log 2>&1
  echo "== variant $v"; grep -E "replay" gpurun_out/r2_perf_v$v.log
  XGM_BM_VARIANT=$v ncu --metrics $M --clock-control none -k regex:xgm_and_bm -s 6 -c 1 --csv --log-file gpurun_out/r2_ncu_v$v.csv python scripts/quick_perf.py 10000000 1000000 4096 > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r2_ncu_v$v.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
if h:
    hd=rows[h[0]]
    for r in rows[h[0]+1:]:
        d=dict(zip(hd,r)); print(d.get("Metric Name"), d.get("Metric Value"))
PY
done
