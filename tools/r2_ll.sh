#!/bin/bash
# usage: tools/r2_ll.sh <variant> <tag>: per-launch durations of one estimator variant (ncu launch list)
set -u
V=$1; TAG=$2
CMD="python tools/run_variant.py $V 2"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'tile_kernel|xpass|rows_kernel|reduce_|mean_est' --csv --log-file gpurun_out/ll_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ll_$TAG.csv")) if len(r)>10]
hdr=rows[0]; iK=hdr.index("Kernel Name"); iM=hdr.index("Metric Name"); iV=hdr.index("Metric Value"); iI=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault(r[iI],{"k":r[iK]})[r[iM]]=r[iV]
for i,v in d.items():
    if "t21" in v["k"] or "kernel" in v["k"]:
        print(i, v["k"][:90], v.get("gpu__time_duration.sum"), v.get("dram__bytes_read.sum"), v.get("dram__bytes_write.sum"))
PY
