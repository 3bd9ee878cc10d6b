#!/bin/bash
# usage: tools/r2_profv.sh <variant> <kernel-regex> <tag> [warp_tiles]
set -u
V=$1; PAT=$2; TAG=$3
CMD="python tools/run_variant.py $V 3"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$PAT -s 2 -c 1 -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_source.csv.gz
python tools/ncu_summary.py gpurun_out/${TAG}_raw.csv
python tools/sass_mix.py gpurun_out/${TAG}_source.csv.gz 12 | head -30
python tools/ncu_regions.py gpurun_out/${TAG}_source.csv.gz ${4:-532480} 20
