#!/bin/bash
# usage: tools/r2_prof.sh <kernel-regex> <tag> [skip] [extra env]   -- one full ncu capture of a kernel inside a short bench run
set -u
PAT=$1; TAG=$2; SKIP=${3:-3}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$PAT -s $SKIP -c 1 -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_source.csv.gz
python tools/ncu_summary.py gpurun_out/${TAG}_raw.csv
python tools/sass_mix.py gpurun_out/${TAG}_source.csv.gz 25
