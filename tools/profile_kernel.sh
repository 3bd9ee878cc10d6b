#!/bin/bash
# usage: tools/profile_kernel.sh <kernel-regex> <tag> [skip]   (runs on the GPU box via gpurun)
# Full ncu capture of one launch of the matching kernel during a 1024^3 bench step; exports the
# raw-page CSV (small) and keeps the .ncu-rep only if it is small enough to travel back.
set -u
PAT=$1; TAG=$2; SKIP=${3:-3}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$PAT -s $SKIP -c 1 -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page details --csv > gpurun_out/${TAG}_details.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_source.csv.gz
ls -la /tmp/prof_$TAG.ncu-rep gpurun_out/
SZ=$(stat -c %s /tmp/prof_$TAG.ncu-rep)
if [ "$SZ" -lt 30000000 ]; then cp /tmp/prof_$TAG.ncu-rep gpurun_out/; fi
