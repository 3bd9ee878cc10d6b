#!/bin/bash
# full ncu captures of the Z and Y passes (both are instances of tile_kernel: picked by launch order -- mean, Z, Y, X, reduce per step)
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_zy.log 2>&1 || { echo "plain run failed"; exit 1; }
for K in "6 zf" "7 yf"; do
  set -- $K
  ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s $1 -c 1 -o /tmp/prof_$2 $CMD > gpurun_out/ncu_$2.log 2>&1
  ncu -i /tmp/prof_$2.ncu-rep --page raw --csv > gpurun_out/r02_$2_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$2.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_$2_source.csv.gz
  (python tools/ncu_summary.py gpurun_out/r02_$2_raw.csv; python tools/sass_mix.py gpurun_out/r02_$2_source.csv.gz 12) > gpurun_out/r02_ncu_final_$2_summary.txt 2>&1
  head -8 gpurun_out/r02_ncu_final_$2_summary.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or long_pencils or mixed_and_odd or large_cube or low_dim or torch_cuda" > gpurun_out/nd_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/nd_pytest.log
python tools/bench_variants.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nd', d['power_spectrum_nd'], '2d', d['power_spectrum_2d'])"
