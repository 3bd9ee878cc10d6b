#!/bin/bash
# Round-2 evidence on one B200: GPU tests, smoke, the bench line, the reference arm, the variants table, the ncu launch
# list of the bench command and one full capture of each pass.  Everything lands in gpurun_out/ (copied to profiles/).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs --durations=6 > gpurun_out/r02_pytest_gpu_1gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_gpu_1gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_1gpu_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 600 python tools/bench_variants.py > gpurun_out/r02_variants_1gpu.json 2> gpurun_out/r02_variants.err; echo "variants rc=$?"
timeout 300 python tools/bench_variants.py --size 256 --quick > gpurun_out/r02_variants_256.json 2>/dev/null
timeout 300 python tools/bench_variants.py --size 512 --quick > gpurun_out/r02_variants_512.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k regex:'tile_kernel|xpass|mean_estimate|reduce_partials' -c 60 --csv --log-file gpurun_out/r02_ncu_launches_bench_1024.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_ncu_ll.log 2>&1; echo "ncu ll rc=$?"
python tools/launch_list.py gpurun_out/r02_ncu_launches_bench_1024.csv > gpurun_out/r02_ncu_launch_list_bench_1024.txt 2>&1
for K in "xpass32_kernel x32" "tile_kernel.*ZPass zf" "tile_kernel.*YPass yf"; do
  set -- $K
  ncu --set full --clock-control none --import-source on -k regex:"$1" -s 3 -c 1 -o /tmp/prof_$2 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$2.log 2>&1
  ncu -i /tmp/prof_$2.ncu-rep --page raw --csv > gpurun_out/r02_$2_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$2.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_$2_source.csv.gz
  (python tools/ncu_summary.py gpurun_out/r02_$2_raw.csv; python tools/sass_mix.py gpurun_out/r02_$2_source.csv.gz 12) > gpurun_out/r02_ncu_final_$2_summary.txt 2>&1
done
head -30 gpurun_out/r02_ncu_final_x32_summary.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], {k:round(v.get("ms",0),3) for k,v in d["passes"].items()})
print("roofline", d["roofline"]); print("e2e", d["e2e"]); print("cpu", d["cpu_baseline"])
print(open("gpurun_out/r02_bench_1gpu_reference_arm.json").read()[:600])
print(open("gpurun_out/r02_variants_1gpu.json").read())
PY
bash tools/r2_np2.sh
bash tools/r2_512.sh
