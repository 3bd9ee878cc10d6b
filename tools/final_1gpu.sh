#!/bin/bash
# Round-end evidence on one B200: GPU tests, smoke, the bench line, the reference arm, the ncu launch list
# of the same bench command, and one full capture per pass.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; echo "bench rc=$?"
if [ "${1:-}" = "all" ]; then
  timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"
  timeout 600 python tools/bench_variants.py > gpurun_out/final_variants.json 2> gpurun_out/final_variants.err
fi
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k regex:'tile_kernel|xpass_kernel|mean_estimate|reduce_partials' -c 60 --csv --log-file gpurun_out/final_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/final_ncu_ll.log 2>&1; echo "ncu ll rc=$?"
bash tools/profile_kernel.sh ${2:-xpass_kernel} ${3:-xfin} ${4:-3} > gpurun_out/prof_last.log 2>&1
