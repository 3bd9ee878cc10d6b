#!/bin/bash
# round 2, GPU call 1: packed-f32x2 build vs scalar build (tests + bench), and the staging / issue-rate microbenchmarks
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest_packed.log 2>&1; echo "pytest packed rc=$?"; tail -3 gpurun_out/c1_pytest_packed.log
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/c1_bench_packed.json 2> gpurun_out/c1_bench_packed.err; echo "bench packed rc=$?"
T21PS_LIB=$PWD/build_alt/libt21ps_scalar.so timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/c1_bench_scalar.json 2> gpurun_out/c1_bench_scalar.err; echo "bench scalar rc=$?"
timeout 300 tools/ubench/fp_rate > gpurun_out/c1_fp_rate.txt 2>&1; echo "fp_rate rc=$?"
timeout 300 tools/ubench/tma_tile > gpurun_out/c1_tma_tile.txt 2>&1; echo "tma_tile rc=$?"
timeout 600 python tools/bench_variants.py > gpurun_out/c1_variants_packed.json 2> gpurun_out/c1_variants_packed.err; echo "variants rc=$?"
cat gpurun_out/c1_fp_rate.txt gpurun_out/c1_tma_tile.txt
python - <<'PY'
import json
for t in ("packed","scalar"):
    try:
        d=json.loads(open("gpurun_out/c1_bench_%s.json"%t).read().strip().splitlines()[-1])
        print(t, d["value"], d["ms_per_step"], {k:round(v.get("ms",0),3) for k,v in d["passes"].items()})
    except Exception as e: print(t, "ERR", e)
PY
