#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_pencils" > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c2_pytest.log
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/c2_bench_x32.json 2> gpurun_out/c2_bench_x32.err; echo "bench x32 rc=$?"; tail -3 gpurun_out/c2_bench_x32.err
T21_PENCIL32=0 timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/c2_bench_old.json 2> gpurun_out/c2_bench_old.err; echo "bench old rc=$?"
python - <<'PY'
import json
for t in ("x32","old"):
    try:
        d=json.loads(open("gpurun_out/c2_bench_%s.json"%t).read().strip().splitlines()[-1])
        print(t, d["value"], d["ms_per_step"], {k:round(v.get("ms",0),3) for k,v in d["passes"].items()}, d["n_modes_total"])
    except Exception as e: print(t, "ERR", e)
PY
