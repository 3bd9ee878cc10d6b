#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_inputs.py -m gpu -x -q > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c18_pytest.log
timeout 900 python bench.py --no-cpu > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c18_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c18_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"])
PY
nproc; free -g | head -2
