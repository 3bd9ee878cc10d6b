#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_pencils or warp_per_pencil or golden" > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c7_pytest.log
timeout 600 python tools/bench_variants.py --quick > gpurun_out/c7_variants.json 2> gpurun_out/c7_variants.err; echo "variants rc=$?"; cat gpurun_out/c7_variants.json
