#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/nd_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/nd_pytest.log
