#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q -k "golden or multipoles or abi" > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c10_pytest.log
python - <<'PY'
import torch, time, numpy as np
import tools21cm_b200 as t2c
from tools21cm_b200 import synthetic
a = synthetic.toy_dt_cube_torch((1024,)*3, seed=1, device="cuda")
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print("power_spectrum_2d ms", timed(lambda: t2c.power_spectrum_2d(a, kbins=10, box_dims=714.0)))
print("multipoles los0 ms", timed(lambda: t2c.power_spectrum_multipoles(a, kbins=15, box_dims=714.0, los_axis=0)))
print("multipoles los2 ms", timed(lambda: t2c.power_spectrum_multipoles(a, kbins=15, box_dims=714.0, los_axis=2)))
print("ps1d ms", timed(lambda: t2c.power_spectrum_1d(a, kbins=15, box_dims=714.0)))
PY
