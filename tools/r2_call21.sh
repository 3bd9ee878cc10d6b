#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "mixed_and_odd or golden or c2ray_grid or low_dimensional" > gpurun_out/c21_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c21_pytest.log
python - <<'PY'
import torch, time, numpy as np
import tools21cm_b200 as t2c
from tools21cm_b200 import synthetic
for n in (250, 300, 600):
    a = synthetic.toy_dt_cube_torch((n,)*3, seed=1, device="cuda")
    def timed(fn, k=10):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/k
    ms = timed(lambda: t2c.power_spectrum_1d(a, kbins=15, box_dims=244/0.7))
    b = 4*n**3 + 4*8*n*n*(n//2+1)
    print("power_spectrum_1d %d^3 fp32: %.3f ms per call, %.0f GB/s of algorithmic bytes (%.2f of 6557.8)" % (n, ms, b/ms/1e6, b/ms/1e6/6557.8))
PY
