#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q -k "golden or integrated or abi" > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c17_pytest.log
python - <<'PY'
import torch, time, numpy as np
import tools21cm_b200 as t2c
from tools21cm_b200 import synthetic
a = synthetic.toy_dt_cube_torch((1024,)*3, seed=1, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0=time.perf_counter()
    r = t2c.integrated_bispectrum(a, Ncuts=4, kbins=15, box_dims=714.0, verbose=False)
    torch.cuda.synchronize(); print("integrated_bispectrum 1024^3 Ncuts=4 (64 sub-boxes): %.1f ms" % ((time.perf_counter()-t0)*1e3))
PY
