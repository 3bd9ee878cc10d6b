timeout 900 python -m pytest tests -m gpu -q -x -k "nd or golden or long_pencils or parseval or properties" 2>&1 | tail -6
timeout 600 python - <<'PY'
import sys, json, time, torch
sys.path.insert(0, '.')
import tools21cm_b200 as t2c
from tools21cm_b200 import synthetic, _lib
import ctypes as C
a = synthetic.toy_dt_cube_torch((1024,)*3, seed=1, device=torch.device('cuda', 0))
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
out = {"power_spectrum_nd": timed(lambda: t2c.power_spectrum_nd(a, box_dims=714.0))}
lib = _lib.load(); lib.t21_profile(1); t2c.power_spectrum_nd(a, box_dims=714.0); buf = (C.c_float * 5)(); lib.t21_profile_last(buf, 5); lib.t21_profile(0)
out["passes_ms"] = [round(v, 3) for v in buf]
# exact check against the other route at full size
import os
ref = t2c.power_spectrum_nd(a, box_dims=714.0)
os.environ["T21_PENCIL32_ND"] = "0"
old = t2c.power_spectrum_nd(a, box_dims=714.0)
out["max_abs_diff_vs_cta_per_tile_route"] = float((ref - old).abs().max()); out["max"] = float(old.max())
out["equal_fraction"] = float((ref == old).float().mean())
print(json.dumps(out))
PY
