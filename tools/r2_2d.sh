timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "2d or golden" 2>&1 | tail -15
timeout 600 python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import tools21cm_b200 as t2c
from tools21cm_b200 import synthetic
a = synthetic.toy_dt_cube_torch((1024,)*3, seed=1, device=torch.device('cuda', 0))
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
out = {}
for nu in (0, 1, 2):
    out["power_spectrum_2d_nu%d" % nu] = timed(lambda: t2c.power_spectrum_2d(a, kbins=10, box_dims=714.0, nu_axis=nu))
print(json.dumps(out))
PY
