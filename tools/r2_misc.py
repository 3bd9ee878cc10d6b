"""Per-pass times (library CUDA events) of a few configurations outside the bench line: fp64 at 1024^3, fp32 at 2048^3."""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tools21cm_b200 as t2c
from tools21cm_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
def passes(fn, reps=3):
    for _ in range(2): fn()
    lib.t21_profile(1); acc = [0.0] * 5; buf = (C.c_float * 5)()
    for _ in range(reps):
        fn(); lib.t21_profile_last(buf, 5)
        acc = [a + b for a, b in zip(acc, buf)]
    lib.t21_profile(0)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return {"ms_per_call": e0.elapsed_time(e1) / reps, "mean_z_y_x_reduce_ms": [round(a / reps, 3) for a in acc]}
out = {}
g = torch.Generator(device=dev).manual_seed(3)
a = torch.randn((1024,) * 3, generator=g, device=dev, dtype=torch.float32).double()
out["power_spectrum_1d_1024_f64"] = passes(lambda: t2c.power_spectrum_1d(a, kbins=15, box_dims=714.0))
if "--f64-only" in sys.argv:
    print(json.dumps(out)); sys.exit(0)
del a; torch.cuda.empty_cache()
a = torch.randn((2048,) * 3, generator=g, device=dev, dtype=torch.float32)
out["power_spectrum_1d_2048_f32"] = passes(lambda: t2c.power_spectrum_1d(a, kbins=15, box_dims=714.0), reps=2)
print(json.dumps(out))
