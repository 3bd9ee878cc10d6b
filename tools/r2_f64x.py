"""fp64 cross spectra (two fields) at 512^3 and 1024^3: ms per call."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tools21cm_b200 as t2c
dev = torch.device("cuda", 0)
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
out = {}
g = torch.Generator(device=dev).manual_seed(3)
for n in (512, 1024):
    a = torch.randn((n,) * 3, generator=g, device=dev, dtype=torch.float32).double()
    b = torch.randn((n,) * 3, generator=g, device=dev, dtype=torch.float32).double()
    out["cross_1d_f64_%d" % n] = timed(lambda: t2c.cross_power_spectrum_1d(a, b, kbins=15, box_dims=714.0))
    out["cross_mu_f64_%d" % n] = timed(lambda: t2c.cross_power_spectrum_mu(a, b, kbins=10, mubins=5, box_dims=714.0))
    del a, b; torch.cuda.empty_cache()
print(json.dumps(out))
