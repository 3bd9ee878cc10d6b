"""Sweep of the staged host->device copy (pageable numpy -> device): chunk size is read once per process
(T21_H2D_CHUNK_MB), threads per call.  usage: T21_H2D_CHUNK_MB=64 python tools/r2_h2d.py"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools21cm_b200 import _lib
lib = _lib.load()
n = 1024 ** 3
a = np.ones(n, dtype=np.float32)
t = torch.empty(n, dtype=torch.float32, device="cuda")
p = torch.empty(n, dtype=torch.float32).pin_memory()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(fn, reps=4):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
out = {"chunk_mb": os.environ.get("T21_H2D_CHUNK_MB", "32"), "nt": os.environ.get("T21_H2D_NT", "1")}
out["pinned"] = a.nbytes / run(lambda: t.copy_(p, non_blocking=True)) / 1e9
for th in (8, 12, 16):
    out["threads_%d" % th] = a.nbytes / run(lambda: lib.t21_h2d_staged(t.data_ptr(), a.ctypes.data, a.nbytes, th, st)) / 1e9
print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in out.items()}, "GB/s; cores", os.cpu_count())
