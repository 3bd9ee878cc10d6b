"""Experiment: sensitivity of the pass times to the placement of the input cube and the workspace."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools21cm_b200 import _lib, binspec as bs, power_spectrum as ps, synthetic

n = 1024
lib = _lib.load()
dev = torch.device("cuda", 0)
x0 = synthetic.toy_dt_cube_torch((n, n, n), seed=1235, device=dev)
d = ps._device(0)
plan = d.plan((n, n, n), _lib.T21_F32)
spec, cspec, own = d.spec(((n, n, n), "zexp"), lambda: bs.build_binspec((n, n, n), bs.get_dims(714.0, (n, n, n)), 15))
nb = spec.nbins
need = lib.t21_workspace_bytes(plan, 1, nb)
xb = x0.numel() * 4
big = torch.empty(xb + need + (64 << 20), dtype=torch.uint8, device=dev)
base = big.data_ptr()
base_al = (base + (2 << 20) - 1) // (2 << 20) * (2 << 20)
sums = torch.empty(nb, dtype=torch.float64, device=dev)
counts = torch.empty(nb, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
buf = (C.c_float * 5)()

def run(xoff, woff, reps=6):
    xp = base_al + xoff
    wp = base_al + (32 << 20) + xb + woff
    xv = torch.as_tensor(ps_view(xp, x0.numel()), device=dev) if False else None
    # copy the cube into place
    C.cdll.LoadLibrary("libcudart.so").cudaMemcpy(C.c_void_p(xp), C.c_void_p(x0.data_ptr()), C.c_size_t(xb), 3)
    acc = np.zeros(5)
    lib.t21_profile(1)
    for i in range(reps):
        rc = lib.t21_ps_binned(plan, xp, None, 1, C.byref(cspec), sums.data_ptr(), counts.data_ptr(), wp, need, st)
        assert rc == 0, lib.t21_last_error()
        lib.t21_profile_last(buf, 5)
        if i >= 2:
            acc += np.array(list(buf))
    lib.t21_profile(0)
    acc /= (reps - 2)
    return acc

print("base %x aligned %x need %d" % (base, base_al, need))
for xoff, woff in [(0, 0), (0, 256), (0, 4096), (0, 65536), (0, 1 << 20), (256, 0), (4096, 0), (65536, 0), (1 << 20, 0),
                   (0, 512), (0, 1024), (0, 2048), (0, 8192), (0, 16384), (0, 32768), (0, 131072), (0, 262144), (0, 524288)]:
    a = run(xoff, woff)
    print("xoff %8d woff %8d  z %.3f y %.3f x %.3f" % (xoff, woff, a[1], a[2], a[3]), flush=True)
print(counts.sum().item())
