"""How fast does cudaHostRegister pin an existing numpy array on this host? (decides whether chunk-wise registration can
replace the staging copy of t21_h2d_staged)"""
import os, sys, time, ctypes as C
import numpy as np, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]; rt.cudaHostUnregister.argtypes = [C.c_void_p]
n = 1024 ** 3
a = np.ones(n, dtype=np.float32)
base = a.ctypes.data
for chunk_mb in (4096, 256, 64):
    ch = chunk_mb << 20
    t0 = time.perf_counter(); rc = 0
    for off in range(0, a.nbytes, ch):
        rc |= rt.cudaHostRegister(base + off, min(ch, a.nbytes - off), 0)
    t1 = time.perf_counter()
    for off in range(0, a.nbytes, ch):
        rc |= rt.cudaHostUnregister(base + off)
    t2 = time.perf_counter()
    print("chunk %4d MB: register %.1f GB/s, unregister %.1f GB/s, rc %d" % (chunk_mb, a.nbytes / (t1 - t0) / 1e9, a.nbytes / (t2 - t1) / 1e9, rc))
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
