"""Where the host side of a call goes: cProfile of 3000 power_spectrum_1d calls on a 64^3 device tensor."""
import sys, os, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tools21cm_b200 as t2c
from tools21cm_b200 import synthetic
a = synthetic.toy_dt_cube_torch((64,) * 3, seed=1, device=torch.device("cuda", 0))
for _ in range(20): t2c.power_spectrum_1d(a, kbins=15, box_dims=200.0, return_n_modes=True)
pr = cProfile.Profile(); pr.enable()
for _ in range(3000): t2c.power_spectrum_1d(a, kbins=15, box_dims=200.0, return_n_modes=True)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:4500])
