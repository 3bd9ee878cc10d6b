cat > /tmp/nd_once.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import tools21cm_b200 as t2c
from tools21cm_b200 import synthetic
a = synthetic.toy_dt_cube_torch((1024,)*3, seed=1, device=torch.device('cuda', 0))
for _ in range(3): t2c.power_spectrum_nd(a, box_dims=714.0)
torch.cuda.synchronize()
PY
for V in 1 0; do T21_PENCIL32_ND=$V ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/nd_ll_$V.csv python /tmp/nd_once.py > /dev/null 2>&1; echo "PENCIL32_ND=$V"; python tools/launch_list.py gpurun_out/nd_ll_$V.csv 2>&1 | head -8; done
