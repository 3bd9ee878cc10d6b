for G in 0 4 8 16 32 64; do T21_ZY_GROUP=$G timeout 300 python bench.py --no-cpu --no-e2e > gpurun_out/zyg_$G.json 2> gpurun_out/zyg_$G.err; python -c "
import json
d=json.loads(open('gpurun_out/zyg_$G.json').read().strip().splitlines()[-1])
print('G=$G', round(d['ms_per_step'],3), {k:round(v.get('ms',0),3) for k,v in d['passes'].items()})
"; done
