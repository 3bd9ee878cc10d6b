timeout 900 python -m pytest tests -m gpu -q -x -k "not 512_against_oracle" 2>&1 | tail -8
for S in 1 0; do T21_X32_SEGMENTS=$S timeout 300 python bench.py --no-cpu --no-e2e > gpurun_out/seg_$S.json 2> gpurun_out/seg_$S.err; python -c "
import json
d=json.loads(open('gpurun_out/seg_$S.json').read().strip().splitlines()[-1])
print('SEG=$S', round(d['ms_per_step'],3), {k:round(v.get('ms',0),3) for k,v in d['passes'].items()}, d['n_modes_total'])
"; tail -3 gpurun_out/seg_$S.err; done
timeout 600 python tools/bench_variants.py --quick > gpurun_out/seg_variants.json 2> gpurun_out/seg_variants.err; cat gpurun_out/seg_variants.json; tail -3 gpurun_out/seg_variants.err
