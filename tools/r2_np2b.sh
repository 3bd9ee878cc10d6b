for n in 250 300 600; do timeout 300 python bench.py --size $n --no-cpu --no-e2e > gpurun_out/r02_bench_np2_$n.json 2> gpurun_out/r02_bench_np2_$n.err; echo "rc=$? $n"; python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_bench_np2_$n.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:round(v.get('ms',0),4) for k,v in d['passes'].items()})
"; done
