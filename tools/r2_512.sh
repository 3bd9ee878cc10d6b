for n in 256 512; do timeout 300 python bench.py --size $n --no-cpu --no-e2e > gpurun_out/r02_bench_$n.json 2> gpurun_out/r02_bench_$n.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_$n.json').read().strip().splitlines()[-1])
print($n, d['ms_per_step'], {k:round(v.get('ms',0),4) for k,v in d['passes'].items()})
"; done
