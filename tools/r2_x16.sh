timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nx512" 2>&1 | tail -15
for P in 1 0; do T21_PENCIL16=$P timeout 300 python tools/bench_variants.py --quick --size 512 > gpurun_out/x16_$P.json 2> gpurun_out/x16_$P.err; echo "PENCIL16=$P"; cat gpurun_out/x16_$P.json; tail -2 gpurun_out/x16_$P.err; done
T21_PENCIL16=1 timeout 300 python bench.py --size 512 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], {k:round(v.get('ms',0),4) for k,v in d['passes'].items()})"
