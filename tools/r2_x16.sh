timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nx512" 2>&1 | tail -5 || exit 1
T21_PENCIL16=1 timeout 120 python tools/bench_variants.py --quick --size 512 > gpurun_out/x16_1.json 2> gpurun_out/x16_1.err; cat gpurun_out/x16_1.json; tail -2 gpurun_out/x16_1.err
T21_PENCIL16=1 timeout 120 python bench.py --size 512 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], {k:round(v.get('ms',0),4) for k,v in d['passes'].items()})"
