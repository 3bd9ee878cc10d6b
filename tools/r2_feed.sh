timeout 900 python -m pytest tests/test_gpu_inputs.py -m gpu -q -x 2>&1 | tail -6
for F in 1 0; do T21_HOST_FEED=$F timeout 600 python bench.py --no-cpu > gpurun_out/feed_$F.json 2> gpurun_out/feed_$F.err; python -c "
import json
d=json.loads(open('gpurun_out/feed_$F.json').read().strip().splitlines()[-1])
print('FEED=$F', round(d['ms_per_step'],3), d['e2e'])
"; tail -2 gpurun_out/feed_$F.err; done
