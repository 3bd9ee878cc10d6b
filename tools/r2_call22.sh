#!/bin/bash
for n in 250 600; do
python bench.py --size $n --no-cpu --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print($n, d['ms_per_step'], {k:round(v.get('ms',0),3) for k,v in d['passes'].items()})"
done
