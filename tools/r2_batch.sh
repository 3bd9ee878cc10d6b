for B in 8 4 16; do L=""; [ $B != 8 ] && L="T21PS_LIB=$PWD/build_alt/b$B/libt21ps.so"; env $L timeout 120 python bench.py --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch $B', round(d['ms_per_step'],4), round(d['passes']['x']['ms'],4))"; done
