T21_ZY_GROUP=${G:-8} timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --cache-control none \
   -k regex:'tile_kernel' -s 600 -c 40 --csv --log-file gpurun_out/zyg${G:-8}_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/zyg${G:-8}_ncu.log 2>&1; echo rc=$?
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/zyg${G:-8}_ncu.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; break
i0=rows.index(hdr)
kn=hdr.index('Kernel Name'); mn=hdr.index('Metric Name'); mv=hdr.index('Metric Value'); idc=hdr.index('ID')
cur={}
for r in rows[i0+1:]:
    if len(r)<=mv: continue
    cur.setdefault(r[idc],{'k':r[kn][:40]})[r[mn]]=r[mv]
for k,v in list(cur.items())[:10]: print(k,v)
PY
