#!/usr/bin/env python
"""Stall-reason breakdown of an instruction index range of an `ncu --page source --csv` dump.
usage: ncu_stalls.py <source.csv[.gz]> <first> <last> [top]"""
import csv, gzip, io, collections, sys
path, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 12
txt = gzip.open(path, 'rt').read() if path.endswith('.gz') else open(path).read()
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iSt = hdr.index('Warp Stall Sampling (All Samples)')
body = [r for r in rows[2:] if len(r) > iSt]
tot = collections.Counter(); per = []
for idx, r in enumerate(body):
    if idx < a or idx > b: continue
    for i in cols:
        try: tot[hdr[i]] += int(r[i])
        except ValueError: pass
    per.append((int(r[iSt] or 0), idx, r[iS].strip(), {hdr[i]: int(r[i] or 0) for i in cols if int(r[i] or 0) > 0}))
s = sum(tot.values()) or 1
print("range %d-%d: %d samples" % (a, b, s))
print(", ".join("%s %.0f%%" % (k.replace('stall_', ''), 100.0 * v / s) for k, v in tot.most_common(10)))
for st, idx, src, d in sorted(per, reverse=True)[:top]:
    print("%5d %6d  %-60s %s" % (idx, st, src[:60], ", ".join("%s:%d" % (k.replace('stall_', ''), v) for k, v in sorted(d.items(), key=lambda t: -t[1])[:3])))
