#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: instruction mix, hottest regions, stall samples."""
import csv, gzip, io, sys, collections
path = sys.argv[1]
txt = gzip.open(path, 'rt').read() if path.endswith('.gz') else open(path).read()
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
iS, iE, iSt = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
tot = 0; stall_tot = 0
byop = collections.Counter(); stall_by = collections.Counter()
body = []
for r in rows[2:]:
    try: n = int(float(r[iE])); st = int(float(r[iSt]))
    except Exception: continue
    toks = r[iS].split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
    base = op.split('.')[0]
    byop[base] += n; stall_by[base] += st; tot += n; stall_tot += st
    body.append((n, st, r[iS].strip()))
print('static instrs', len(body), 'executed warp-instrs', tot, 'stall samples', stall_tot)
print('%-12s %14s %6s %8s' % ('op', 'executed', '%', 'stall%'))
for k, v in byop.most_common(28):
    print('%-12s %14d %6.1f %8.1f' % (k, v, 100.0 * v / tot, 100.0 * stall_by[k] / max(1, stall_tot)))
if len(sys.argv) > 2:
    print('--- top stalled instructions')
    for n, st, s in sorted(body, key=lambda t: -t[1])[:int(sys.argv[2])]:
        print('%10d %7d  %s' % (n, st, s))
