#!/usr/bin/env python
"""Group the instructions of an `ncu --page source --csv` dump by how often they ran per warp-tile, and list the
contiguous code regions that carry the work.  usage: ncu_regions.py <source.csv[.gz]> <warp_tiles> [min instr/wt]"""
import csv, gzip, io, collections, sys
path, WT = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 15
txt = gzip.open(path, 'rt').read() if path.endswith('.gz') else open(path).read()
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; iE = hdr.index('Instructions Executed'); iS = hdr.index('Source'); iSt = hdr.index('Warp Stall Sampling (All Samples)')
cls = collections.Counter(); n = collections.Counter(); st = collections.Counter(); seq = []
for r in rows[2:]:
    try: e = int(r[iE]); s = int(r[iSt])
    except Exception: continue
    seq.append((e, s, r[iS].strip()))
    key = round(e / WT, 2)
    cls[key] += e; n[key] += 1; st[key] += s
tot = sum(cls.values())
print("instructions per warp-tile: %.1f" % (tot / WT))
for k, v in sorted(cls.items(), key=lambda t: -t[1])[:20]:
    print("exec/warp-tile %7.2f : static %5d  instr/warp-tile %8.1f  stall samples %6d" % (k, n[k], v / WT, st[k]))
print("---- contiguous regions")
cur = None; start = 0; acc = 0; sacc = 0; out = []
for i, (e, s, src) in enumerate(seq):
    key = round(e / WT, 2)
    if key != cur:
        if cur is not None: out.append((start, i - 1, cur, acc, sacc))
        cur = key; start = i; acc = 0; sacc = 0
    acc += e; sacc += s
out.append((start, len(seq) - 1, cur, acc, sacc))
for a, b, k, acc, sacc in out:
    if acc / WT > thr: print("instr %5d-%5d exec/wt %6.2f  instr/wt %7.1f stall %6d   %s" % (a, b, k, acc / WT, sacc, seq[a][2][:60]))
