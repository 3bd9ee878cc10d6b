#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch
list: per kernel the number of launches, its share of the summed kernel time, the average duration and the
DRAM bytes per launch."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0] != "ID"]
per = collections.defaultdict(lambda: collections.defaultdict(dict))
for r in rows:
    per[r[4]][r[0]][r[12]] = (float(r[14].replace(",", "")), r[13])
unit = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
bunit = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
stats = []
for k, launches in per.items():
    us = [v["gpu__time_duration.sum"][0] * unit[v["gpu__time_duration.sum"][1]] for v in launches.values()]
    rd = [v["dram__bytes_read.sum"][0] * bunit[v["dram__bytes_read.sum"][1]] for v in launches.values() if "dram__bytes_read.sum" in v]
    wr = [v["dram__bytes_write.sum"][0] * bunit[v["dram__bytes_write.sum"][1]] for v in launches.values() if "dram__bytes_write.sum" in v]
    stats.append((sum(us), k, len(us), sum(us) / len(us), sum(rd) / max(1, len(rd)), sum(wr) / max(1, len(wr))))
tot = sum(s[0] for s in stats)
for s in sorted(stats, reverse=True):
    print("%-62.62s launches %3d  share of step %5.1f%%  avg %8.1f us  dram R / W per launch %.3f / %.3f GB" %
          (s[1].replace("void ", "").replace("t21::", ""), s[2], 100 * s[0] / tot, s[3], s[4], s[5]))
