#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

path = sys.argv[1]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
last = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    i = int(row["ID"])
    if i < first or i >= last:
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
    k = row["Kernel Name"][:90]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
    n += 1
tot = sum(a[1] for a in agg.values())
print("launches %d, total %.3f ms" % (n, tot))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 25]:
    print("%-92s n=%5d %9.3f ms %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
