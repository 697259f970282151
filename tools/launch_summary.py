"""per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python tools/launch_summary.py FILE [regex-to-list-in-order]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    seq.append((name, v / 1e3 if u.startswith("n") else v if u.startswith("u") else v * 1e3))
agg = collections.OrderedDict()
for n, v in seq:
    agg.setdefault(n, []).append(v)
print("%d launches (times in us; cold-cache, serialised)" % len(seq))
for n, v in agg.items():
    print("%-64s n=%4d mean %9.1f min %9.1f max %9.1f" % (n[:64], len(v), sum(v) / len(v), min(v), max(v)))
if len(sys.argv) > 2:
    pat = re.compile(sys.argv[2])
    for n in agg:
        if pat.search(n):
            print(n, [round(x) for x in agg[n]])
