"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: total / count / mean per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        try:
            d[r[ki][:70]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v)/1e3:10.1f} us {100*sum(v)/tot:5.1f}% {len(v):5d} x {sum(v)/len(v)/1e3:8.1f} us  {k}")
