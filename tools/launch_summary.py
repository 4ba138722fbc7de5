"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, mean ms, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
d = collections.defaultdict(list)
for r in rows[start + 1:]:
    if len(r) < len(h):
        continue
    x = dict(zip(h, r))
    if x["Metric Name"] == "gpu__time_duration.sum":
        v, u = float(x["Metric Value"]), x["Metric Unit"]
        d[x["Kernel Name"][:70]].append(v/1e6 if u.startswith("n") else v/1e3 if u.startswith("u") else v)
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} n={len(v):3d} mean={sum(v)/len(v):8.4f} ms share={sum(v)/tot*100:5.1f}%")
