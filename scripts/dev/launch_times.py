"""Summarise an ncu launch list (gpu__time_duration.sum per launch) by kernel name: count, mean, last value."""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"].split("(")[0][-60:], float(r["Metric Value"].replace(",", "")) / 1e3))
agg = collections.OrderedDict()
for k, v in rows:
    agg.setdefault(k, []).append(v)
for k, v in agg.items():
    print(f"{k:62s} n={len(v):3d} mean {sum(v)/len(v):9.1f} us  last {v[-1]:9.1f} us")
