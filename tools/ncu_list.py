"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel count, avg, total, share."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg.setdefault(r[ki][:100], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{len(v):4d} x {sum(v) / len(v) / 1e3:10.1f} us  total {sum(v) / 1e6:8.2f} ms {sum(v) / tot * 100:5.1f}%  {k}")
