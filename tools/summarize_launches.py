"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name (and grid), shares."""
import csv
import collections
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r["Kernel Name"]
        name = name.split("(")[0][-70:]
        rows.append((name, r.get("Grid Size", ""), ns))
tot = sum(r[2] for r in rows)
agg = collections.OrderedDict()
for n, g, ns in rows:
    k = (n, g)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ns
print("total %.1f us over %d launches" % (tot / 1e3, len(rows)))
for (n, g), (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%6.2f%%  %9.1f us  x%-4d %8.1f us/launch  grid %-14s %s" % (100 * ns / tot, ns / 1e3, c, ns / 1e3 / c, g, n))
