"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` log: per kernel launches, total and mean time, share."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((r["Kernel Name"], us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
rows = rows[skip:skip + count]
agg = collections.OrderedDict()
for k, us in rows:
    k = re.sub(r"\(.*", "", k)[:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | us total | us / launch | share |\n|---|---|---|---|---|")
for k, (n, us) in agg.items():
    print("| %s | %d | %.1f | %.1f | %.1f%% |" % (k, n, us, us / n, 100 * us / tot))
print("| total | %d | %.1f | | |" % (len(rows), tot))
