"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(r[ix["Kernel Name"]].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `{}` | {} | {:.3f} | {:.1f}% |".format(k, a[0], a[1] / 1e6, 100 * a[1] / tot))
