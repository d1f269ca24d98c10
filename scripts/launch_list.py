"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list: per kernel count / mean / total / share."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    k = r[iK].split("(")[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[iV].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':32s} {'launches':>8s} {'mean_us':>10s} {'total_us':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:32s} {n:8d} {t / n / 1e3:10.2f} {t / 1e3:10.2f} {100 * t / tot:6.1f}%")
