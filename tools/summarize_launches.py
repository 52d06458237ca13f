"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
tot = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): v *= 1e3
    elif unit in ("ms", "msecond"): v *= 1e6
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print("total %.3f ms over %d launches" % (tot / 1e6, sum(a[0] for a in agg.values())))
for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.3f ms %5.1f%% %5d x  %s" % (v / 1e6, 100 * v / tot, n, name[:110]))
