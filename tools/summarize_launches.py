"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
    agg[r[ki]].append(v)
tot = sum(sum(v) for v in agg.values())
print("| kernel | launches | mean (us) | total (ms) | share |")
print("|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("| `%s` | %d | %.1f | %.1f | %.2f %% |" % (k[:90], len(v), sum(v) / len(v), sum(v) / 1e3, 100 * sum(v) / tot))
print("\ntotal kernel time %.2f s over %d launches" % (tot / 1e6, sum(len(v) for v in agg.values())))
