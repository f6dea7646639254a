"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14]
hdr = rows[0]; iK = hdr.index("Kernel Name"); iV = hdr.index("Metric Value"); iU = hdr.index("Metric Unit")
n = collections.Counter(); t = collections.Counter()
for r in rows[1:]:
    if r[iK].startswith("void at::") or "at::native" in r[iK]: continue     # torch's input generation, outside the step
    v = float(r[iV].replace(",", "")) * ({"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0}.get(r[iU], 1e-6))
    n[r[iK]] += 1; t[r[iK]] += v
tot = sum(t.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in t.most_common(20): print("| `%s` | %d | %.3f | %.1f%% |" % (k[:100], n[k], v, 100 * v / tot))
print("\nTotal %.2f ms" % tot)
