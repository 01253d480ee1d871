"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) > vi:
        d[r[ki][:60]][0] += 1
        d[r[ki][:60]][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-62s %4d %10.1f us %8.1f us each %5.1f%%" % (k, v[0], v[1] / 1e3, v[1] / 1e3 / v[0], 100 * v[1] / tot))
