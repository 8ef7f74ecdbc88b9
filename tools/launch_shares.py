"""Summarises an `ncu --csv --metrics gpu__time_duration.sum` launch list: per-kernel count, total, average, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:64]
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"| `{k}` | {v[0]} | {v[1]:.0f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |")
