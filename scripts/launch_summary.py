"""Kernel shares of one timed bench step from an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v *= {'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'ms': 1.0}.get(u, 1.0)
    name = r[ki].split('(')[0]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {len(rows) - 1} launches, {tot:.3f} ms summed (cold-cache, serialised ncu timings)")
print(f"# {'kernel':70s} {'n':>5s} {'ms':>10s} {'share':>7s} {'avg ms':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {v[0]:5d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}% {v[1] / v[0]:9.4f}")
