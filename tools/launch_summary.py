"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; data = rows[hi + 1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict(); n = 0
for r in data:
    if len(r) <= vi: continue
    name = r[ki].split('(')[0][:70]; v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += v; n += 1
tot = sum(v for _, v in agg.values())
print(f"launches {n}  total {tot:.1f} us")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{c:5d} {v:10.1f} us {v / tot * 100:5.1f}%  avg {v / c:8.2f}  {k}")
