"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= mv:
        continue
    name = re.sub(r'\(.*', '', r[kn])[:70]
    v = float(r[mv].replace(',', ''))
    v = v / 1e3 if r[mu] == 'ns' else v * 1e3 if r[mu] == 'ms' else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{v[1]/1e3:9.3f} ms {100*v[1]/tot:5.1f}%  n={v[0]:5d}  avg={v[1]/v[0]:8.1f} us  {k}")
