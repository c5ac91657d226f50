"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1e-3)
    name = r[ki].split('(')[0].replace('<unnamed>::', '')
    agg[name[:58]].append(v * scale)
tot = sum(sum(v) for v in agg.values())
print("total %.2f ms over %d launches" % (tot / 1e3, sum(len(v) for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-60s n=%5d  mean %9.2f us  total %8.2f ms  share %5.1f%%" % (k, len(v), sum(v) / len(v), sum(v) / 1e3, 100 * sum(v) / tot))
