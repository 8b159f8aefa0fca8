"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki][:84]
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6}.get(r[ui], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print("total %.3f ms over %d kernel launches" % (tot / 1e6, n))
print("%-84s %5s %9s %9s %6s" % ("kernel", "count", "total_ms", "us/launch", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%-84s %5d %9.3f %9.1f %5.1f%%" % (k, a[0], a[1] / 1e6, a[1] / a[0] / 1e3, 100 * a[1] / tot))
