"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols = rows[hdr]
ki, vi = cols.index('Kernel Name'), cols.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki]).split('::')[-1][:44]
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in agg.values())
print('%-46s %6s %12s %10s %7s' % ('kernel', 'n', 'total_us', 'avg_us', 'share'))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-46s %6d %12.1f %10.1f %6.1f%%' % (k, v[0], v[1] / 1e3, v[1] / v[0] / 1e3, v[1] / tot * 100))
