"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum,... --csv` launch list (phase-kernel pipeline)."""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, mi, vi = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('<')[0].replace('void ', '')
    agg[name][r[mi]].append(float(r[vi].replace(',', '')))
tot = 0.0
for n, m in agg.items():
    t = m['gpu__time_duration.sum']
    tot += sum(t)
    extra = ' '.join('%s=%.1f' % (k.split('.')[0].replace('smsp__', '').replace('sm__', '').replace('l1tex__', ''), v[1 if len(v) > 1 else 0])
                     for k, v in m.items() if k != 'gpu__time_duration.sum')
    print('%-18s launches %3d total %7.2f ms  first %s  %s' % (n, len(t), sum(t) / 1e6, [round(x / 1e6, 2) for x in t[:4]], extra))
print('sum of kernel times %.2f ms' % (tot / 1e6))
