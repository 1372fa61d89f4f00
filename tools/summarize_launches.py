"""ncu launch list (gpu__time_duration.sum, --csv) -> per-kernel totals and shares:  python tools/summarize_launches.py <in.csv> <out.csv> "<header comment>" """
import collections, csv, sys
src, dst, comment = sys.argv[1:4]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(',', ''))
    v = v / 1000 if r[iu] in ('ns', 'nsecond') else (v * 1000 if r[iu] in ('ms', 'msecond') else v)
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(a[1] for a in agg.values())
with open(dst, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['# ' + comment])
    w.writerow(['kernel', 'launches', 'total_us', 'share_pct', 'us_per_launch'])
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([name, a[0], '%.1f' % a[1], '%.2f' % (100 * a[1] / total), '%.2f' % (a[1] / a[0])])
    w.writerow(['TOTAL', sum(a[0] for a in agg.values()), '%.1f' % total, '100.00', ''])
print('wrote', dst, 'kernels', len(agg), 'total_us %.1f' % total)
