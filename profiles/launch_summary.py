#!/usr/bin/env python3
"""launches_*.csv (ncu --metrics gpu__time_duration.sum --csv --log-file ...) -> per-kernel summary CSV,
plus the timed-step view: the last `--steps` (k_sampler<128>, k_verify) pairs of the headline loop.
Usage: launch_summary.py launches_r1.csv > launches_r1_summary.csv"""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
iN, iV, iG = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
agg, order = {}, []
for r in rows:
    m = re.search(r'(k_[a-z0-9_]+(<[^>]*>)?)', r[iN].replace('(int)', ''))
    name = m.group(1) if m else 'other (torch)'
    ns = float(r[iV].replace(',', ''))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
    order.append((name, ns, r[iG]))
tot = sum(v[1] for v in agg.values())
print('kernel,launches,total_ms,avg_ms,share_pct')
for k, (n, ns) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f'{k},{n},{ns / 1e6:.3f},{ns / n / 1e6:.4f},{100 * ns / tot:.2f}')
# headline step: full-size verify launches and the challenge sampler right before each
big = max((int(g.strip('()').split(',')[0]) for n, _, g in order if n.startswith('k_verify')), default=0)
pairs = [(order[i - 1][1], order[i][1]) for i in range(1, len(order))
         if order[i][0].startswith('k_verify') and int(order[i][2].strip('()').split(',')[0]) == big
         and order[i - 1][0].startswith('k_sampler')]
if pairs:
    s = sum(p[0] for p in pairs) / len(pairs) / 1e6
    v = sum(p[1] for p in pairs) / len(pairs) / 1e6
    print(f'# headline step under ncu (serialised, cold): sampler {s:.3f} ms + verify {v:.3f} ms over {len(pairs)} steps;'
          f' verify share {100 * v / (s + v):.1f} %')
