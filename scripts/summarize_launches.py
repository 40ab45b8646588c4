"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (and per kernel+grid) totals.
usage: python scripts/summarize_launches.py launches.csv [--by-grid]"""
import csv, sys, collections, re

path = sys.argv[1]
by_grid = '--by-grid' in sys.argv
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    key = (name, r['Grid Size'] if by_grid else '')
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    us = v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
    t = tot.setdefault(key, [0, 0.0])
    t[0] += 1
    t[1] += us
total = sum(t[1] for t in tot.values())
print('%-86s %6s %12s %7s %10s' % ('kernel' + (' / grid' if by_grid else ''), 'n', 'total_us', 'share', 'avg_us'))
for (name, grid), (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    label = (name + ('  ' + grid if by_grid else ''))[:86]
    print('%-86s %6d %12.1f %6.1f%% %10.1f' % (label, n, us, 100 * us / total, us / n))
print('%-86s %6d %12.1f' % ('TOTAL', sum(t[0] for t in tot.values()), total))
