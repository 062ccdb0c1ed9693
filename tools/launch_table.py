"""Ordered launch table from an ncu `--metrics gpu__time_duration.sum --csv` log: python tools/launch_table.py <csv> [filter]"""
import csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, mi, ui, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
flt = sys.argv[2] if len(sys.argv) > 2 else None
tot = {}
for idx, r in enumerate(rows[1:]):
    try:
        v = float(r[mi].replace(',', ''))
    except ValueError:
        continue
    if r[ui] == 'ns':
        v /= 1e3
    n = re.sub(r'\(.*', '', r[ki])
    tot.setdefault(n, [0, 0.0])
    tot[n][0] += 1
    tot[n][1] += v
    if flt is None or re.search(flt, n):
        print('%4d %-46s %8.1f %s' % (idx, n[-46:], v, r[gi]))
print('---- totals')
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:22]:
    print('%-60s %5d %10.1f us' % (n[:60], c, t))
print('sum', sum(t for _, t in tot.values()))
