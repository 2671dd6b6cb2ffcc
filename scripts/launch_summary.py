"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST step."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
def us(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
idx = [i for i, r in enumerate(rows) if 'import_' in r['Kernel Name']]
step = rows[idx[-1]:]
tot = sum(us(r) for r in step)
print('last step: %d launches, %.2f ms' % (len(step), tot / 1e3))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    k = re.sub(r'<.*', '', r['Kernel Name'].split('(')[0]).replace('cb200::', '').replace('void ', '')
    agg[k][0] += 1; agg[k][1] += us(r)
for k, (c, v) in sorted(agg.items(), key=lambda t: -t[1][1]):
    print('%-34s n=%3d %9.1f us %5.1f%%' % (k, c, v, 100 * v / tot))
if len(sys.argv) > 2:
    for r in step:
        if re.search(sys.argv[2], r['Kernel Name']):
            print('%9.1f us grid=%-16s blk=%-10s %s' % (us(r), r['Grid Size'], r['Block Size'], r['Kernel Name'][:90]))
