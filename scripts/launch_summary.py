"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST step."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
def us(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
# a step starts at the network's first kernel (first-layer forward, or the input import); the list of a bench.py run holds
# training steps, attribution passes and inference steps: the LAST segment with the most launches is a full training step
idx = [i for i, r in enumerate(rows) if 'import_input' in r['Kernel Name'] or 'conv_first_fwd_kernel' in r['Kernel Name']]
segs = [rows[a:b] for a, b in zip(idx, idx[1:] + [len(rows)])] if idx else [rows]
# (timed steps = the most frequent length among the long segments; the attribution pass updates layer by layer and is longer)
longest = max(len(g) for g in segs)
most = collections.Counter(len(g) for g in segs if 2 * len(g) > longest).most_common(1)[0][0]
step = [g for g in segs if len(g) == most][0 if len(sys.argv) > 3 else -1]
tot = sum(us(r) for r in step)
print('%d launches in the list, %d steps; a timed training step: %d launches, %.2f ms serialised under ncu' % (len(rows), len(segs), len(step), tot / 1e3))
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
