"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import re
import sys


def main(path, skip_until=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1e-3)
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        name = re.sub(r'^void ', '', name)
        name = re.sub(r'\(anonymous namespace\)::', '', name)
        rows.append((int(r['ID']), name, v * scale, r.get('Grid Size', ''), r.get('Block Size', '')))
    agg = collections.OrderedDict()
    for _, name, us, g, b in rows:
        a = agg.setdefault(name, [0, 0.0, g, b])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print('| kernel | launches | total us | avg us | share | grid | block |')
    print('|---|---|---|---|---|---|---|')
    for name, (n, us, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.1f | %.2f | %.1f%% | %s | %s |' % (name[:90], n, us, us / n, 100 * us / total, g, b))
    print('\ntotal %.1f us over %d launches' % (total, len(rows)))


if __name__ == '__main__':
    main(sys.argv[1])
