"""Turn an `ncu --csv` log into a per-kernel summary table (count, total ms, share, and any extra summed metrics)."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path, errors='ignore') if not l.startswith('==')]
    return list(csv.DictReader(lines))


def main():
    path = sys.argv[1]
    last = int(sys.argv[2]) if len(sys.argv) > 2 else 0       # only the last N launches (one step)
    rows = load(path)
    by_id = collections.OrderedDict()
    for r in rows:
        by_id.setdefault(r['ID'], {'name': re.sub(r'\(.*', '', r['Kernel Name'])[:70]})[r['Metric Name']] = (r['Metric Value'], r['Metric Unit'])
    launches = list(by_id.values())
    if last:
        launches = launches[-last:]
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for l in launches:
        a = agg[l['name']]
        a['n'] += 1
        for k, v in l.items():
            if k == 'name':
                continue
            val = float(v[0].replace(',', ''))
            unit = v[1]
            if k == 'gpu__time_duration.sum':
                val *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(unit, 1e-6)
            if 'bytes' in k:
                val *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit, 1e-6)
            a[k] += val
    tot = sum(a['gpu__time_duration.sum'] for a in agg.values())
    keys = sorted({k for a in agg.values() for k in a if k not in ('n', 'gpu__time_duration.sum')})
    print('%-72s %5s %10s %6s %s' % ('kernel', 'n', 'ms', 'share', ' '.join(k.replace('dram__bytes_', 'dram_').replace('.sum', '_MB') for k in keys)))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
        print('%-72s %5d %10.3f %5.1f%% %s' % (name, a['n'], a['gpu__time_duration.sum'], 100 * a['gpu__time_duration.sum'] / tot,
                                             ' '.join('%10.1f' % a[k] for k in keys)))
    print('%-72s %5d %10.3f' % ('TOTAL', sum(a['n'] for a in agg.values()), tot))


if __name__ == '__main__':
    main()
