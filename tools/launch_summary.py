"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    tot = defaultdict(float); cnt = defaultdict(int)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        if unit in ('us', 'usecond'):
            v *= 1e3
        elif unit in ('ms', 'msecond'):
            v *= 1e6
        name = r['Kernel Name'].split('(')[0]
        tot[name] += v; cnt[name] += 1
    total = sum(tot.values())
    print(f'{"kernel":70s} {"n":>6s} {"total_ms":>10s} {"avg_us":>9s} {"share":>7s}')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f'{k[:70]:70s} {cnt[k]:6d} {v / 1e6:10.3f} {v / cnt[k] / 1e3:9.1f} {100 * v / total:6.1f}%')
    print(f'{"TOTAL":70s} {sum(cnt.values()):6d} {total / 1e6:10.3f}')


if __name__ == '__main__':
    main(sys.argv[1])
