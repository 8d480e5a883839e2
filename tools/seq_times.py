"""Per-launch durations of one kernel from an ncu launch list (order of launch): python tools/seq_times.py file.csv substring [skip] [count]"""
import csv
import sys


def main(path, sub, skip=0, count=24):
    lines = [l for l in open(path) if not l.startswith('==')]
    out = []
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum' or sub not in r['Kernel Name']:
            continue
        v = float(r['Metric Value'].replace(',', ''))
        u = r.get('Metric Unit', 'ns')
        v *= {'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6}.get(u, 1.0)
        out.append(round(v / 1e3))
    print(sub, len(out), out[skip:skip + count])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0, int(sys.argv[4]) if len(sys.argv) > 4 else 24)
