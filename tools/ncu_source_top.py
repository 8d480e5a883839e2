"""Top SASS lines by stall samples from an `ncu --page source --csv` export (needs -lineinfo / --import-source on)."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, isamp, iexec = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for n, r in enumerate(rows[2:]):
        if len(r) <= isamp or not r[isamp].isdigit():
            continue
        stalls = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        data.append((int(r[isamp]), n, r[isrc].strip(), int(r[iexec] or 0), stalls))
    total = sum(d[0] for d in data)
    print(f'total samples {total}, instructions {len(data)}')
    for s, n, src, ex, st in sorted(data, reverse=True)[:top]:
        print(f'{100 * s / total:5.1f}%  #{n:5d}  exec={ex:9d}  {src[:70]:70s} {st}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
