"""Join an `ncu --page source --csv` export (per-SASS-instruction executed counts) with `nvdisasm -g -c` line info of the
same kernel and print executed warp instructions per task, aggregated by CUDA source line.

    python tools/sass_by_line.py <ncu source csv> <nvdisasm txt> <kernel substring> <tasks>
"""
import csv
import re
import sys
from collections import defaultdict


def disasm_lines(path, kernel):
    """[(line_tag, sass_text)] in program order for the function whose name contains `kernel`."""
    out, cur, inside = [], None, False
    for l in open(path, errors='replace'):
        if l.startswith('.text.') or re.match(r'^\s*\.section\s+\.text\.', l):
            inside = kernel in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (m.group(1).split('/')[-1], int(m.group(2)), (inl.group(1).split('/')[-1], int(inl.group(2))) if inl else None)
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main(ncu_csv, dis_txt, kernel, tasks):
    rows = list(csv.reader(open(ncu_csv)))
    hdr = rows[1]
    isrc, iex = hdr.index('Source'), hdr.index('Instructions Executed')
    ex = []
    for r in rows[2:]:
        if len(r) > iex and r[iex].isdigit():
            ex.append((r[isrc].strip(), int(r[iex])))
    dis = disasm_lines(dis_txt, kernel)
    n = len(dis)
    ex = ex[:n]                                   # first captured launch
    assert len(ex) == n, (len(ex), n)
    agg = defaultdict(int)
    for (tag, _), (_, e) in zip(dis, ex):
        agg[tag] += e
    tot = sum(agg.values())
    print(f'{kernel}: {tot / tasks:.1f} warp instructions per task over {n} SASS instructions')
    for tag, e in sorted(agg.items(), key=lambda kv: -kv[1])[:45]:
        if tag is None:
            name = '?'
        else:
            name = f'{tag[0]}:{tag[1]}' + (f'  <- {tag[2][0]}:{tag[2][1]}' if tag[2] else '')
        print(f'{e / tasks:8.1f}  {name}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]))
