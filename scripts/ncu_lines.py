#!/usr/bin/env python
"""Per-source-line totals (stall samples, executed warp instructions, static SASS size) from
`ncu -i rep --page source --csv --print-source cuda,sass > mix.csv`."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    agg, cur, hdr, last = {}, None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            continue
        if r[0] == 'Line No':
            hdr = r
            iS, iI = hdr.index('# Samples'), hdr.index('Instructions Executed')
            continue
        if cur is None or hdr is None:
            continue
        if r[0].strip().isdigit():
            last = (cur, int(r[0]))
            a = agg.setdefault(last, [0.0, 0.0, 0, r[1]])
            try:
                a[0] += float(r[iS] or 0)
                a[1] += float(r[iI] or 0)
            except ValueError:
                pass
        elif last is not None and len(r) > 2 and r[2].strip():
            agg[last][2] += 1
    ts, ti, tz = (sum(a[i] for a in agg.values()) for i in range(3))
    print('total samples %d  warp-instructions %d  static SASS %d' % (ts, ti, tz))
    byf = {}
    for (f, _), a in agg.items():
        b = byf.setdefault(f, [0, 0, 0])
        for i in range(3):
            b[i] += a[i]
    for f, b in sorted(byf.items(), key=lambda kv: -kv[1][1]):
        print('  %-28s samples %5.1f%%  inst %5.1f%%  sass %6d' % (f, 100 * b[0] / ts, 100 * b[1] / ti, b[2]))
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%-22s %4d samp %5.2f%% inst %5.2f%% sass %5d | %s' % (f, l, 100 * a[0] / ts, 100 * a[1] / ti, a[2], a[3].strip()[:88]))


if __name__ == '__main__':
    main()
