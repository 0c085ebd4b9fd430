#!/usr/bin/env python
"""Stall samples by source line and reason from `ncu -i rep --page source --csv --print-source cuda,sass > mix.csv`."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors='replace')))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    names = {'samp': '# Samples', 'inst': 'Instructions Executed', 'long': 'stall_long_sb', 'wait': 'stall_wait',
             'short': 'stall_short_sb', 'noinst': 'stall_no_inst', 'barrier': 'stall_barrier', 'lg': 'stall_lg',
             'math': 'stall_math', 'branch': 'stall_branch_resolving', 'sel': 'stall_selected', 'mio': 'stall_mio',
             'notsel': 'stall_not_selected', 'dispatch': 'stall_dispatch'}
    agg, cur, idx = {}, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            idx = {h: i for i, h in enumerate(r)}
            continue
        if cur is None or idx is None or not r[0].strip().isdigit():
            continue
        a = agg.setdefault((cur, int(r[0])), dict(src=r[1], **{k: 0.0 for k in names}))
        for k, col in names.items():
            try:
                a[k] += float(r[idx[col]] or 0)
            except (ValueError, KeyError, IndexError):
                pass
    tot = {k: sum(a[k] for a in agg.values()) for k in names}
    print('totals: ' + '  '.join('%s %d (%.1f%%)' % (k, tot[k], 100 * tot[k] / max(tot['samp'], 1)) for k in names if k != 'inst'))
    print('warp instructions executed: %d' % tot['inst'])
    print('top lines by stall samples (share of all samples | of long-scoreboard samples | of instructions):')
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]['samp'])[:top]:
        print('%-20s %4d samp %5.2f%% long %5.2f%% inst %5.2f%% | %s' % (
            key[0][:20], key[1], 100 * a['samp'] / tot['samp'], 100 * a['long'] / max(tot['long'], 1), 100 * a['inst'] / tot['inst'], a['src'].strip()[:84]))


if __name__ == '__main__':
    main()
