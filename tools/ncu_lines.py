#!/usr/bin/env python3
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per source line.
usage: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name ... | tools/ncu_lines.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
fname, hdr, agg, cur = None, None, {}, None
for r in rows:
    if not r:
        continue
    if r[0] in ('File Name', 'File Path'):
        if r[0] == 'File Name':
            fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if hdr is None:
        continue
    if r[0].isdigit():
        cur = (fname, int(r[0]), r[1].strip()[:100])
        agg.setdefault(cur, [0, 0])
    elif len(r) > isamp and r[2] not in ('', '...') and cur:
        try:
            agg[cur][0] += int(r[ie])
            agg[cur][1] += int(r[isamp])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print('total inst', tot, 'samples', ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100 * v[0] / max(tot, 1), 100 * v[1] / max(ts, 1), k[0], k[1], k[2]))
