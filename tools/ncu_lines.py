#!/usr/bin/env python3
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per source line.
usage: ncu -i rep --page source --csv --print-source cuda,sass | tools/ncu_lines.py [N]
Prints executed warp instructions, stall samples and the no_instruction / long_scoreboard share
per source line (top N by instructions), then the same per file."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
fname, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] in ('File Name', 'File Path'):
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
        ini, ilsb = hdr.index('stall_no_inst'), hdr.index('stall_long_sb')
        continue
    if hdr is None or not r[0].isdigit():
        continue
    try:
        agg[(fname, int(r[0]), r[1].strip()[:90])] = [int(r[ie]), int(r[isamp]), int(r[ini]), int(r[ilsb])]
    except (ValueError, IndexError):
        pass
tot = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print('total inst', tot, 'samples', ts, 'no_inst', sum(v[2] for v in agg.values()), 'long_sb', sum(v[3] for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% samp (noinst %4d lsb %4d)  %s:%d  %s" %
          (100 * v[0] / max(tot, 1), 100 * v[1] / max(ts, 1), v[2], v[3], k[0], k[1], k[2]))
files = {}
for k, v in agg.items():
    f = files.setdefault(k[0], [0, 0])
    f[0] += v[0]
    f[1] += v[1]
for f, v in sorted(files.items(), key=lambda kv: -kv[1][0]):
    print("file %-28s %5.1f%% inst %5.1f%% samp" % (f, 100 * v[0] / max(tot, 1), 100 * v[1] / max(ts, 1)))
