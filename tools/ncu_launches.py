#!/usr/bin/env python3
"""Aggregate an ncu launch list (--csv --metrics gpu__time_duration.sum[,smsp__inst_executed.sum]) per kernel.
usage: ncu_launches.py file.csv"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
agg = {}
for r in rows[1:]:
    a = agg.setdefault(r[ik][:64], {})
    a.setdefault(r[im], []).append(float(r[iv].replace(',', '')))
tot = sum(sum(a.get('gpu__time_duration.sum', [0])) for a in agg.values())
for k, a in agg.items():
    t = a.get('gpu__time_duration.sum', [0])
    ins = a.get('smsp__inst_executed.sum')
    s = f"{k:64s} n={len(t):4d} total={sum(t) / 1e3:10.1f} us ({100 * sum(t) / max(tot, 1):5.1f}%) each={sum(t) / len(t) / 1e3:9.1f}"
    if ins:
        s += f"  inst/launch={sum(ins) / len(ins) / 1e6:8.2f}M"
    print(s)
