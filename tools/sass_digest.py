#!/usr/bin/env python3
"""Digest of `cuobjdump -sass` for the kernels of libdav1d_cuda.so: size, mnemonic histogram and the lines
that show how data moves (async copies, packed dot products, atomics, fences, cluster barriers).
usage: tools/sass_digest.py > profiles/r2_sass_digest.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "dav1d-mirror_b200", "build")
KERNELS = ["mc_put_tma_kernelIt", "mc_compound_tma_kernelIt", "mc_put_kernelItLb0", "mc_put_kernelItLb1", "mc_compound_kernelItLb0", "warp_batch_kernelIt",
           "itx2_task_kernelItLi0", "itx2_task_kernelItLi1", "itx2_task_kernelItLi2",
           "intra_exec_kernelIt", "intra_levels_kernel", "intra_sort_kernelILb0", "intra_scan_kernel",
           "intra_mark_kernel"]
SHOW = re.compile(r"\b(LDGSTS|LDGDEPBAR|DEPBAR|IDP|ATOMG|ATOMS|RED|MEMBAR|NANOSLEEP|UCGABAR_ARV|UCGABAR_WAIT|CGAERRBAR|CCTL|UBLKCP|UTMALDG|SYNCS|"
                  r"LD\.E\.[A-Z0-9.]*STRONG|REDUX|MATCH|VOTE|SHFL|PRMT|SHF)\b")
ins = re.compile(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);")
for obj in sorted(os.listdir(BUILD)):
    if not obj.endswith(".o"):
        continue
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in txt.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            body[cur] = []
        elif cur:
            m = ins.match(line)
            if m:
                body[cur].append(m.group(2).strip())
    for name, lines in body.items():
        if not any(k in name for k in KERNELS):
            continue
        hist = collections.Counter()
        shown = collections.OrderedDict()
        for l in lines:
            op = l.split()[1] if l.startswith("@") else l.split()[0]
            hist[op.split(".")[0]] += 1
            m = SHOW.search(l)
            if m:
                key = re.sub(r"R\d+|UR\d+|P\d|0x[0-9a-f]+", "_", l)
                shown.setdefault(key, l)
        print(f"== {name}   ({obj})")
        print(f"   {len(lines)} instructions = {len(lines) * 16 / 1024:.1f} KB")
        print("   " + "  ".join(f"{k}:{v}" for k, v in hist.most_common(22)))
        first = [l for l in shown.values() if re.search(r"UTMALDG|SYNCS|UBLKCP", l)]
        for l in (first + [l for l in shown.values() if l not in first])[:28]:
            print("      " + l)
        print()
