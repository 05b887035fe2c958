#!/usr/bin/env python3
"""A group of 4K 10-bit synthetic frames through dav1d_cuda_recon_group_submit, a few times -
the command profiled with ncu (see profiles/README.md).
usage: profile_group.py [frames] [reps] [phase_mask]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mask = int(sys.argv[3]) if len(sys.argv) > 3 else 31
w, h, bd = 3840, 2160, 0x3ff
hfs = [F.HostFrame(w, h, bd, 1000 + i) for i in range(min(n, 4))]
ctx = F.open_context(0)
dfs = []
for s in range(n):
    hf = hfs[s % len(hfs)]
    df = F.DeviceFrame(ctx, hf)
    df.upload_descriptors()
    for r in range(2):
        df.upload_picture(df.refs[r], F.random_planes(hf, 7 + r))
    df.upload_picture(df.dst, F.random_planes(hf, 99))
    dfs.append(df)
mf = F.MultiFrame(ctx, dfs, phase_mask=mask)
for _ in range(reps):
    mf.launch()
pkg.lib().dav1d_cuda_synchronize(ctx)
pkg.check_error()
print("ok frames", n, "launches", pkg.lib().dav1d_cuda_launch_count())
