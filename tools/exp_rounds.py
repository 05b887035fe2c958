#!/usr/bin/env python3
"""Experiment (needs `make -C dav1d-mirror_b200 EXTRA=-DD1_EXPERIMENT`): per-round timeline of the
intra executor for one group of 4K frames.  usage: exp_rounds.py [frames]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = pkg.lib()
hfs = [F.HostFrame(3840, 2160, 0x3ff, 1000 + i) for i in range(4)]
ctx = F.open_context(0)
dfs = []
for s in range(n):
    hf = hfs[s % 4]
    df = F.DeviceFrame(ctx, hf)
    df.upload_descriptors()
    for r in range(2):
        df.upload_picture(df.refs[r], F.random_planes(hf, 7 + r))
    df.upload_picture(df.dst, F.random_planes(hf, 99))
    dfs.append(df)
mf = F.MultiFrame(ctx, dfs, phase_mask=16)
for _ in range(3):
    mf.launch()
L.dav1d_cuda_synchronize(ctx)
tr = np.zeros(256 * 6, dtype=np.uint64)
L.dav1d_cuda_debug_rounds_trace.argtypes = [C.c_void_p, C.c_int]
assert L.dav1d_cuda_debug_rounds_trace(tr.ctypes.data, 256) == 0
tr = tr.reshape(256, 6)
t00 = int(tr[0, 0])
print("round  pend    entries  scan_us  sort_us  exec_us   t_end_us")
for r in range(256):
    if tr[r, 0] == 0:
        break
    t0, t1, t2, t3, npend, nready = [int(v) for v in tr[r]]
    print(f"{r:4d} {npend:8d} {nready:8d} {(t1 - t0) / 1e3:8.1f} {(t2 - t1) / 1e3:8.1f} {(t3 - t2) / 1e3:8.1f} {(t3 - t00) / 1e3:10.1f}")
pkg.check_error()
