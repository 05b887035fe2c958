#!/usr/bin/env python3
"""All intra-class ops of a real synthetic frame launched as ONE level (dependencies ignored:
timing only) - separates per-op cost from level/tail effects."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402

L = pkg.lib()
hf = F.HostFrame(3840, 2160, 0x3ff, 1000)
hf.schedule()
ctx = F.open_context(0)
df = F.DeviceFrame(ctx, hf, dataflow=False)
df.upload_descriptors()
for r in range(2):
    df.upload_picture(df.refs[r], F.random_planes(hf, 7 + r))
df.upload_picture(df.dst, F.random_planes(hf, 99))
L.dav1d_cuda_synchronize(ctx)
e0, e1 = L.dav1d_cuda_event_create(), L.dav1d_cuda_event_create()
b = df.batch
rec = hf.intra_sorted.reshape(-1, 40)


def run(sel, label, reps=5):
    sub = np.ascontiguousarray(rec[sel]).reshape(-1)
    n = int(sel.sum()) if sel.dtype == bool else len(sel)
    dev = L.dav1d_cuda_malloc(max(sub.nbytes, 64))
    L.dav1d_cuda_upload(ctx, dev, sub.ctypes.data, sub.nbytes)
    ls = (C.c_int32 * 2)(0, n)
    best = 1e9
    for _ in range(reps):
        L.dav1d_cuda_event_record(ctx, e0)
        L.dav1d_cuda_intra_batch(ctx, b.dst, b.bw4, b.bh4, b.cf, dev, ls, 1, b.pal, b.pal_idx)
        L.dav1d_cuda_event_record(ctx, e1)
        best = min(best, L.dav1d_cuda_event_elapsed_ms(e0, e1))
    L.dav1d_cuda_free(dev)
    print(f"{label:40s} n={n:6d}  {best*1e3:8.1f} us  {best*1e6/max(n,1):7.2f} ns/op", flush=True)


tw, th, mode = rec[:, 13].astype(int), rec[:, 14].astype(int), rec[:, 15].astype(int)
eob = rec[:, 20:22].copy().view(np.int16).reshape(-1)
allsel = np.ones(len(rec), bool)
run(allsel, "all ops, one launch")
tx, txtp = rec[:, 22].astype(int), rec[:, 23].astype(int)
has_res = (rec[:, 20:22].copy().view(np.int16).reshape(-1) >= 0).astype(int)
key = ((has_res * 32 + tx * has_res) * 32 + txtp * has_res) * 256 + rec[:, 15].astype(int)
order = np.argsort(key, kind="stable")
run(order, "all ops, sorted by (tx, txtp, mode)")
key2 = (has_res * 32 + tx * has_res) * 256 + rec[:, 15].astype(int)
run(np.argsort(key2, kind="stable"), "all ops, sorted by (tx, mode)")
run(np.argsort(has_res * 32 + tx * has_res, kind="stable"), "all ops, sorted by tx")
run((tw <= 4) & (th <= 4), "ops <= 16x16")
run((tw <= 2) & (th <= 2), "ops <= 8x8")
run((tw >= 8) | (th >= 8), "ops with a side >= 32")
run((tw == 16) | (th == 16), "ops with a side == 64")
run(eob < 0, "no residual")
run(eob >= 0, "with residual")
for m, name in ((14, "CFL"), (13, "FILTER"), (15, "PAL"), (255, "NONE")):
    run(mode == m, name)
run((mode >= 1) & (mode <= 8), "directional")
run((mode == 0) | ((mode >= 9) & (mode <= 12)), "dc/smooth/paeth")
pkg.check_error()
