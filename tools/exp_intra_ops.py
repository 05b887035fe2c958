#!/usr/bin/env python3
"""Micro-benchmark: throughput of the intra-class kernel per (size, mode, residual) on
independent operations tiling a 4K 10-bit luma plane."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402
from dav1d_mirror_b200 import binding as B  # noqa: E402

L = pkg.lib()
W, H, BD = 3840, 2160, 0x3ff
ctx = F.open_context(0)
pic = B.Picture()
L.dav1d_cuda_picture_alloc(ctx, C.byref(pic), W, H, 1, 1, BD)
TXD = B.TX_DIMS
cf_elems = 16 << 20
cf = L.dav1d_cuda_malloc(cf_elems * 4)
rng = np.random.default_rng(0)
host_cf = rng.integers(-2000, 2000, size=cf_elems, dtype=np.int32)
L.dav1d_cuda_upload(ctx, cf, host_cf.ctypes.data, host_cf.nbytes)
e0, e1 = L.dav1d_cuda_event_create(), L.dav1d_cuda_event_create()


def run(tx, mode, residual, eob=5, txtp=0, reps=5):
    w, h = TXD[tx]
    n_x, n_y = W // w, H // h
    n = n_x * n_y
    descs = (B.IntraDesc * n)()
    sw, sh = min(w, 32), min(h, 32)
    k = 0
    for y in range(n_y):
        for x in range(n_x):
            d = descs[k]
            d.x4, d.y4 = x * w // 4, y * h // 4
            d.tile_x4_end, d.tile_y4_end = W // 4, H // 4
            d.plane, d.tw4, d.th4, d.mode = 0, w // 4, h // 4, mode
            d.flags = 1024
            d.eob = eob if residual else -1
            d.tx, d.txtp = tx, txtp
            d.coef_off = (k * sw * sh) % (cf_elems - 4096)
            d.level = 1
            k += 1
    dev = L.dav1d_cuda_malloc(C.sizeof(descs))
    L.dav1d_cuda_upload(ctx, dev, C.addressof(descs), C.sizeof(descs))
    ls = (C.c_int32 * 2)(0, n)
    best = 1e9
    for _ in range(reps):
        L.dav1d_cuda_event_record(ctx, e0)
        L.dav1d_cuda_intra_batch(ctx, C.byref(pic), W // 4, H // 4, cf, dev, ls, 1, None, None)
        L.dav1d_cuda_event_record(ctx, e1)
        best = min(best, L.dav1d_cuda_event_elapsed_ms(e0, e1))
    L.dav1d_cuda_free(dev)
    px = n * w * h
    print(f"tx={w:2d}x{h:<2d} mode={mode:3d} resid={int(residual)} eob={eob:4d}: {best*1e3:8.1f} us  n={n:6d} "
          f"{best*1e6/n:8.1f} ns/op  {px/best/1e6:8.1f} Gpix/s", flush=True)


for tx in (0, 1, 2, 3, 4):
    for mode, resid, eob in ((0, False, 0), (12, False, 0), (3, False, 0), (0, True, 0), (0, True, 5), (255, True, 5)):
        run(tx, mode, resid, eob)
pkg.check_error()
