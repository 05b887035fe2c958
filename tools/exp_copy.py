#!/usr/bin/env python3
"""Experiment: bidirectional pinned copies through the C ABI - same stream vs dedicated streams,
1-D vs 2-D downloads."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402
from dav1d_mirror_b200 import binding as B  # noqa: E402

L = pkg.lib()
main = F.open_context(0)
N = 4
ctxs = [F.open_context(0) for _ in range(N)]
upc, dnc = F.open_context(0), F.open_context(0)
n = 64 << 20
hu = [L.dav1d_cuda_host_alloc(n) for _ in range(N)]
hd = [L.dav1d_cuda_host_alloc(n) for _ in range(N)]
du = [L.dav1d_cuda_malloc(n) for _ in range(N)]
dd = [L.dav1d_cuda_malloc(n) for _ in range(N)]
pics = []
for i in range(N):
    p = B.Picture()
    L.dav1d_cuda_picture_alloc(ctxs[i], C.byref(p), 3840, 2160, 1, 1, 0x3ff)
    pics.append(p)
e0, e1 = L.dav1d_cuda_event_create(), L.dav1d_cuda_event_create()


def run(mode, reps=6):
    all_ctx = ctxs + [upc, dnc]
    L.dav1d_cuda_synchronize(main)
    L.dav1d_cuda_event_record(main, e0)
    for c in all_ctx:
        L.dav1d_cuda_stream_wait_event(c, e0)
    nb = 0
    for _ in range(reps):
        for i in range(N):
            if mode == "same-1d":
                L.dav1d_cuda_upload(ctxs[i], du[i], hu[i], n)
                L.dav1d_cuda_download(ctxs[i], hd[i], dd[i], n)
                nb += 2 * n
            elif mode == "same-2d":
                L.dav1d_cuda_upload(ctxs[i], du[i], hu[i], n)
                L.dav1d_cuda_picture_download(ctxs[i], C.byref(pics[i]), 0, hd[i], 7680)
                nb += n + 7680 * 2160
            elif mode == "split-1d":
                L.dav1d_cuda_upload(upc, du[i], hu[i], n)
                L.dav1d_cuda_download(dnc, hd[i], dd[i], n)
                nb += 2 * n
            elif mode.startswith("chunks"):
                # per stream: K uploads of n/K bytes, then 24 downloads of n/24 (like one group of 8 frames)
                K = int(mode[6:])
                for k in range(K):
                    L.dav1d_cuda_upload(ctxs[i], du[i] + k * (n // K), hu[i] + k * (n // K), n // K)
                for k in range(24):
                    L.dav1d_cuda_download(ctxs[i], hd[i] + k * (n // 24), dd[i] + k * (n // 24), n // 24)
                nb += 2 * n
            elif mode == "up":
                L.dav1d_cuda_upload(ctxs[i], du[i], hu[i], n)
                nb += n
            elif mode == "down":
                L.dav1d_cuda_download(ctxs[i], hd[i], dd[i], n)
                nb += n
    dones = [L.dav1d_cuda_event_create() for _ in all_ctx]
    for c, d in zip(all_ctx, dones):
        L.dav1d_cuda_event_record(c, d)
        L.dav1d_cuda_stream_wait_event(main, d)
    L.dav1d_cuda_event_record(main, e1)
    ms = L.dav1d_cuda_event_elapsed_ms(e0, e1)
    print(f"{mode:10s} {nb / ms / 1e6:6.1f} GB/s", flush=True)


for m in ("same-1d", "chunks1", "chunks8", "chunks96", "chunks960"):
    run(m, 2)
    run(m)
pkg.check_error()
