#!/usr/bin/env python3
"""Experiment: where does the end-to-end step go?  32 frames in 4 groups; per step and frame
pinned H2D of descriptors/coefficients, the group's graph, pinned D2H of the picture."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402

L = pkg.lib()
S, G = 32, 8
hfs = [F.HostFrame(3840, 2160, 0x3ff, 1000 + i) for i in range(4)]
for hf in hfs:
    hf.schedule()
main = F.open_context(0)
units = []
for g0 in range(0, S, G):
    ctx = F.open_context(0)
    dfs = []
    for s in range(g0, g0 + G):
        df = F.DeviceFrame(ctx, hfs[s % 4])
        df.upload_descriptors()
        for r in range(2):
            df.upload_picture(df.refs[r], F.random_planes(df.hf, 7 + r))
        df.upload_picture(df.dst, F.random_planes(df.hf, 99))
        df.alloc_pinned()
        dfs.append(df)
    L.dav1d_cuda_synchronize(ctx)
    units.append((ctx, F.MultiFrame(ctx, dfs), dfs))
e0, e1 = L.dav1d_cuda_event_create(), L.dav1d_cuda_event_create()
h2d = sum(df.arena_bytes for _, _, dfs in units for df in dfs)
d2h = sum(df.pinned_out_bytes for _, _, dfs in units for df in dfs)
print("per step: H2D %.0f MB  D2H %.0f MB" % (h2d / 1e6, d2h / 1e6))


def run(up, comp, down, steps=6):
    dones = [L.dav1d_cuda_event_create() for _ in units]
    L.dav1d_cuda_synchronize(main)
    t0 = time.time()
    L.dav1d_cuda_event_record(main, e0)
    for c, _, _ in units:
        L.dav1d_cuda_stream_wait_event(c, e0)
    for _ in range(steps):
        for c, mf, dfs in units:
            if up:
                for df in dfs:
                    df.upload_descriptors_pinned()
            if comp:
                mf.launch()
            if down:
                for df in dfs:
                    df.download_pinned()
    t_issue = time.time() - t0
    for (c, _, _), d in zip(units, dones):
        L.dav1d_cuda_event_record(c, d)
        L.dav1d_cuda_stream_wait_event(main, d)
    L.dav1d_cuda_event_record(main, e1)
    ms = L.dav1d_cuda_event_elapsed_ms(e0, e1) / steps
    gb = ((h2d if up else 0) + (d2h if down else 0)) / 1e9
    print(f"up={up} comp={comp} down={down}: {ms:7.2f} ms/step  issue {t_issue / steps * 1e3:6.2f} ms/step"
          f"  {gb / (ms * 1e-3):6.1f} GB/s  {S * hfs[0].luma_px / (ms * 1e-3) / 1e6:8.0f} Mpix/s", flush=True)


for cfg in [(1, 0, 0), (0, 0, 1), (1, 0, 1), (0, 1, 0), (1, 1, 0), (0, 1, 1), (1, 1, 1)]:
    run(*cfg, steps=2)
    run(*cfg)
pkg.check_error()
