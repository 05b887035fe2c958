#!/usr/bin/env python3
"""Experiment driver: timings of one 4K frame (phases and whole) and of S concurrent streams."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402

L = pkg.lib()
kw = dict(a.split("=") for a in sys.argv[1:])
S = int(kw.pop("S", 8))
reps = int(kw.pop("reps", 10))
w, h, bd = int(kw.pop("w", 3840)), int(kw.pop("h", 2160)), int(kw.pop("bd", "0x3ff"), 0)
gkw = {k: (float(v) if "." in v else int(v, 0)) for k, v in kw.items()}
hfs = [F.HostFrame(w, h, bd, 1000 + i, **gkw) for i in range(min(S, 4))]
for hf in hfs:
    hf.schedule()
print("levels", [hf.n_levels for hf in hfs], "intra ops", hfs[0].n_intra, "deps", hfs[0].deps.nbytes // 4,
      "algoMB %.1f" % (hfs[0].algo_bytes / 1e6), flush=True)
main = F.open_context(0)
e0, e1 = L.dav1d_cuda_event_create(), L.dav1d_cuda_event_create()


def build(dataflow, n, classes=False, tasks=True):
    ctxs, dfs = [], []
    for s in range(n):
        ctx = F.open_context(0)
        df = F.DeviceFrame(ctx, hfs[s % len(hfs)], dataflow=dataflow, classes=classes, tasks=tasks)
        df.upload_descriptors()
        for r in range(2):
            df.upload_picture(df.refs[r], F.random_planes(df.hf, 7 + r))
        df.upload_picture(df.dst, F.random_planes(df.hf, 99))
        df.build_graph()
        L.dav1d_cuda_synchronize(ctx)
        ctxs.append(ctx)
        dfs.append(df)
    return ctxs, dfs


def timed(ctxs, dfs, steps):
    dones = [L.dav1d_cuda_event_create() for _ in ctxs]
    L.dav1d_cuda_event_record(main, e0)
    for c in ctxs:
        L.dav1d_cuda_stream_wait_event(c, e0)
    for _ in range(steps):
        for df in dfs:
            df.launch_graph()
    for c, d in zip(ctxs, dones):
        L.dav1d_cuda_event_record(c, d)
        L.dav1d_cuda_stream_wait_event(main, d)
    L.dav1d_cuda_event_record(main, e1)
    ms = L.dav1d_cuda_event_elapsed_ms(e0, e1)
    for d in dones:
        L.dav1d_cuda_event_destroy(d)
    return ms


for dataflow, classes, tasks in ((False, False, 1), (False, False, 0)):
    for n in sorted(set([1, 2, S])):
        ctxs, dfs = build(dataflow, n, classes, tasks)
        timed(ctxs, dfs, 3)
        ms = timed(ctxs, dfs, reps)
        per_frame = ms / (reps * n)
        print(f"dataflow={int(dataflow)} tasks={int(tasks)} streams={n}: {per_frame*1e3:8.1f} us/frame  "
              f"{hfs[0].luma_px / per_frame / 1e3:9.0f} Mpix/s  nodes={dfs[0].graph_nodes}", flush=True)
        if n == 1 and tasks == 1:
            print("   classes(ms):", {k: round(v, 4) for k, v in dfs[0].time_classes(reps=5).items()}, flush=True)
        if n == S and tasks == 1:
            print("   classes rotating over streams (ms):", {k: round(v, 4) for k, v in F.time_classes(dfs, reps=3).items()}, flush=True)
        for df in dfs:
            df.close()
        for c in ctxs:
            L.dav1d_cuda_close(c)
# ---- multi-frame batched graph: all S frames in one graph on one context
for n, mt in [(1, 1), (S, 0), (S, 1), (2 * S, 1)]:
    ctx = F.open_context(0)
    dfs = []
    for s_ in range(n):
        df = F.DeviceFrame(ctx, hfs[s_ % len(hfs)], dataflow=False, tasks=mt)
        df.upload_descriptors()
        for r in range(2):
            df.upload_picture(df.refs[r], F.random_planes(df.hf, 7 + r))
        df.upload_picture(df.dst, F.random_planes(df.hf, 99))
        dfs.append(df)
    L.dav1d_cuda_synchronize(ctx)
    mf = F.MultiFrame(ctx, dfs)
    for _ in range(3):
        mf.launch()
    L.dav1d_cuda_event_record(ctx, e0)
    for _ in range(reps):
        mf.launch()
    L.dav1d_cuda_event_record(ctx, e1)
    ms = L.dav1d_cuda_event_elapsed_ms(e0, e1)
    per_frame = ms / (reps * n)
    print(f"multi-frame graph tasks={mt} frames={n}: {per_frame*1e3:8.1f} us/frame  {hfs[0].luma_px / per_frame / 1e3:9.0f} Mpix/s"
          f"  nodes={mf.graph_nodes}", flush=True)
    mf.close()
    for df in dfs:
        df.close()
    L.dav1d_cuda_close(ctx)
pkg.check_error()
# ---- several multi-frame graphs in flight: G groups of n frames, each group on its own context
for G, n in [(2, S), (4, S // 2), (2, 2 * S)]:
    if n < 1:
        continue
    ctxs, mfs, alldfs = [], [], []
    for g in range(G):
        ctx = F.open_context(0)
        dfs = []
        for s_ in range(n):
            df = F.DeviceFrame(ctx, hfs[(g * n + s_) % len(hfs)], dataflow=False, tasks=1)
            df.upload_descriptors()
            for r in range(2):
                df.upload_picture(df.refs[r], F.random_planes(df.hf, 7 + r))
            df.upload_picture(df.dst, F.random_planes(df.hf, 99))
            dfs.append(df)
        L.dav1d_cuda_synchronize(ctx)
        ctxs.append(ctx)
        mfs.append(F.MultiFrame(ctx, dfs))
        alldfs.append(dfs)
    for _ in range(3):
        for mf in mfs:
            mf.launch()
    for c in ctxs:
        L.dav1d_cuda_synchronize(c)
    dones = [L.dav1d_cuda_event_create() for _ in ctxs]
    L.dav1d_cuda_event_record(main, e0)
    for c in ctxs:
        L.dav1d_cuda_stream_wait_event(c, e0)
    for _ in range(reps):
        for mf in mfs:
            mf.launch()
    for c, d in zip(ctxs, dones):
        L.dav1d_cuda_event_record(c, d)
        L.dav1d_cuda_stream_wait_event(main, d)
    L.dav1d_cuda_event_record(main, e1)
    ms = L.dav1d_cuda_event_elapsed_ms(e0, e1)
    per_frame = ms / (reps * n * G)
    print(f"multi-frame graphs in flight={G} x frames={n}: {per_frame*1e3:8.1f} us/frame"
          f"  {hfs[0].luma_px / per_frame / 1e3:9.0f} Mpix/s", flush=True)
    for mf in mfs:
        mf.close()
    for dfs in alldfs:
        for df in dfs:
            df.close()
    for c in ctxs:
        L.dav1d_cuda_close(c)
pkg.check_error()
