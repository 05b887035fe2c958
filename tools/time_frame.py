#!/usr/bin/env python3
"""Per-launch-class timing of the group submission (4K 10-bit by default): `groups` contexts,
each submitting the frames of `per_group` streams; CUDA events around `reps` steps."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402

L = pkg.lib()
groups = int(sys.argv[1]) if len(sys.argv) > 1 else 4
per_group = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
only = sys.argv[4].split(",") if len(sys.argv) > 4 and sys.argv[4] != "all" else None
record = len(sys.argv) > 5 and sys.argv[5] == "record"      # the recorder's level pass instead of the device's
w, h, bd = 3840, 2160, 0x3ff
hfs = [F.HostFrame(w, h, bd, 1000 + i) for i in range(4)]
if record:
    for hf in hfs:
        hf.record_levels()
planes = {k: [F.random_planes(hfs[k], 7 + r + 10 * k) for r in range(3)] for k in range(4)}
main = F.open_context(0)
units = []
for g in range(groups):
    ctx = F.open_context(0)
    dfs = []
    for s in range(per_group):
        k = (g * per_group + s) % 4
        df = F.DeviceFrame(ctx, hfs[k])
        df.upload_descriptors()
        for r in range(2):
            df.upload_picture(df.refs[r], planes[k][r])
        df.upload_picture(df.dst, planes[k][2])
        dfs.append(df)
    L.dav1d_cuda_synchronize(ctx)
    units.append((ctx, dfs))
e0, e1 = L.dav1d_cuda_event_create(), L.dav1d_cuda_event_create()
evs = [L.dav1d_cuda_event_create() for _ in units]
S = groups * per_group
for name, mask in (("all", 31), ("mc_put", 1), ("mc_compound", 2), ("warp", 4), ("itx", 8), ("intra", 16),
                   ("all_graph", 31)):
    if only and name not in only:
        continue
    mfs = [F.MultiFrame(ctx, dfs, phase_mask=mask, graph=name.endswith("graph")) for ctx, dfs in units]
    for _ in range(2):
        for m in mfs:
            m.launch()
    for ctx, _ in units:
        L.dav1d_cuda_synchronize(ctx)
    t0 = time.perf_counter()
    L.dav1d_cuda_event_record(main, e0)
    for ctx, _ in units:
        L.dav1d_cuda_stream_wait_event(ctx, e0)
    for _ in range(reps):
        for m in mfs:
            m.launch()
    t1 = time.perf_counter()
    for (ctx, _), ev in zip(units, evs):
        L.dav1d_cuda_event_record(ctx, ev)
        L.dav1d_cuda_stream_wait_event(main, ev)
    L.dav1d_cuda_event_record(main, e1)
    ms = L.dav1d_cuda_event_elapsed_ms(e0, e1)
    per = ms / (reps * S) * 1e3
    alg = sum(hf.algo_bytes if mask == 31 else hf.algo_class[name] for hf in hfs) / 4
    print(f"{name:12s} {per:8.1f} us/frame  {alg / per / 1e3:8.1f} GB/s  host submit {(t1 - t0) / (reps * S) * 1e6:6.1f} us/frame",
          flush=True)
    for m in mfs:
        m.close()
pkg.check_error()
print("launches", L.dav1d_cuda_launch_count())
