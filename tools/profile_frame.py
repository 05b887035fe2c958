#!/usr/bin/env python3
"""One 4K 10-bit synthetic frame through dav1d_cuda_recon_submit (no graph),
a few times - the command profiled with ncu (see profiles/README.md)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402

w, h, bd = 3840, 2160, 0x3ff
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
kw = {}
for a in sys.argv[2:]:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v else int(v, 0)
w, h, bd = kw.pop("w", w), kw.pop("h", h), kw.pop("bd", bd)
hf = F.HostFrame(w, h, bd, 1000, **kw)
ctx = F.open_context(0)
df = F.DeviceFrame(ctx, hf, tasks=int(os.environ.get('D1_TASKS', '1')))
df.upload_descriptors()
for r in range(2):
    df.upload_picture(df.refs[r], F.random_planes(hf, 7 + r))
df.upload_picture(df.dst, F.random_planes(hf, 99))
for _ in range(reps):
    df.submit()
pkg.lib().dav1d_cuda_synchronize(ctx)
pkg.check_error()
print("ok intra ops", hf.n_intra, "launches", pkg.lib().dav1d_cuda_launch_count())
df.close()
