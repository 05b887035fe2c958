#!/usr/bin/env python3
"""Device deblocking of 4K 10-bit 4:2:0 frames: time per frame with CUDA events, `n` frames in flight on one
context.  Masks / levels come from the reference's lf_mask.c over the generator's block records (oracle)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import binding as B  # noqa: E402
from dav1d_mirror_b200 import frame as F  # noqa: E402
import refdsp  # noqa: E402
import reflf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w, h, bd = 3840, 2160, 0x3ff
L = pkg.lib()
ref = refdsp.RefDSP()
hf = F.HostFrame(w, h, bd, 1000, real_blocks=1, p_wedge=0.0, p_warp=0.0)
src = reflf.blocky_planes(hf, 5)
_, st = reflf.run_reference_lf(ref, hf, [p.copy() for p in src], 7, run=False)
ctx = F.open_context(0)
pics = []
for k in range(n):
    pic = B.Picture()
    assert L.dav1d_cuda_picture_alloc(ctx, C.byref(pic), w, h, 1, 1, bd) == 0
    for pl, a in enumerate(src):
        L.dav1d_cuda_picture_upload(ctx, C.byref(pic), pl, a.ctypes.data, a.strides[0])
    pics.append(pic)
d_masks = L.dav1d_cuda_malloc(st["masks"].nbytes)
d_level = L.dav1d_cuda_malloc(st["level"].nbytes)
L.dav1d_cuda_upload(ctx, d_masks, st["masks"].ctypes.data, st["masks"].nbytes)
L.dav1d_cuda_upload(ctx, d_level, st["level"].ctypes.data, st["level"].nbytes)
lf = B.LfFrame()
lf.w4, lf.h4, lf.b4_stride, lf.sb128w, lf.filter_uv = st["w4"], st["h4"], st["b4_stride"], st["sb128w"], 1
lf.masks, lf.level = d_masks, d_level
C.memmove(lf.lut_e, st["lut"].ctypes.data, 64)
C.memmove(lf.lut_i, st["lut"].ctypes.data + 64, 64)
e0, e1 = L.dav1d_cuda_event_create(), L.dav1d_cuda_event_create()
for _ in range(2):
    for pic in pics:
        L.dav1d_cuda_loopfilter_frame(ctx, C.byref(pic), C.byref(lf))
L.dav1d_cuda_synchronize(ctx)
L.dav1d_cuda_event_record(ctx, e0)
for _ in range(reps):
    for pic in pics:
        L.dav1d_cuda_loopfilter_frame(ctx, C.byref(pic), C.byref(lf))
L.dav1d_cuda_event_record(ctx, e1)
L.dav1d_cuda_synchronize(ctx)
ms = L.dav1d_cuda_event_elapsed_ms(e0, e1)
per = ms / (reps * n) * 1e3
samples = w * h * 1.5
alg = samples * 2 * 2 * 2          # two passes, each reads and writes the picture once (an upper bound: the
                                   # second pass of a fused design would not re-read) - 2 bytes per sample
print(f"deblock 4K 10-bit 4:2:0: {per:.1f} us/frame, {w * h / per / 1e3:.1f} Gpix/s, "
      f"{alg / per / 1e3:.0f} GB/s of 2-pass picture traffic ({alg / 1e6:.1f} MB/frame); "
      f"masks {st['masks'].nbytes / 1e6:.2f} MB + levels {st['level'].nbytes / 1e6:.2f} MB per frame")

# CDEF, out of place, same pictures: strengths of a typical stream (primary 4..12, secondary 1..2), every 64x64 set
_, cst = reflf.run_reference_cdef(ref, hf, [p.copy() for p in src], 9, 5, [20, 33, 18, 49, 9, 26, 38, 45],
                                  [16, 9, 21, 34, 5, 18, 25, 10], p_unset=0, run=False)
out = B.Picture()
assert L.dav1d_cuda_picture_alloc(ctx, C.byref(out), w, h, 1, 1, bd) == 0
d_cm = L.dav1d_cuda_malloc(cst["masks"].nbytes)
L.dav1d_cuda_upload(ctx, d_cm, cst["masks"].ctypes.data, cst["masks"].nbytes)
cp = B.CdefFrame()
cp.bw, cp.bh, cp.sb128w, cp.damping = cst["bw"], cst["bh"], cst["sb128w"], cst["damping"]
for k in range(8):
    cp.y_strength[k], cp.uv_strength[k] = cst["y_strength"][k], cst["uv_strength"][k]
cp.masks = d_cm
for pic in pics:
    L.dav1d_cuda_cdef_frame(ctx, C.byref(out), C.byref(pic), C.byref(cp))
L.dav1d_cuda_synchronize(ctx)
L.dav1d_cuda_event_record(ctx, e0)
for _ in range(reps):
    for pic in pics:
        L.dav1d_cuda_cdef_frame(ctx, C.byref(out), C.byref(pic), C.byref(cp))
L.dav1d_cuda_event_record(ctx, e1)
L.dav1d_cuda_synchronize(ctx)
ms = L.dav1d_cuda_event_elapsed_ms(e0, e1)
per = ms / (reps * n) * 1e3
alg = samples * 2 * 2              # read the deblocked picture once, write the filtered one once
import numpy as np  # noqa: E402
ns = cst["masks"].reshape(-1, 1348)[:, 1284:1348]
print(f"cdef    4K 10-bit 4:2:0: {per:.1f} us/frame, {w * h / per / 1e3:.1f} Gpix/s, {alg / per / 1e3:.0f} GB/s of "
      f"read-once / write-once picture traffic ({alg / 1e6:.1f} MB/frame); "
      f"{np.unpackbits(ns).mean() * 100:.0f}% of the 8x8 blocks carry coefficients (the others are copied)")

# loop restoration, out of place: CDEF output (`out`) + deblocked picture (pics[0]) -> pics[1]; 64x64 luma units,
# 32x32 chroma units, every unit restored (half Wiener, half self-guided)
_, pst = reflf.run_reference_chain(ref, hf, [p.copy() for p in src], 11, deblock=False, cdef=False, lr=True,
                                   unit_size_log2=(6, 5), p_lr_none=0, run=False)
d_lr = L.dav1d_cuda_malloc(pst["lr_mask"].nbytes)
L.dav1d_cuda_upload(ctx, d_lr, pst["lr_mask"].ctypes.data, pst["lr_mask"].nbytes)
q = B.LrFrame()
q.w, q.h, q.sb128w, q.sb128 = w, h, pst["sb128w"], 0
q.unit_size_log2[0], q.unit_size_log2[1] = 6, 5
q.restore_planes, q.lr_mask = 7, d_lr
dsts = pics[1:]
if not dsts:
    dsts = [B.Picture()]
    assert L.dav1d_cuda_picture_alloc(ctx, C.byref(dsts[0]), w, h, 1, 1, bd) == 0
for d in dsts[:2]:
    L.dav1d_cuda_lr_frame(ctx, C.byref(d), C.byref(out), C.byref(pics[0]), C.byref(q))
L.dav1d_cuda_synchronize(ctx)
L.dav1d_cuda_event_record(ctx, e0)
for _ in range(reps):
    for d in dsts:
        L.dav1d_cuda_lr_frame(ctx, C.byref(d), C.byref(out), C.byref(pics[0]), C.byref(q))
L.dav1d_cuda_event_record(ctx, e1)
L.dav1d_cuda_synchronize(ctx)
ms = L.dav1d_cuda_event_elapsed_ms(e0, e1)
per = ms / (reps * len(dsts)) * 1e3
alg = samples * 2 * 2
print(f"lr      4K 10-bit 4:2:0: {per:.1f} us/frame, {w * h / per / 1e3:.1f} Gpix/s, {alg / per / 1e3:.0f} GB/s of "
      f"read-once / write-once picture traffic ({alg / 1e6:.1f} MB/frame); all units restored, Wiener and self-guided mixed")

# super-resolution: a 2560-wide coded frame upscaled to 3840 (denominator 12), all three planes; with loop
# restoration the deblocked picture is upscaled too (two of these per frame)
sw = 2560
step, start = zip(reflf.resize_params(sw, w), reflf.resize_params(sw // 2, w // 2))
st_, sa_ = (C.c_int32 * 2)(*step), (C.c_int32 * 2)(*start)
small = B.Picture()
assert L.dav1d_cuda_picture_alloc(ctx, C.byref(small), sw, h, 1, 1, bd) == 0
for pl, a in enumerate(src):
    c_ = np.ascontiguousarray(a[:, :sw >> (1 if pl else 0)])
    L.dav1d_cuda_picture_upload(ctx, C.byref(small), pl, c_.ctypes.data, c_.strides[0])
for d in dsts[:2]:
    assert L.dav1d_cuda_resize_frame(ctx, C.byref(d), C.byref(small), st_, sa_) == 0
L.dav1d_cuda_synchronize(ctx)
L.dav1d_cuda_event_record(ctx, e0)
for _ in range(reps):
    for d in dsts:
        L.dav1d_cuda_resize_frame(ctx, C.byref(d), C.byref(small), st_, sa_)
L.dav1d_cuda_event_record(ctx, e1)
L.dav1d_cuda_synchronize(ctx)
ms = L.dav1d_cuda_event_elapsed_ms(e0, e1)
per = ms / (reps * len(dsts)) * 1e3
alg = samples * 2 + samples * 2 * sw / w      # read the coded picture once, write the upscaled one once
print(f"resize  2560 -> 3840 x 2160 10-bit 4:2:0: {per:.1f} us/frame, {w * h / per / 1e3:.1f} Gpix/s, "
      f"{alg / per / 1e3:.0f} GB/s of read-once / write-once picture traffic ({alg / 1e6:.1f} MB/frame)")
pkg.check_error()
