"""Deblocking loop filter of a frame on the device (dav1d_cuda_loopfilter_frame) against the reference's OWN
loop-filter path: masks and levels from dav1d_create_lf_mask_intra / _inter (src/lf_mask.c) over the block
records of synthetic frames, limits from dav1d_calc_eih, filtering by dav1d_loopfilter_sbrow_cols / _rows
(src/lf_apply_tmpl.c) + loop_filter_sb (src/loopfilter_tmpl.c), all compiled where they lie (oracle/ref_lf.c).
The device gets the reference's mask / level / limit structures as they are when the filter starts."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import _d1pkg
import reflf

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import binding as B  # noqa: E402
from dav1d_mirror_b200 import frame as F  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loopfilter_md5.json")

CASES = {
    # name: (w, h, bdmax, seed, frame kwargs, sharpness)
    "420_8b": (256, 192, 0xff, 61, {"p_intra": 0.4}, 0),
    "420_10b_sharp3": (320, 256, 0x3ff, 62, {"p_intra": 0.3, "p_tx_split": 0.6}, 3),
    "444_12b": (256, 192, 0xfff, 63, {"ss_hor": 0, "ss_ver": 0, "p_intra": 0.5}, 0),
    "422_10b_sharp7": (264, 200, 0x3ff, 64, {"ss_hor": 1, "ss_ver": 0, "p_intra": 0.5}, 7),
    "luma_8b_intra": (256, 256, 0xff, 65, {"no_chroma": 1, "p_intra": 1.0}, 1),
    "420_10b_ragged": (328, 200, 0x3ff, 66, {"p_intra": 0.2, "p_residual": 0.3}, 0),
    "420_8b_inter_skip": (384, 320, 0xff, 67, {"p_intra": 0.0, "p_residual": 0.2}, 2),
    "420_10b_720p": (1280, 720, 0x3ff, 68, {"p_intra": 0.3}, 0),
}


def make(name):
    w, h, bd, seed, kw, sharp = CASES[name]
    hf = F.HostFrame(w, h, bd, seed, real_blocks=1, p_wedge=0.0, p_warp=0.0, **kw)
    return hf, reflf.blocky_planes(hf, seed + 1000), seed, sharp


def md5_planes(planes):
    m = hashlib.md5()
    for p in planes:
        m.update(np.ascontiguousarray(p).tobytes())
    return m.hexdigest()


@pytest.mark.parametrize("name", list(CASES))
def test_reference_loopfilter_matches_golden(ref, name):
    """CPU: generator + the reference's loop filter reproduce the committed checksums (tools/make_golden.py), and
    the filter really has work to do on these pictures."""
    hf, src, seed, sharp = make(name)
    out, st = reflf.run_reference_lf(ref, hf, [p.copy() for p in src], seed, sharpness=sharp)
    assert st["sizeof_av1filter"] == 1348
    assert all(0.05 < float((a != b).mean()) for a, b in zip(src, out))
    with open(GOLDEN) as f:
        assert md5_planes(out) == json.load(f)[name], name


def run_gpu(hf, src, st):
    L = pkg.lib()
    ctx = F.open_context(0)
    pic = B.Picture()
    assert L.dav1d_cuda_picture_alloc(ctx, C.byref(pic), hf.w, hf.h, hf.ss_hor, hf.ss_ver, hf.bdmax) == 0
    d_masks = L.dav1d_cuda_malloc(st["masks"].nbytes)
    d_level = L.dav1d_cuda_malloc(st["level"].nbytes)
    try:
        for pl, a in enumerate(src):
            L.dav1d_cuda_picture_upload(ctx, C.byref(pic), pl, a.ctypes.data, a.strides[0])
        L.dav1d_cuda_upload(ctx, d_masks, st["masks"].ctypes.data, st["masks"].nbytes)
        L.dav1d_cuda_upload(ctx, d_level, st["level"].ctypes.data, st["level"].nbytes)
        lf = B.LfFrame()
        lf.w4, lf.h4, lf.b4_stride, lf.sb128w = st["w4"], st["h4"], st["b4_stride"], st["sb128w"]
        lf.filter_uv = 0 if hf.no_chroma else 1
        lf.masks, lf.level = d_masks, d_level
        C.memmove(lf.lut_e, st["lut"].ctypes.data, 64)
        C.memmove(lf.lut_i, st["lut"].ctypes.data + 64, 64)
        assert L.dav1d_cuda_loopfilter_frame(ctx, C.byref(pic), C.byref(lf)) == 0
        out = []
        for pl, a in enumerate(src):
            o = np.zeros_like(a)
            L.dav1d_cuda_picture_download(ctx, C.byref(pic), pl, o.ctypes.data, o.strides[0])
            out.append(o)
        L.dav1d_cuda_synchronize(ctx)
        pkg.check_error()
    finally:
        L.dav1d_cuda_free(d_masks)
        L.dav1d_cuda_free(d_level)
        L.dav1d_cuda_picture_free(ctx, C.byref(pic))
        L.dav1d_cuda_close(ctx)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_loopfilter_equals_the_reference(ref, name):
    hf, src, seed, sharp = make(name)
    want, st = reflf.run_reference_lf(ref, hf, [p.copy() for p in src], seed, sharpness=sharp)
    got = run_gpu(hf, src, st)
    for pl, (a, b) in enumerate(zip(want, got)):
        bad = np.argwhere(a != b)
        assert bad.size == 0, f"{name}: plane {pl}: {len(bad)} pixels differ, first at (y,x)={bad[0]}"


@pytest.mark.gpu
def test_cuda_loopfilter_random_frames(ref):
    rng = np.random.default_rng(20261020)
    for k in range(10):
        lay = [(1, 1), (1, 0), (0, 0)][rng.integers(3)]
        w, h = int(rng.integers(8, 50)) * 8, int(rng.integers(8, 36)) * 8
        bd = [0xff, 0x3ff, 0xfff][rng.integers(3)]
        hf = F.HostFrame(w, h, bd, 700 + k, real_blocks=1, p_wedge=0.0, p_warp=0.0, ss_hor=lay[0], ss_ver=lay[1],
                         p_intra=float(rng.choice([0.0, 0.3, 1.0])), p_tx_split=float(rng.choice([0, 0.5, 1.0])),
                         p_residual=float(rng.choice([0.2, 0.6, 1.0])))
        src = reflf.blocky_planes(hf, 800 + k)
        want, st = reflf.run_reference_lf(ref, hf, [p.copy() for p in src], 900 + k, sharpness=int(rng.integers(8)),
                                          p_zero_level=int(rng.choice([0, 100, 400])))
        got = run_gpu(hf, src, st)
        assert all(np.array_equal(a, b) for a, b in zip(want, got)), (k, w, h, hex(bd), lay)


def test_loopfilter_rejects_bad_arguments():
    L = pkg.lib()
    assert L.dav1d_cuda_loopfilter_frame(None, None, None) == -22
