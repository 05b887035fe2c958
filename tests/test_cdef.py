"""CDEF of a frame on the device (dav1d_cuda_cdef_frame) against the reference's OWN path: dav1d_filter_sbrow_cdef
-> dav1d_cdef_brow (src/cdef_apply_tmpl.c, with its pre-filter line / column backups) -> cdef_find_dir_c /
cdef_filter_block_c (src/cdef_tmpl.c), compiled where they lie (oracle/ref_cdef.c).  The skip mask comes from the
block records of synthetic frames, the strength index per 64x64 is random (incl. "unset")."""
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

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cdef_md5.json")

CASES = {
    # name: (w, h, bdmax, seed, frame kwargs, damping)
    "420_8b": (256, 192, 0xff, 71, {"p_intra": 0.4}, 3),
    "420_10b": (320, 256, 0x3ff, 72, {"p_intra": 0.3, "p_residual": 0.8}, 6),
    "444_12b": (256, 192, 0xfff, 73, {"ss_hor": 0, "ss_ver": 0, "p_intra": 0.5}, 5),
    "422_10b": (264, 200, 0x3ff, 74, {"ss_hor": 1, "ss_ver": 0, "p_intra": 0.5, "p_residual": 0.9}, 4),
    "luma_8b": (256, 256, 0xff, 75, {"no_chroma": 1, "p_intra": 1.0}, 4),
    "420_10b_ragged_mostly_skip": (328, 200, 0x3ff, 76, {"p_intra": 0.2, "p_residual": 0.3}, 3),
    "420_10b_720p": (1280, 720, 0x3ff, 78, {"p_intra": 0.3}, 5),
}


def make(name):
    w, h, bd, seed, kw, damping = CASES[name]
    hf = F.HostFrame(w, h, bd, seed, real_blocks=1, p_wedge=0.0, p_warp=0.0, **kw)
    rng = np.random.default_rng(seed)
    ys = [int(v) for v in rng.integers(0, 64, 8)]
    us = [0] * 8 if hf.no_chroma else [int(v) for v in rng.integers(0, 64, 8)]
    ys[0], us[1] = 0, 0                    # one index without luma, one without chroma filtering
    if not hf.no_chroma:
        ys[2], us[2] = 0, 0                # and one that filters nothing
        ys[3], us[3] = 3, 8                # secondary only / primary only
    return hf, reflf.blocky_planes(hf, seed + 1000), seed, damping, ys, us


def md5_planes(planes):
    m = hashlib.md5()
    for p in planes:
        m.update(np.ascontiguousarray(p).tobytes())
    return m.hexdigest()


@pytest.mark.parametrize("name", list(CASES))
def test_reference_cdef_matches_golden(ref, name):
    hf, src, seed, damping, ys, us = make(name)
    out, st = reflf.run_reference_cdef(ref, hf, [p.copy() for p in src], seed, damping, ys, us)
    assert 0.05 < float((src[0] != out[0]).mean())
    with open(GOLDEN) as f:
        assert md5_planes(out) == json.load(f)[name], name


def run_gpu(hf, src, st):
    L = pkg.lib()
    ctx = F.open_context(0)
    pics = [B.Picture(), B.Picture()]
    for pic in pics:
        assert L.dav1d_cuda_picture_alloc(ctx, C.byref(pic), hf.w, hf.h, hf.ss_hor, hf.ss_ver, hf.bdmax) == 0
    d_masks = L.dav1d_cuda_malloc(st["masks"].nbytes)
    try:
        for pl, a in enumerate(src):
            L.dav1d_cuda_picture_upload(ctx, C.byref(pics[0]), pl, a.ctypes.data, a.strides[0])
            junk = np.full_like(a, 1)
            L.dav1d_cuda_picture_upload(ctx, C.byref(pics[1]), pl, junk.ctypes.data, junk.strides[0])
        L.dav1d_cuda_upload(ctx, d_masks, st["masks"].ctypes.data, st["masks"].nbytes)
        p = B.CdefFrame()
        p.bw, p.bh, p.sb128w, p.damping = st["bw"], st["bh"], st["sb128w"], st["damping"]
        for k in range(8):
            p.y_strength[k], p.uv_strength[k] = st["y_strength"][k], st["uv_strength"][k]
        p.masks = d_masks
        assert L.dav1d_cuda_cdef_frame(ctx, C.byref(pics[1]), C.byref(pics[0]), C.byref(p)) == 0
        assert L.dav1d_cuda_cdef_frame(ctx, C.byref(pics[0]), C.byref(pics[0]), C.byref(p)) == -22     # in place: refused
        out = []
        for pl, a in enumerate(src):
            o = np.zeros_like(a)
            L.dav1d_cuda_picture_download(ctx, C.byref(pics[1]), pl, o.ctypes.data, o.strides[0])
            out.append(o)
        L.dav1d_cuda_synchronize(ctx)
        pkg.check_error()
    finally:
        L.dav1d_cuda_free(d_masks)
        for pic in pics:
            L.dav1d_cuda_picture_free(ctx, C.byref(pic))
        L.dav1d_cuda_close(ctx)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_cdef_equals_the_reference(ref, name):
    hf, src, seed, damping, ys, us = make(name)
    want, st = reflf.run_reference_cdef(ref, hf, [p.copy() for p in src], seed, damping, ys, us)
    got = run_gpu(hf, src, st)
    for pl, (a, b) in enumerate(zip(want, got)):
        bad = np.argwhere(a != b)
        assert bad.size == 0, f"{name}: plane {pl}: {len(bad)} pixels differ, first at (y,x)={bad[0]}"


@pytest.mark.gpu
def test_cuda_cdef_random_frames(ref):
    rng = np.random.default_rng(20261021)
    for k in range(10):
        lay = [(1, 1), (1, 0), (0, 0)][rng.integers(3)]
        w, h = int(rng.integers(8, 50)) * 8, int(rng.integers(8, 36)) * 8
        bd = [0xff, 0x3ff, 0xfff][rng.integers(3)]
        hf = F.HostFrame(w, h, bd, 720 + k, real_blocks=1, p_wedge=0.0, p_warp=0.0, ss_hor=lay[0], ss_ver=lay[1],
                         p_intra=float(rng.choice([0.0, 0.3, 1.0])), p_residual=float(rng.choice([0.2, 0.6, 1.0])))
        src = reflf.blocky_planes(hf, 820 + k) if k % 3 else F.random_planes(hf, 820 + k)
        ys = [int(v) for v in rng.integers(0, 64, 8)]
        us = [int(v) for v in rng.integers(0, 64, 8)]
        want, st = reflf.run_reference_cdef(ref, hf, [p.copy() for p in src], 920 + k, 3 + int(rng.integers(4)), ys, us,
                                            p_unset=int(rng.choice([0, 100, 400])))
        got = run_gpu(hf, src, st)
        assert all(np.array_equal(a, b) for a, b in zip(want, got)), (k, w, h, hex(bd), lay)
