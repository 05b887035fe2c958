"""Loop restoration on the device (dav1d_cuda_lr_frame) and the whole device post-filter chain (deblock in place ->
CDEF out of place -> [super-resolution out of place] -> loop restoration out of place) against the reference's OWN chain: dav1d_filter_sbrow()
(src/recon_tmpl.c:2149-2160) per superblock row with dav1d_copy_lpf's line backups, dav1d_cdef_brow and
dav1d_lr_sbrow / lr_stripe / wiener_c / sgr_*_c, compiled where they lie (oracle/ref_pf.c)."""
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

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postfilter_md5.json")

CASES = {
    # name: (w, h, bdmax, seed, frame kwargs, (deblock, cdef, lr), (luma, chroma) log2 unit size)
    "lr_420_8b_u64": (256, 192, 0xff, 81, {"p_intra": 0.4}, (0, 0, 1), (6, 6)),
    "lr_420_10b_u128_u64": (320, 256, 0x3ff, 82, {"p_intra": 0.3}, (0, 0, 1), (7, 6)),
    "lr_444_12b_u64": (256, 200, 0xfff, 83, {"ss_hor": 0, "ss_ver": 0, "p_intra": 0.5}, (0, 0, 1), (6, 6)),
    "lr_422_10b_u64_u32": (264, 200, 0x3ff, 84, {"ss_hor": 1, "ss_ver": 0, "p_intra": 0.5}, (0, 0, 1), (6, 5)),
    "lr_luma_8b_u256": (512, 320, 0xff, 85, {"no_chroma": 1, "p_intra": 1.0}, (0, 0, 1), (8, 8)),
    "lr_420_10b_ragged_u64_u32": (328, 184, 0x3ff, 86, {"p_intra": 0.2}, (0, 0, 1), (6, 5)),
    "chain_420_8b": (256, 192, 0xff, 91, {"p_intra": 0.4}, (1, 1, 1), (6, 6)),
    "chain_420_10b": (384, 256, 0x3ff, 92, {"p_intra": 0.3, "p_residual": 0.8}, (1, 1, 1), (6, 5)),
    "chain_444_12b": (256, 192, 0xfff, 93, {"ss_hor": 0, "ss_ver": 0, "p_intra": 0.5}, (1, 1, 1), (7, 7)),
    "chain_422_10b": (264, 200, 0x3ff, 94, {"ss_hor": 1, "ss_ver": 0, "p_intra": 0.5}, (1, 1, 1), (6, 6)),
    "chain_luma_8b": (256, 256, 0xff, 95, {"no_chroma": 1, "p_intra": 1.0}, (1, 1, 1), (6, 6)),
    "chain_420_10b_no_cdef": (320, 200, 0x3ff, 96, {"p_intra": 0.3}, (1, 0, 1), (6, 6)),
    "chain_420_10b_720p": (1280, 720, 0x3ff, 97, {"p_intra": 0.3}, (1, 1, 1), (6, 6)),
    # super-resolution (9th field: the upscaled width, frame_hdr->width[1]; denominators 9..16 of 8): the reference
    # runs dav1d_filter_sbrow_resize between CDEF and loop restoration, backs up RESIZED deblocked lines and indexes
    # the restoration units with f->sr_sb128w
    "sr_chain_420_8b_d16": (128, 192, 0xff, 101, {"p_intra": 0.4}, (1, 1, 1), (6, 6), 256),
    "sr_chain_420_10b_d11": (256, 200, 0x3ff, 102, {"p_intra": 0.3}, (1, 1, 1), (6, 5), 352),
    "sr_chain_444_12b_d9": (320, 192, 0xfff, 103, {"ss_hor": 0, "ss_ver": 0, "p_intra": 0.5}, (1, 1, 1), (7, 7), 360),
    "sr_chain_422_10b_d13": (192, 136, 0x3ff, 104, {"ss_hor": 1, "ss_ver": 0, "p_intra": 0.5}, (1, 1, 1), (6, 6), 312),
    "sr_chain_luma_8b_d12": (256, 256, 0xff, 105, {"no_chroma": 1, "p_intra": 1.0}, (1, 1, 1), (8, 8), 384),
    "sr_lr_only_420_10b_d14": (232, 184, 0x3ff, 106, {"p_intra": 0.2}, (0, 0, 1), (6, 5), 406),
    "sr_no_cdef_420_8b_d15": (328, 200, 0xff, 107, {"p_intra": 0.3}, (1, 0, 1), (6, 6), 615),
    "sr_no_lr_420_10b_d10": (256, 192, 0x3ff, 108, {"p_intra": 0.3}, (1, 1, 0), (6, 6), 320),
    # 128x128 superblocks (10th field: seq_hdr->sb128): the reference filters 128-row superblock rows and looks the
    # restoration unit of a whole row up at its first line; units are at least 128 luma pixels then (obu.c:944-954)
    "sb128_chain_420_10b": (384, 328, 0x3ff, 111, {"p_intra": 0.3}, (1, 1, 1), (7, 6), 0, 1),
    "sb128_chain_444_8b_u256": (320, 264, 0xff, 112, {"ss_hor": 0, "ss_ver": 0, "p_intra": 0.5}, (1, 1, 1), (8, 8), 0, 1),
    "sb128_chain_422_12b": (264, 200, 0xfff, 113, {"ss_hor": 1, "ss_ver": 0, "p_intra": 0.5}, (1, 1, 1), (7, 7), 0, 1),
    "sb128_sr_chain_420_10b_d12": (256, 392, 0x3ff, 114, {"p_intra": 0.3}, (1, 1, 1), (8, 7), 384, 1),
}


def make(name):
    w, h, bd, seed, kw, stages, units = CASES[name][:7]
    hf = F.HostFrame(w, h, bd, seed, real_blocks=1, p_wedge=0.0, p_warp=0.0, **kw)
    rng = np.random.default_rng(seed)
    par = dict(deblock=bool(stages[0]), cdef=bool(stages[1]), lr=bool(stages[2]), sharpness=int(rng.integers(8)),
               damping=3 + int(rng.integers(4)), y_strength=[int(v) for v in rng.integers(0, 64, 8)],
               uv_strength=[int(v) for v in rng.integers(0, 64, 8)], unit_size_log2=units)
    if len(CASES[name]) > 7:
        par["sr_w"] = CASES[name][7]
    if len(CASES[name]) > 8:
        par["sb128"] = CASES[name][8]
    return hf, reflf.blocky_planes(hf, seed + 1000), seed, par


def md5_planes(planes):
    m = hashlib.md5()
    for p in planes:
        m.update(np.ascontiguousarray(p).tobytes())
    return m.hexdigest()


@pytest.mark.parametrize("name", list(CASES))
def test_reference_chain_matches_golden(ref, name):
    hf, src, seed, par = make(name)
    out, st = reflf.run_reference_chain(ref, hf, [p.copy() for p in src], seed, **par)
    assert st["sizeof_av1restoration"] == 108
    if out[0].shape == src[0].shape:
        assert 0.2 < float((src[0] != out[0]).mean())
    else:
        assert out[0].shape == (hf.h, par["sr_w"]) and st["sr_sb128w"] == (par["sr_w"] + 127) >> 7
    with open(GOLDEN) as f:
        assert md5_planes(out) == json.load(f)[name], name


def test_reference_superres_is_a_whole_plane_resize(ref):
    """What the device chain relies on: dav1d_filter_sbrow_resize (recon_tmpl.c:2104-2137), one superblock row at a
    time with its 8-row lag, leaves in f->sr_cur exactly mc.resize of every row of the filtered f->cur."""
    name = "sr_no_lr_420_10b_d10"
    hf, src, seed, par = make(name)
    cur = [p.copy() for p in src]
    out, st = reflf.run_reference_chain(ref, hf, cur, seed, **par)     # cur: deblocked + CDEF, in place
    R = ref.bpc[hf.hbd]
    for pl, (p, o) in enumerate(zip(cur, out)):
        d = np.zeros_like(o)
        args = [d.ctypes.data, d.strides[0], p.ctypes.data, p.strides[0], o.shape[1], p.shape[0], p.shape[1],
                st["resize_step"][pl > 0], st["resize_start"][pl > 0]]
        R.resize(*(args + ([hf.bdmax] if hf.hbd else [])))
        assert np.array_equal(d, o), f"plane {pl}"
        assert not np.array_equal(p, src[pl])


@pytest.mark.gpu
def test_lr_sb128_small_units_are_refused(ref):
    """sb128 with 64-pixel luma units is not a stream (obu.c:944-954) and the reference's unit lookup differs
    there (lr_apply_tmpl.c:137-143): -EINVAL, not a silently different picture."""
    hf, src, seed, par = make("chain_420_10b")
    _, st = reflf.run_reference_chain(ref, hf, [p.copy() for p in src], seed, run=False, **par)
    par = dict(par, deblock=False, cdef=False, sb128=1, unit_size_log2=(6, 6))
    st = dict(st, unit_size_log2=(6, 6))
    with pytest.raises(AssertionError):
        run_gpu(hf, src, st, par)


def run_gpu(hf, src, st, par):
    """deblock (in place, picture 0) -> CDEF (0 -> 1) -> loop restoration (src 1, deblocked 0 -> 2).
    With super-resolution (st["sr_w"]): the CDEF output AND the deblocked picture are upscaled (dav1d_cuda_resize_frame,
    1 -> 3, 0 -> 4: the reference resizes the deblocked lines it backs up row by row, lf_apply_tmpl.c:76-91, which is
    the same as taking them from the resized deblocked picture) and loop restoration runs on pictures of the
    upscaled width (src 3, deblocked 4 -> 2) with f->sr_sb128w."""
    L = pkg.lib()
    ctx = F.open_context(0)
    sr_w = st.get("sr_w", 0)
    pics = [B.Picture() for _ in range(5 if sr_w else 3)]
    for k, pic in enumerate(pics):
        pw = sr_w if sr_w and k >= 2 else hf.w
        assert L.dav1d_cuda_picture_alloc(ctx, C.byref(pic), pw, hf.h, hf.ss_hor, hf.ss_ver, hf.bdmax) == 0
    bufs = {k: L.dav1d_cuda_malloc(st[k].nbytes) for k in ("masks", "level", "lr_mask")}
    try:
        for pl, a in enumerate(src):
            L.dav1d_cuda_picture_upload(ctx, C.byref(pics[0]), pl, a.ctypes.data, a.strides[0])
        for k, d in bufs.items():
            L.dav1d_cuda_upload(ctx, d, st[k].ctypes.data, st[k].nbytes)
        cur = 0
        if par["deblock"]:
            lf = B.LfFrame()
            lf.w4, lf.h4, lf.b4_stride, lf.sb128w = st["w4"], st["h4"], st["b4_stride"], st["sb128w"]
            lf.filter_uv = 0 if hf.no_chroma else 1
            lf.masks, lf.level = bufs["masks"], bufs["level"]
            C.memmove(lf.lut_e, st["lut"].ctypes.data, 64)
            C.memmove(lf.lut_i, st["lut"].ctypes.data + 64, 64)
            assert L.dav1d_cuda_loopfilter_frame(ctx, C.byref(pics[0]), C.byref(lf)) == 0
        pre = 0
        if par["cdef"]:
            p = B.CdefFrame()
            p.bw, p.bh, p.sb128w, p.damping = st["bw"], st["bh"], st["sb128w"], st["damping"]
            for k in range(8):
                p.y_strength[k], p.uv_strength[k] = st["y_strength"][k], st["uv_strength"][k]
            p.masks = bufs["masks"]
            assert L.dav1d_cuda_cdef_frame(ctx, C.byref(pics[1]), C.byref(pics[0]), C.byref(p)) == 0
            cur = 1
        if sr_w:
            step, start = (C.c_int32 * 2)(*st["resize_step"]), (C.c_int32 * 2)(*st["resize_start"])
            assert L.dav1d_cuda_resize_frame(ctx, C.byref(pics[3]), C.byref(pics[cur]), step, start) == 0
            if par["lr"] and pre != cur:
                assert L.dav1d_cuda_resize_frame(ctx, C.byref(pics[4]), C.byref(pics[pre]), step, start) == 0
            cur, pre = 3, (4 if pre != cur else 3)
        if par["lr"]:
            q = B.LrFrame()
            q.w, q.h, q.sb128w, q.sb128 = sr_w or hf.w, hf.h, st["sr_sb128w"], par.get("sb128", 0)
            q.unit_size_log2[0], q.unit_size_log2[1] = st["unit_size_log2"]
            q.restore_planes, q.lr_mask = st["restore_planes"], bufs["lr_mask"]
            assert L.dav1d_cuda_lr_frame(ctx, C.byref(pics[2]), C.byref(pics[cur]), C.byref(pics[pre]), C.byref(q)) == 0
            cur = 2
        out = []
        for pl, a in enumerate(src):
            o = np.zeros_like(a) if not sr_w else np.zeros((a.shape[0], sr_w if pl == 0 else (sr_w + hf.ss_hor) >> hf.ss_hor), a.dtype)
            L.dav1d_cuda_picture_download(ctx, C.byref(pics[cur]), pl, o.ctypes.data, o.strides[0])
            out.append(o)
        L.dav1d_cuda_synchronize(ctx)
        pkg.check_error()
    finally:
        for d in bufs.values():
            L.dav1d_cuda_free(d)
        for pic in pics:
            L.dav1d_cuda_picture_free(ctx, C.byref(pic))
        L.dav1d_cuda_close(ctx)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_chain_equals_the_reference(ref, name):
    hf, src, seed, par = make(name)
    want, st = reflf.run_reference_chain(ref, hf, [p.copy() for p in src], seed, **par)
    got = run_gpu(hf, src, st, par)
    for pl, (a, b) in enumerate(zip(want, got)):
        bad = np.argwhere(a != b)
        assert bad.size == 0, f"{name}: plane {pl}: {len(bad)} pixels differ, first at (y,x)={bad[0]} {a[tuple(bad[0])]} vs {b[tuple(bad[0])]}"


@pytest.mark.gpu
def test_cuda_chain_random_frames(ref):
    rng = np.random.default_rng(20261022)
    for k in range(12):
        lay = [(1, 1), (1, 0), (0, 0)][rng.integers(3)]
        w, h = int(rng.integers(8, 60)) * 8, int(rng.integers(8, 40)) * 8
        bd = [0xff, 0x3ff, 0xfff][rng.integers(3)]
        hf = F.HostFrame(w, h, bd, 740 + k, real_blocks=1, p_wedge=0.0, p_warp=0.0, ss_hor=lay[0], ss_ver=lay[1],
                         p_intra=float(rng.choice([0.0, 0.3, 1.0])), p_residual=float(rng.choice([0.2, 0.6, 1.0])))
        src = reflf.blocky_planes(hf, 840 + k) if k % 3 else F.random_planes(hf, 840 + k)
        lu = int(rng.integers(6, 9))
        par = dict(deblock=bool(rng.integers(2)), cdef=bool(rng.integers(2)), lr=True, sharpness=int(rng.integers(8)),
                   damping=3 + int(rng.integers(4)), y_strength=[int(v) for v in rng.integers(0, 64, 8)],
                   uv_strength=[int(v) for v in rng.integers(0, 64, 8)],
                   unit_size_log2=(lu, lu - int(rng.integers(2)) if lay[0] else lu),
                   restore_planes=int(rng.integers(1, 8)), p_lr_none=int(rng.choice([0, 150, 500])))
        want, st = reflf.run_reference_chain(ref, hf, [p.copy() for p in src], 940 + k, **par)
        got = run_gpu(hf, src, st, par)
        assert all(np.array_equal(a, b) for a, b in zip(want, got)), (k, w, h, hex(bd), lay, par)


@pytest.mark.gpu
def test_cuda_chain_random_superres_and_sb128(ref):
    """Random frames with super-resolution (denominators 9..16) and / or 128x128 superblocks, every stage switched
    at random, against dav1d_filter_sbrow of the reference."""
    rng = np.random.default_rng(20261023)
    for k in range(12):
        lay = [(1, 1), (1, 0), (0, 0)][rng.integers(3)]
        w, h = int(rng.integers(8, 48)) * 8, int(rng.integers(8, 40)) * 8
        bd = [0xff, 0x3ff, 0xfff][rng.integers(3)]
        hf = F.HostFrame(w, h, bd, 1740 + k, real_blocks=1, p_wedge=0.0, p_warp=0.0, ss_hor=lay[0], ss_ver=lay[1],
                         p_intra=float(rng.choice([0.0, 0.3, 1.0])), p_residual=float(rng.choice([0.2, 0.6, 1.0])))
        src = reflf.blocky_planes(hf, 1840 + k) if k % 3 else F.random_planes(hf, 1840 + k)
        sb128 = int(k % 3 != 0)
        lu = int(rng.integers(7 if sb128 else 6, 9))
        par = dict(deblock=bool(rng.integers(2)), cdef=bool(rng.integers(2)), lr=bool(k % 4), sharpness=int(rng.integers(8)),
                   damping=3 + int(rng.integers(4)), y_strength=[int(v) for v in rng.integers(0, 64, 8)],
                   uv_strength=[int(v) for v in rng.integers(0, 64, 8)],
                   unit_size_log2=(lu, lu - int(rng.integers(2)) if lay == (1, 1) else lu),
                   restore_planes=int(rng.integers(1, 8)), p_lr_none=int(rng.choice([0, 150, 500])), sb128=sb128)
        if k % 3 != 1:
            par["sr_w"] = (w * int(rng.integers(9, 17)) + 4) >> 3
        want, st = reflf.run_reference_chain(ref, hf, [p.copy() for p in src], 1940 + k, **par)
        got = run_gpu(hf, src, st, par)
        assert all(np.array_equal(a, b) for a, b in zip(want, got)), (k, w, h, hex(bd), lay, par)
