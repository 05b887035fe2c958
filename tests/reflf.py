"""Run the reference's own loop-filter path (oracle/ref_lf.c inside oracle/_ref/libdav1d_ref.so) on the block
records of a HostFrame.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np


class OracleLfFrame(C.Structure):
    _fields_ = [("dst", C.c_void_p * 3), ("dst_stride", C.c_ssize_t * 3),
                ("w", C.c_int32), ("h", C.c_int32), ("ss_hor", C.c_int32), ("ss_ver", C.c_int32),
                ("bitdepth_max", C.c_int32), ("no_chroma", C.c_int32),
                ("blocks", C.c_void_p), ("n_blocks", C.c_int32), ("seed", C.c_uint64),
                ("sharpness", C.c_int32), ("p_zero_level", C.c_int32), ("run", C.c_int32),
                ("masks", C.c_void_p), ("level", C.c_void_p), ("lut", C.c_void_p),
                ("b4_stride", C.c_int32), ("sb128w", C.c_int32), ("sb128h", C.c_int32), ("w4", C.c_int32),
                ("h4", C.c_int32), ("sizeof_av1filter", C.c_int32)]


def blocky_planes(hf, seed):
    """A picture the loop filter has work on: flat 4x4 / 8x8 / 16x16 patches with steps of all sizes between
    them plus low noise, so that every branch (filter mask on / off, hev, flat8in, flat8out) is taken."""
    rng = np.random.default_rng(seed)
    dt = np.uint16 if hf.hbd else np.uint8
    sh = {0xff: 0, 0x3ff: 2, 0xfff: 4}[hf.bdmax]
    out = []
    for pl in range(1 if hf.no_chroma else 3):
        ph, pw = hf.plane_shape(pl)
        acc = np.zeros((ph, pw), dtype=np.int64)
        for cell, amp in ((16, 40), (8, 10), (4, 3)):
            gh, gw = (ph + cell - 1) // cell, (pw + cell - 1) // cell
            g = rng.integers(-amp, amp + 1, size=(gh, gw))
            acc += np.kron(g, np.ones((cell, cell), dtype=np.int64))[:ph, :pw]
        acc += rng.integers(-1, 2, size=(ph, pw)) * (rng.random((ph, pw)) < 0.3)
        base = rng.integers(60, 190)
        out.append(np.clip((acc + base) << sh, 0, hf.bdmax).astype(dt))
    return out


def run_reference_lf(ref, hf, planes, seed, sharpness=0, p_zero_level=100, run=True):
    """Builds masks / levels / limit table through the reference's lf_mask.c and (run) filters `planes` in place
    through lf_apply_tmpl.c.  Returns (planes, state) with state = what the device path is given."""
    assert hf.n_block_recs > 0, "generate the frame with real_blocks=1"
    of = OracleLfFrame()
    for pl, a in enumerate(planes):
        of.dst[pl] = a.ctypes.data
        of.dst_stride[pl] = a.strides[0]
    of.w, of.h, of.ss_hor, of.ss_ver = hf.w, hf.h, hf.ss_hor, hf.ss_ver
    of.bitdepth_max, of.no_chroma = hf.bdmax, hf.no_chroma
    of.blocks, of.n_blocks = hf.blocks.ctypes.data, hf.n_block_recs
    of.seed, of.sharpness, of.p_zero_level, of.run = seed, sharpness, p_zero_level, 1 if run else 0
    sfx = "16bpc" if hf.hbd else "8bpc"
    geo = getattr(ref.lib, "oracle_lf_geometry_" + sfx)
    geo.argtypes = [C.POINTER(OracleLfFrame)]
    geo.restype = None
    geo(C.byref(of))
    masks = np.zeros(of.sb128w * of.sb128h * of.sizeof_av1filter, dtype=np.uint8)
    level = np.zeros(of.b4_stride * 32 * of.sb128h * 4, dtype=np.uint8)
    lut = np.zeros(144, dtype=np.uint8)
    of.masks, of.level, of.lut = masks.ctypes.data, level.ctypes.data, lut.ctypes.data
    fn = getattr(ref.lib, "oracle_lf_frame_" + sfx)
    fn.argtypes = [C.POINTER(OracleLfFrame)]
    fn.restype = C.c_int
    r = fn(C.byref(of))
    if r:
        raise RuntimeError(f"oracle_lf_frame: {r}")
    state = {"masks": masks, "level": level, "lut": lut, "b4_stride": of.b4_stride, "sb128w": of.sb128w,
             "sb128h": of.sb128h, "w4": of.w4, "h4": of.h4, "sizeof_av1filter": of.sizeof_av1filter}
    return planes, state


class OracleCdefFrame(C.Structure):
    _fields_ = [("dst", C.c_void_p * 3), ("dst_stride", C.c_ssize_t * 3),
                ("w", C.c_int32), ("h", C.c_int32), ("ss_hor", C.c_int32), ("ss_ver", C.c_int32),
                ("bitdepth_max", C.c_int32), ("no_chroma", C.c_int32),
                ("blocks", C.c_void_p), ("n_blocks", C.c_int32), ("seed", C.c_uint64),
                ("damping", C.c_int32), ("y_strength", C.c_uint8 * 8), ("uv_strength", C.c_uint8 * 8),
                ("p_unset", C.c_int32), ("run", C.c_int32), ("masks", C.c_void_p),
                ("sb128w", C.c_int32), ("sb128h", C.c_int32), ("bw", C.c_int32), ("bh", C.c_int32)]


def run_reference_cdef(ref, hf, planes, seed, damping, y_strength, uv_strength, p_unset=100, run=True):
    """CDEF of `planes` in place through dav1d_filter_sbrow_cdef / dav1d_cdef_brow; returns (planes, state) with
    state = the Av1Filter array (cdef_idx, noskip_mask) and the frame parameters the device path is given."""
    assert hf.n_block_recs > 0 and hf.w % 8 == 0 and hf.h % 8 == 0
    of = OracleCdefFrame()
    for pl, a in enumerate(planes):
        of.dst[pl] = a.ctypes.data
        of.dst_stride[pl] = a.strides[0]
    of.w, of.h, of.ss_hor, of.ss_ver = hf.w, hf.h, hf.ss_hor, hf.ss_ver
    of.bitdepth_max, of.no_chroma = hf.bdmax, hf.no_chroma
    of.blocks, of.n_blocks = hf.blocks.ctypes.data, hf.n_block_recs
    of.seed, of.damping, of.p_unset, of.run = seed, damping, p_unset, 1 if run else 0
    for k in range(8):
        of.y_strength[k], of.uv_strength[k] = y_strength[k], uv_strength[k]
    sb128w, sb128h = (hf.w + 127) // 128, (hf.h + 127) // 128
    masks = np.zeros(sb128w * sb128h * 1348, dtype=np.uint8)
    of.masks = masks.ctypes.data
    fn = getattr(ref.lib, "oracle_cdef_frame_" + ("16bpc" if hf.hbd else "8bpc"))
    fn.argtypes = [C.POINTER(OracleCdefFrame)]
    fn.restype = C.c_int
    r = fn(C.byref(of))
    if r:
        raise RuntimeError(f"oracle_cdef_frame: {r}")
    assert (of.sb128w, of.sb128h) == (sb128w, sb128h)
    return planes, {"masks": masks, "sb128w": sb128w, "bw": of.bw, "bh": of.bh, "damping": damping,
                    "y_strength": list(y_strength), "uv_strength": list(uv_strength)}


class OraclePfFrame(C.Structure):
    _fields_ = [("dst", C.c_void_p * 3), ("dst_stride", C.c_ssize_t * 3),
                ("w", C.c_int32), ("h", C.c_int32), ("ss_hor", C.c_int32), ("ss_ver", C.c_int32),
                ("bitdepth_max", C.c_int32), ("no_chroma", C.c_int32),
                ("blocks", C.c_void_p), ("n_blocks", C.c_int32), ("seed", C.c_uint64),
                ("do_deblock", C.c_int32), ("do_cdef", C.c_int32), ("do_lr", C.c_int32), ("run", C.c_int32),
                ("sharpness", C.c_int32), ("p_zero_level", C.c_int32), ("damping", C.c_int32), ("p_unset", C.c_int32),
                ("y_strength", C.c_uint8 * 8), ("uv_strength", C.c_uint8 * 8),
                ("unit_size_log2", C.c_int32 * 2), ("restore_planes", C.c_int32), ("p_lr_none", C.c_int32),
                ("masks", C.c_void_p), ("level", C.c_void_p), ("lut", C.c_void_p), ("lr_mask", C.c_void_p),
                ("b4_stride", C.c_int32), ("sb128w", C.c_int32), ("sb128h", C.c_int32), ("w4", C.c_int32),
                ("h4", C.c_int32), ("bw", C.c_int32), ("bh", C.c_int32), ("sizeof_av1filter", C.c_int32),
                ("sizeof_av1restoration", C.c_int32),
                ("sr_w", C.c_int32), ("sr_sb128w", C.c_int32),
                ("resize_step", C.c_int32 * 2), ("resize_start", C.c_int32 * 2),
                ("sr_dst", C.c_void_p * 3), ("sr_stride", C.c_ssize_t * 2), ("sb128", C.c_int32)]


def _c_div(a, b):
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def resize_params(src_w, dst_w):
    """f->resize_step / f->resize_start (decode.c:3365-3369 get_upscale_x0, :3576-3583 scale_fac)."""
    dx = ((src_w << 14) + (dst_w >> 1)) // dst_w
    err = dst_w * dx - (src_w << 14)
    x0 = _c_div(-((dst_w - src_w) << 13) + (dst_w >> 1), dst_w) + 128 - _c_div(err, 2)
    return dx, x0 & 0x3fff


def run_reference_chain(ref, hf, planes, seed, deblock=True, cdef=True, lr=True, sharpness=0, p_zero_level=100,
                        damping=4, y_strength=(0,) * 8, uv_strength=(0,) * 8, p_unset=100, unit_size_log2=(6, 6),
                        restore_planes=7, p_lr_none=150, run=True, sr_w=0, sb128=0):
    """The reference's own post-filter chain (dav1d_filter_sbrow per superblock row) on `planes`, in place, with the
    stages switched by deblock / cdef / lr.  Returns (planes, state): masks, levels, limit table, restoration units
    and the frame parameters the device calls take.
    sr_w > hf.w: super-resolution to that width - the result is f->sr_cur (new planes of the upscaled width; `planes`
    then holds f->cur after deblocking and CDEF), state carries sr_w, sr_sb128w, resize_step / resize_start."""
    assert hf.n_block_recs > 0 and hf.w % 8 == 0 and hf.h % 8 == 0
    of = OraclePfFrame()
    for pl, a in enumerate(planes):
        of.dst[pl] = a.ctypes.data
        of.dst_stride[pl] = a.strides[0]
    of.w, of.h, of.ss_hor, of.ss_ver = hf.w, hf.h, hf.ss_hor, hf.ss_ver
    of.bitdepth_max, of.no_chroma = hf.bdmax, hf.no_chroma
    of.blocks, of.n_blocks, of.seed = hf.blocks.ctypes.data, hf.n_block_recs, seed
    of.do_deblock, of.do_cdef, of.do_lr, of.run = int(deblock), int(cdef), int(lr), int(run)
    of.sharpness, of.p_zero_level, of.damping, of.p_unset = sharpness, p_zero_level, damping, p_unset
    for k in range(8):
        of.y_strength[k] = y_strength[k]
        of.uv_strength[k] = 0 if hf.no_chroma else uv_strength[k]
    of.unit_size_log2[0], of.unit_size_log2[1] = unit_size_log2
    of.restore_planes, of.p_lr_none = restore_planes, p_lr_none
    of.sb128 = sb128
    sfx = "16bpc" if hf.hbd else "8bpc"
    geo = getattr(ref.lib, "oracle_pf_geometry_" + sfx)
    geo.argtypes = [C.POINTER(OraclePfFrame)]
    geo.restype = None
    sr_planes = None
    if sr_w > hf.w:
        of.sr_w = sr_w
        sr_planes = [np.zeros((a.shape[0], sr_w if pl == 0 else (sr_w + hf.ss_hor) >> hf.ss_hor), dtype=a.dtype)
                     for pl, a in enumerate(planes)]
        for pl, a in enumerate(sr_planes):
            of.sr_dst[pl] = a.ctypes.data
        of.sr_stride[0] = sr_planes[0].strides[0]
        of.sr_stride[1] = sr_planes[-1].strides[0]
        in_cw, out_cw = (hf.w + hf.ss_hor) >> hf.ss_hor, (sr_w + hf.ss_hor) >> hf.ss_hor
        (of.resize_step[0], of.resize_start[0]), (of.resize_step[1], of.resize_start[1]) = \
            resize_params(hf.w, sr_w), resize_params(in_cw, out_cw)
    geo(C.byref(of))
    n128 = of.sb128w * of.sb128h
    masks = np.zeros(n128 * of.sizeof_av1filter, dtype=np.uint8)
    level = np.zeros(of.b4_stride * 32 * of.sb128h * 4, dtype=np.uint8)
    lut = np.zeros(144, dtype=np.uint8)
    lr_mask = np.zeros(of.sr_sb128w * of.sb128h * of.sizeof_av1restoration, dtype=np.uint8)
    of.masks, of.level, of.lut, of.lr_mask = masks.ctypes.data, level.ctypes.data, lut.ctypes.data, lr_mask.ctypes.data
    fn = getattr(ref.lib, "oracle_pf_frame_" + sfx)
    fn.argtypes = [C.POINTER(OraclePfFrame)]
    fn.restype = C.c_int
    r = fn(C.byref(of))
    if r:
        raise RuntimeError(f"oracle_pf_frame: {r}")
    state = {"masks": masks, "level": level, "lut": lut, "lr_mask": lr_mask, "b4_stride": of.b4_stride,
             "sb128w": of.sb128w, "sb128h": of.sb128h, "w4": of.w4, "h4": of.h4, "bw": of.bw, "bh": of.bh,
             "damping": damping, "y_strength": list(y_strength),
             "uv_strength": [0] * 8 if hf.no_chroma else list(uv_strength),
             "unit_size_log2": tuple(unit_size_log2), "restore_planes": restore_planes & (1 if hf.no_chroma else 7),
             "sizeof_av1restoration": of.sizeof_av1restoration, "sr_w": sr_w if sr_planes else 0,
             "sr_sb128w": of.sr_sb128w, "resize_step": tuple(of.resize_step), "resize_start": tuple(of.resize_start)}
    return (sr_planes if sr_planes else planes), state
