"""Run the sequential frame-level oracle (oracle/ref_frame.c inside
oracle/_ref/libdav1d_ref.so) on a HostFrame. TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

import _d1pkg

pkg = _d1pkg.load_pkg()


class OracleFrame(C.Structure):
    _fields_ = [("dst", C.c_void_p * 3), ("dst_stride", C.c_ssize_t * 3),
                ("ref", (C.c_void_p * 3) * 7), ("ref_stride", (C.c_ssize_t * 3) * 7),
                ("w", C.c_int32), ("h", C.c_int32), ("ss_hor", C.c_int32), ("ss_ver", C.c_int32),
                ("bitdepth_max", C.c_int32), ("bw4", C.c_int32), ("bh4", C.c_int32),
                ("cf", C.c_void_p), ("masks", C.c_void_p), ("pal", C.c_void_p), ("pal_idx", C.c_void_p),
                ("mc_put", C.c_void_p), ("mc_comp", C.c_void_p), ("warp", C.c_void_p), ("itx", C.c_void_p),
                ("intra", C.c_void_p), ("order", C.c_void_p), ("n_order", C.c_int32),
                ("mc_obmc", C.c_void_p), ("mc_scaled", C.c_void_p), ("ref_w", C.c_int32 * 7), ("ref_h", C.c_int32 * 7)]


def make_oracle_frame(hf, dst_planes, ref_planes_list, keep):
    """dst_planes: list of numpy planes (modified in place); ref_planes_list: list (per ref) of plane lists."""
    of = OracleFrame()
    npl = len(dst_planes)
    for pl in range(3):
        if pl < npl:
            of.dst[pl] = dst_planes[pl].ctypes.data
            of.dst_stride[pl] = dst_planes[pl].strides[0]
    for r, planes in enumerate(ref_planes_list):
        for pl in range(npl):
            of.ref[r][pl] = planes[pl].ctypes.data
            of.ref_stride[r][pl] = planes[pl].strides[0]
    of.w, of.h = hf.w, hf.h
    of.ss_hor, of.ss_ver = hf.ss_hor, hf.ss_ver
    of.bitdepth_max = hf.bdmax
    of.bw4, of.bh4 = hf.bw4, hf.bh4
    masks = hf.masks.copy()       # W_MASK writes into it
    keep.append(masks)
    for name, arr in (("cf", hf.cf), ("masks", masks), ("pal", hf.pal), ("pal_idx", hf.pal_idx),
                      ("mc_put", hf.mc_put), ("mc_comp", hf.mc_comp), ("warp", hf.warp), ("itx", hf.itx),
                      ("intra", hf.intra), ("order", hf.order), ("mc_obmc", hf.mc_obmc), ("mc_scaled", hf.mc_scaled)):
        setattr(of, name, arr.ctypes.data if arr.nbytes else None)
    of.n_order = hf.order.nbytes // 4
    for r in range(7):
        of.ref_w[r], of.ref_h[r] = hf.params.ref_w[r], hf.params.ref_h[r]
    return of


def run_oracle(ref, hf, dst_planes, ref_planes_list):
    keep = []
    of = make_oracle_frame(hf, dst_planes, ref_planes_list, keep)
    fn = ref.lib.oracle_ref_frame_run
    fn.argtypes = [C.POINTER(OracleFrame)]
    fn.restype = None
    fn(C.byref(of))
    return dst_planes


class OracleReconFrame(C.Structure):
    """oracle/ref_recon.c: input of the checker that runs the reference's own dav1d_recon_b_intra."""
    _fields_ = [("dst", C.c_void_p * 3), ("dst_stride", C.c_ssize_t * 3),
                ("w", C.c_int32), ("h", C.c_int32), ("ss_hor", C.c_int32), ("ss_ver", C.c_int32),
                ("bitdepth_max", C.c_int32), ("no_chroma", C.c_int32), ("intra_edge_filter", C.c_int32),
                ("blocks", C.c_void_p), ("n_blocks", C.c_int32),
                ("ops", C.c_void_p), ("cf", C.c_void_p), ("pal", C.c_void_p), ("pal_idx", C.c_void_p),
                ("ref", (C.c_void_p * 3) * 7), ("ref_stride", (C.c_ssize_t * 2) * 7), ("n_refs", C.c_int32),
                ("tx_recs", C.c_void_p), ("ref_w", C.c_int32 * 7), ("ref_h", C.c_int32 * 7)]


def run_reference_driver(ref, hf, dst_planes, ref_planes_list=()):
    """Frame generated with real_blocks=1 through dav1d_recon_b_intra / dav1d_recon_b_inter_{8,16}bpc, block
    by block in decode order (modifies dst_planes in place)."""
    assert hf.n_block_recs > 0, "generate the frame with real_blocks=1"
    of = OracleReconFrame()
    for pl, a in enumerate(dst_planes):
        of.dst[pl] = a.ctypes.data
        of.dst_stride[pl] = a.strides[0]
    of.w, of.h, of.ss_hor, of.ss_ver = hf.w, hf.h, hf.ss_hor, hf.ss_ver
    of.bitdepth_max, of.no_chroma = hf.bdmax, hf.no_chroma
    of.intra_edge_filter = hf.params.edge_filter
    of.blocks, of.n_blocks = hf.blocks.ctypes.data, hf.n_block_recs
    for name, arr in (("ops", hf.intra), ("cf", hf.cf), ("pal", hf.pal), ("pal_idx", hf.pal_idx)):
        setattr(of, name, arr.ctypes.data if arr.nbytes else None)
    for r, planes in enumerate(ref_planes_list):
        for pl, a in enumerate(planes):
            of.ref[r][pl] = a.ctypes.data
        of.ref_stride[r][0] = planes[0].strides[0]
        of.ref_stride[r][1] = planes[1].strides[0] if len(planes) > 1 else 0
    of.n_refs = len(ref_planes_list)
    for r in range(7):
        of.ref_w[r], of.ref_h[r] = hf.params.ref_w[r], hf.params.ref_h[r]
    of.tx_recs = hf.tx_recs.ctypes.data if hf.tx_recs.nbytes else None
    fn = getattr(ref.lib, "oracle_recon_frame_16bpc" if hf.hbd else "oracle_recon_frame_8bpc")
    fn.argtypes = [C.POINTER(OracleReconFrame)]
    fn.restype = C.c_int
    r = fn(C.byref(of))
    if r:
        raise RuntimeError(f"oracle_recon_frame: {r}")
    return dst_planes


_warp_tab = None


def reference_warp_tab(ref, n=48, seed=7):
    """Valid local-warp models for real-block frames (HostFrame(warp_tab=...)): random affine matrices near the
    identity whose shear parameters are the reference's own (dav1d_get_shear_params, src/warpmv.c:83-106; models
    it rejects are dropped).  The generator fills in the translation per block."""
    global _warp_tab
    if _warp_tab is not None:
        return _warp_tab
    from dav1d_mirror_b200 import frame as F

    class WMP(C.Structure):          # Dav1dWarpedMotionParams (include/dav1d/headers.h:85-98)
        _fields_ = [("type", C.c_int), ("matrix", C.c_int32 * 6), ("abcd", C.c_int16 * 4)]
    fn = ref.lib.dav1d_get_shear_params
    fn.argtypes = [C.POINTER(WMP)]
    fn.restype = C.c_int
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        w = WMP()
        w.type = 3                   # DAV1D_WM_TYPE_AFFINE
        span = 1 << int(rng.integers(6, 13))
        m = rng.integers(-span, span + 1, size=4)
        w.matrix[2], w.matrix[3], w.matrix[4], w.matrix[5] = 0x10000 + int(m[0]), int(m[1]), int(m[2]), 0x10000 + int(m[3])
        if fn(C.byref(w)):
            continue
        e = F.SynthWarp()
        for k in range(6):
            e.matrix[k] = w.matrix[k]
        for k in range(4):
            e.abcd[k] = w.abcd[k]
        out.append(e)
    _warp_tab = (F.SynthWarp * n)(*out)
    return _warp_tab


_mask_tab = None


def reference_mask_tab(ref):
    """The decoder's wedge / inter-intra mask tables (reference: dav1d_masks, src/wedge.c) as the generator's
    D1SynthMaskTab - for real-block frames with wedge compounds or inter-intra blocks (HostFrame(mask_tab=...))."""
    global _mask_tab
    if _mask_tab is not None:
        return _mask_tab[0]
    from dav1d_mirror_b200 import frame as F
    L = ref.lib
    L.oracle_wedge_mask.restype = C.c_void_p
    L.oracle_wedge_mask.argtypes = [C.c_int] * 5
    L.oracle_ii_mask.restype = C.c_void_p
    L.oracle_ii_mask.argtypes = [C.c_int] * 4
    tab = F.SynthMaskTab()
    blob = bytearray()
    for lay in range(3):
        for wi, w4 in enumerate((2, 4, 8)):
            for hi, h4 in enumerate((2, 4, 8)):
                n = (w4 * 4 >> (lay >= 1)) * (h4 * 4 >> (lay == 2))
                for sign in range(2):
                    for idx in range(16):
                        p = L.oracle_wedge_mask(lay, w4, h4, sign, idx)
                        assert p
                        tab.wedge[lay][wi][hi][sign][idx] = len(blob)
                        blob += C.string_at(p, n)
                for mode in range(4):
                    p = L.oracle_ii_mask(lay, w4, h4, mode)
                    assert p
                    tab.ii[lay][wi][hi][mode] = len(blob)
                    blob += C.string_at(p, n)
    buf = np.frombuffer(bytes(blob), dtype=np.uint8).copy()
    tab.base = buf.ctypes.data
    _mask_tab = (tab, buf)
    return tab
