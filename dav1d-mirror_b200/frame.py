"""Frame-level host glue over the C ABI: synthetic descriptor source
(libd1synth.so), level scheduling, upload of descriptor arrays into HBM and
submission of one frame's reconstruction batch.  Used by tests/ and bench.py;
all device work goes through libdav1d_cuda.so."""
import ctypes as C
import os

import numpy as np

from . import binding as B

HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH_PATH = os.path.join(HERE, "libd1synth.so")


class SynthParams(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("ss_hor", C.c_int32), ("ss_ver", C.c_int32),
                ("bitdepth_max", C.c_int32), ("no_chroma", C.c_int32), ("seed", C.c_uint64),
                ("p_intra", C.c_float), ("p_residual", C.c_float), ("p_tx_split", C.c_float),
                ("p_filter_intra", C.c_float), ("p_palette", C.c_float), ("p_cfl", C.c_float),
                ("p_avg", C.c_float), ("p_w_avg", C.c_float), ("p_wedge", C.c_float),
                ("p_seg", C.c_float), ("p_warp", C.c_float),
                ("mv_range", C.c_int32), ("n_refs", C.c_int32), ("edge_filter", C.c_int32),
                ("only_tx", C.c_int32), ("only_txtp", C.c_int32), ("eob_class", C.c_int32),
                ("dense_coefs", C.c_int32), ("p_obmc", C.c_float), ("p_ii", C.c_float), ("p_ibc", C.c_float),
                ("tile_cols", C.c_int32), ("tile_rows", C.c_int32), ("real_blocks", C.c_int32),
                ("ref_w", C.c_int32 * 7), ("ref_h", C.c_int32 * 7), ("mask_tab", C.c_uint64),
                ("warp_tab", C.c_uint64), ("n_warp_tab", C.c_int32), ("p_sub8x8", C.c_float)]


BLOCK_REC_BYTES = 120          # sizeof(D1SynthBlock)


class SynthWarp(C.Structure):
    """D1SynthWarp of csrc/synth.cpp: a local-warp model (Dav1dWarpedMotionParams: matrix + shear parameters)."""
    _fields_ = [("matrix", C.c_int32 * 6), ("abcd", C.c_int16 * 4)]


class SynthMaskTab(C.Structure):
    """D1SynthMaskTab of csrc/synth.cpp: the decoder's wedge / inter-intra mask tables for real-block frames
    (filled by the tests from the reference's tables; offsets into `base`)."""
    _fields_ = [("base", C.c_void_p), ("wedge", ((((C.c_uint32 * 16) * 2) * 3) * 3) * 3),
                ("ii", (((C.c_uint32 * 4) * 3) * 3) * 3)]


class SynthFrame(C.Structure):
    _fields_ = [("mc_put", C.c_void_p), ("n_mc_put", C.c_int32), ("mc_put_tiles", C.c_void_p),
                ("n_mc_put_tiles", C.c_int32),
                ("mc_comp", C.c_void_p), ("n_mc_comp", C.c_int32), ("mc_comp_tiles", C.c_void_p),
                ("n_mc_comp_tiles", C.c_int32 * 2),
                ("n_mc_put_small", C.c_int32), ("n_mc_comp_small", C.c_int32 * 2),
                ("warp", C.c_void_p), ("n_warp", C.c_int32),
                ("itx", C.c_void_p), ("n_itx", C.c_int32), ("itx_class_count", C.c_int32 * 19),
                ("intra", C.c_void_p), ("n_intra", C.c_int32),
                ("cf", C.c_void_p), ("cf_elems", C.c_uint64),
                ("masks", C.c_void_p), ("masks_bytes", C.c_uint64),
                ("pal", C.c_void_p), ("pal_px", C.c_uint64),
                ("pal_idx", C.c_void_p), ("pal_idx_bytes", C.c_uint64),
                ("order", C.c_void_p), ("n_order", C.c_int32),
                ("bw4", C.c_int32), ("bh4", C.c_int32),
                ("algo_bytes", C.c_double), ("algo_class", C.c_double * 5), ("luma_px", C.c_double),
                ("n_blocks", C.c_int64), ("n_intra_blocks", C.c_int64),
                ("mc_obmc", C.c_void_p), ("n_mc_obmc", C.c_int32), ("mc_obmc_tiles", C.c_void_p),
                ("n_mc_obmc_tiles", C.c_int32 * 2),
                ("intra_itx", C.c_void_p), ("n_intra_itx", C.c_int32), ("intra_itx_class_count", C.c_int32 * 19),
                ("dense_coef_bytes", C.c_double), ("blocks", C.c_void_p), ("n_block_recs", C.c_int32),
                ("tx_recs", C.c_void_p), ("n_tx_recs", C.c_int32),
                ("mc_scaled", C.c_void_p), ("n_mc_scaled", C.c_int32 * 4)]


_synth = None


def synth_lib():
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_PATH):
            raise RuntimeError(f"{SYNTH_PATH} missing: run `make -C dav1d-mirror_b200`")
        _synth = C.CDLL(SYNTH_PATH)
        _synth.d1synth_default_params.argtypes = [C.POINTER(SynthParams), C.c_int, C.c_int, C.c_int, C.c_uint64]
        _synth.d1synth_generate.argtypes = [C.POINTER(SynthParams), C.POINTER(SynthFrame)]
        _synth.d1synth_free.argtypes = [C.POINTER(SynthFrame)]
    return _synth


def _np_from(ptr, nbytes):
    if not nbytes:
        return np.zeros(0, dtype=np.uint8)
    return np.frombuffer((C.c_char * nbytes).from_address(ptr), dtype=np.uint8).copy()


class HostFrame:
    """One synthetic frame on the host: descriptor arrays (numpy byte buffers in
    the C ABI's struct layouts), coefficient stream, pools, decode order."""

    def __init__(self, w, h, bitdepth_max, seed, **kw):
        S = synth_lib()
        p = SynthParams()
        S.d1synth_default_params(C.byref(p), w, h, bitdepth_max, seed)
        for k, v in kw.items():
            if not hasattr(p, k):
                raise KeyError(k)
            if k == "mask_tab":                  # a SynthMaskTab (kept alive by the caller)
                p.mask_tab = C.addressof(v)
                continue
            if k == "warp_tab":                  # a ctypes array of SynthWarp (kept alive by the caller)
                p.warp_tab, p.n_warp_tab = C.addressof(v), len(v)
                continue
            if k in ("ref_w", "ref_h"):          # luma size per reference, 0 = the frame's
                for i, x in enumerate(v):
                    getattr(p, k)[i] = x
                continue
            setattr(p, k, v)
        f = SynthFrame()
        r = S.d1synth_generate(C.byref(p), C.byref(f))
        if r:
            raise RuntimeError(f"d1synth_generate failed: {r}")
        self.params = p
        self.w, self.h, self.bdmax = w, h, bitdepth_max
        self.ss_hor, self.ss_ver, self.no_chroma = p.ss_hor, p.ss_ver, p.no_chroma
        self.hbd = bitdepth_max > 0xff
        self.bw4, self.bh4 = f.bw4, f.bh4
        self.mc_put = _np_from(f.mc_put, f.n_mc_put * C.sizeof(B.McDesc))
        self.mc_put_tiles = _np_from(f.mc_put_tiles, f.n_mc_put_tiles * 4)
        self.mc_comp = _np_from(f.mc_comp, f.n_mc_comp * C.sizeof(B.McDesc))
        ntc = f.n_mc_comp_tiles[0] + f.n_mc_comp_tiles[1]
        self.mc_comp_tiles = _np_from(f.mc_comp_tiles, ntc * 4)
        self.n_mc_put_tiles = f.n_mc_put_tiles
        self.n_mc_comp_tiles = (f.n_mc_comp_tiles[0], f.n_mc_comp_tiles[1])
        self.n_mc_put_small = f.n_mc_put_small
        self.n_mc_comp_small = (f.n_mc_comp_small[0], f.n_mc_comp_small[1])
        self.mc_obmc = _np_from(f.mc_obmc, f.n_mc_obmc * C.sizeof(B.McDesc))
        self.n_mc_obmc_tiles = (f.n_mc_obmc_tiles[0], f.n_mc_obmc_tiles[1])
        self.mc_obmc_tiles = _np_from(f.mc_obmc_tiles, (f.n_mc_obmc_tiles[0] + f.n_mc_obmc_tiles[1]) * 4)
        self.n_mc_scaled = tuple(f.n_mc_scaled[i] for i in range(4))
        self.mc_scaled = _np_from(f.mc_scaled, sum(self.n_mc_scaled) * C.sizeof(B.McScaledDesc))
        self.warp = _np_from(f.warp, f.n_warp * C.sizeof(B.WarpDesc))
        self.n_warp = f.n_warp
        self.itx = _np_from(f.itx, f.n_itx * C.sizeof(B.ItxDesc))
        self.itx_class_count = [f.itx_class_count[i] for i in range(19)]
        self.intra = _np_from(f.intra, f.n_intra * C.sizeof(B.IntraDesc))   # decode order
        self.n_intra = f.n_intra
        self.cf = _np_from(f.cf, f.cf_elems * (4 if self.hbd else 2))
        self.masks = _np_from(f.masks, f.masks_bytes)
        self.pal = _np_from(f.pal, f.pal_px * (2 if self.hbd else 1))
        self.pal_idx = _np_from(f.pal_idx, f.pal_idx_bytes)
        self.order = _np_from(f.order, f.n_order * 4)
        self.algo_bytes, self.luma_px = f.algo_bytes, f.luma_px
        self.algo_class = dict(zip(("mc_put", "mc_compound", "warp", "itx", "intra"), list(f.algo_class)))
        self.n_blocks, self.n_intra_blocks = f.n_blocks, f.n_intra_blocks
        self.dense_coef_bytes = f.dense_coef_bytes
        # real_blocks: one record per coded block (the Av1Block fields the reference's driver reads), 88 bytes each
        self.n_block_recs = f.n_block_recs
        self.blocks = _np_from(f.blocks, f.n_block_recs * BLOCK_REC_BYTES)
        self.tx_recs = _np_from(f.tx_recs, f.n_tx_recs * 12)          # cbi / cf entries of the inter blocks
        # intra-class operations stay in decode order; their residuals are listed a second time as
        # transform descriptors ordered like `itx`
        self.intra_itx = _np_from(f.intra_itx, f.n_intra_itx * C.sizeof(B.ItxDesc))
        self.intra_itx_class_count = [f.intra_itx_class_count[i] for i in range(19)]
        S.d1synth_free(C.byref(f))
        # inter residuals: task codes over the (size, type)-sorted itx array (linear pass)
        L = B.lib()
        n_itx = self.itx.nbytes // C.sizeof(B.ItxDesc)
        tasks = np.zeros(max(n_itx, 1), dtype=np.uint32)
        ns, nb = C.c_int32(), C.c_int32()
        k = L.dav1d_cuda_itx_tasks(self.itx.ctypes.data, n_itx, 0, tasks.ctypes.data, C.byref(ns), C.byref(nb)) \
            if n_itx else 0
        self.itx_tasks = tasks[:max(k, 1)].copy().view(np.uint8)
        self.n_itx_tasks = (ns.value, nb.value)
        n_iitx = self.intra_itx.nbytes // C.sizeof(B.ItxDesc)
        tasks = np.zeros(max(n_iitx, 1), dtype=np.uint32)
        k = L.dav1d_cuda_itx_tasks(self.intra_itx.ctypes.data, n_iitx, 0, tasks.ctypes.data, C.byref(ns), C.byref(nb)) \
            if n_iitx else 0
        self.intra_itx_tasks = tasks[:max(k, 1)].copy().view(np.uint8)
        self.n_intra_itx_tasks = (ns.value, nb.value) if n_iitx else (0, 0)
        self.n_levels = None
        self.cf16 = None                  # pack_coefs()

    def record_levels(self, into=None):
        """The recorder's linear pass over the intra-class descriptors (decode order): dependency
        level + 1 into `reserved`.  into: address of another copy of the descriptor array (a pinned
        mirror) to annotate instead of self.intra."""
        if not self.n_intra:
            return 0
        ssh = 0 if self.no_chroma else self.ss_hor
        ssv = 0 if self.no_chroma else self.ss_ver
        r = B.lib().dav1d_cuda_intra_levels(into or self.intra.ctypes.data, self.n_intra, self.bw4, self.bh4, ssh, ssv)
        if r < 0:
            raise RuntimeError(f"dav1d_cuda_intra_levels: {r}")
        if not into:
            self.n_levels = r
        return r

    def pack_coefs(self):
        """High bit depth: the compact coefficient stream (Dav1dCudaReconBatch.cf_int16) - int16 storage plus an
        escape list for what does not fit.  self.cf stays the int32 stream (what the oracle reads); the device
        arena takes cf16 / cf_esc instead."""
        assert self.hbd
        n = self.cf.nbytes // 4
        self.cf16 = np.zeros(max(n, 1), dtype=np.int16)
        esc = np.zeros((max(n, 1), 2), dtype=np.int32)
        k = B.lib().dav1d_cuda_pack_coefs(self.cf.ctypes.data, n, self.cf16.ctypes.data, esc.ctypes.data, len(esc))
        if k < 0:
            raise RuntimeError(f"dav1d_cuda_pack_coefs: {k}")
        self.cf16 = self.cf16.view(np.uint8)
        self.n_cf_esc = k
        self.cf_esc = esc[:max(k, 1)].copy().view(np.uint8).reshape(-1)
        return k

    def plane_shape(self, pl, ref=None):
        """ref: index of a reference picture (its size may differ from the frame's: params.ref_w / ref_h)."""
        sh = self.ss_hor if pl else 0
        sv = self.ss_ver if pl else 0
        w, h = self.ref_size(ref) if ref is not None else (self.w, self.h)
        return ((h + sv) >> sv, (w + sh) >> sh)

    def ref_size(self, r):
        return (self.params.ref_w[r] or self.w, self.params.ref_h[r] or self.h)

    def host_bytes(self):
        """Bytes a decoder ships host->device for this frame (descriptors + coefficients + pools)."""
        cf = self.cf if self.cf16 is None else np.concatenate([self.cf16, self.cf_esc[:self.n_cf_esc * 8]])
        return sum(a.nbytes for a in (self.mc_put, self.mc_put_tiles, self.mc_comp, self.mc_comp_tiles, self.warp,
                                      self.mc_obmc, self.mc_obmc_tiles, self.mc_scaled, self.itx, self.itx_tasks, cf, self.masks,
                                      self.pal, self.pal_idx, self.intra, self.intra_itx,
                                      self.intra_itx_tasks))


def random_planes(hf, seed, ref=None):
    """Reference / initial picture content: uniform random pixels (checkasm's worst case)."""
    rng = np.random.default_rng(seed)
    dt = np.uint16 if hf.hbd else np.uint8
    return [rng.integers(0, hf.bdmax + 1, size=hf.plane_shape(pl, ref), dtype=np.int32).astype(dt)
            for pl in range(1 if hf.no_chroma else 3)]


class DeviceFrame:
    """Device-resident state of one stream: destination + reference pictures, the cell map of the
    intra executor and ONE descriptor arena that receives, frame after frame, the descriptor set
    of the frame to reconstruct (`hf`, or - use(k) - one of several sets of the same geometry)."""

    def __init__(self, ctx, hf, n_refs=2, tasks=True, more_sets=()):
        """tasks: explicit transform task codes (else the implicit per-size runs)."""
        self.L = B.lib()
        self.ctx = ctx
        L = self.L
        self.dst = B.Picture()
        self.refs = [B.Picture() for _ in range(n_refs)]
        for k, pic in enumerate([self.dst] + self.refs):
            pw, ph = hf.ref_size(k - 1) if k else (hf.w, hf.h)
            r = L.dav1d_cuda_picture_alloc(ctx, C.byref(pic), pw, ph, hf.ss_hor, hf.ss_ver, hf.bdmax)
            if r:
                raise RuntimeError("dav1d_cuda_picture_alloc failed")
        ssh = 0 if hf.no_chroma else hf.ss_hor
        ssv = 0 if hf.no_chroma else hf.ss_ver
        # int16 residual planes of the intra pre-pass: a picture of the same geometry with 16-bit samples
        self.res = B.Picture()
        if L.dav1d_cuda_picture_alloc(ctx, C.byref(self.res), hf.w, hf.h, hf.ss_hor, hf.ss_ver, 0xffff):
            raise RuntimeError("dav1d_cuda_picture_alloc failed")
        self._cellmap_bytes = L.dav1d_cuda_intra_cellmap_bytes(hf.bw4, hf.bh4, ssh, ssv)
        self._cellmap = L.dav1d_cuda_malloc(self._cellmap_bytes)              # cleared by every submission
        # one device arena for descriptors, task lists, coefficients and pools (each array 256-byte
        # aligned): the end-to-end path ships a frame's set with ONE host->device copy from a pinned
        # mirror (every extra copy costs ~10 us of copy-engine time, tools/exp_copy.py)
        self._sets = []
        for h in [hf] + list(more_sets):
            assert (h.w, h.h, h.bdmax, h.ss_hor, h.ss_ver) == (hf.w, hf.h, hf.bdmax, hf.ss_hor, hf.ss_ver)
            self._sets.append(self._layout(h))
        self._arena_cap = max(st["bytes"] for st in self._sets)
        self._arena = L.dav1d_cuda_malloc(self._arena_cap)
        if not self._arena:
            raise RuntimeError("dav1d_cuda_malloc failed")
        for st in self._sets:
            st["batch"] = self._make_batch(st, n_refs, tasks)
        self.graph = None
        self._pinned = None
        self.use(0)

    def _layout(self, hf):
        names = ["mc_put", "mc_put_tiles", "mc_comp", "mc_comp_tiles", "warp", "itx", "intra",
                 "cf", "masks", "pal", "pal_idx", "itx_tasks", "intra_itx", "intra_itx_tasks"]
        alt = {}
        if hf.mc_obmc.nbytes:
            names += ["mc_obmc", "mc_obmc_tiles"]
        if hf.mc_scaled.nbytes:
            names += ["mc_scaled"]
        if hf.cf16 is not None:
            alt["cf"] = hf.cf16
            alt["cf_esc"] = hf.cf_esc
            names += ["cf_esc"]
        host, offs, off = {}, {}, 0
        for name in names:
            arr = alt.get(name, getattr(hf, name))
            host[name] = arr
            offs[name] = off
            off += (max(arr.nbytes, 256) + 255) & ~255
        return {"hf": hf, "host": host, "off": offs, "bytes": off, "pinned": None}

    def _make_batch(self, st, n_refs, tasks):
        hf = st["hf"]
        d = {name: self._arena + o for name, o in st["off"].items()}
        b = B.ReconBatch()
        b.dst = C.pointer(self.dst)
        for i in range(7):
            b.refs[i] = C.pointer(self.refs[i]) if i < n_refs else None
        b.bw4, b.bh4 = hf.bw4, hf.bh4
        b.cf, b.masks, b.pal, b.pal_idx = d["cf"], d["masks"], d["pal"], d["pal_idx"]
        b.mc_put, b.mc_put_tiles, b.n_mc_put_tiles = d["mc_put"], d["mc_put_tiles"], hf.n_mc_put_tiles
        b.mc_comp, b.mc_comp_tiles = d["mc_comp"], d["mc_comp_tiles"]
        b.n_mc_comp_tiles[0], b.n_mc_comp_tiles[1] = hf.n_mc_comp_tiles
        b.n_mc_put_small = hf.n_mc_put_small
        b.n_mc_comp_small[0], b.n_mc_comp_small[1] = hf.n_mc_comp_small
        b.warp, b.n_warp = d["warp"], hf.n_warp
        if hf.mc_obmc.nbytes:
            b.mc_obmc, b.mc_obmc_tiles = d["mc_obmc"], d["mc_obmc_tiles"]
            b.n_mc_obmc_tiles[0], b.n_mc_obmc_tiles[1] = hf.n_mc_obmc_tiles
        if hf.cf16 is not None:
            b.cf_int16, b.cf_esc, b.n_cf_esc = 1, d["cf_esc"], hf.n_cf_esc
        if hf.mc_scaled.nbytes:
            b.mc_scaled = d["mc_scaled"]
            for i in range(4):
                b.n_mc_scaled[i] = hf.n_mc_scaled[i]
        b.itx = d["itx"]
        for i in range(19):
            b.itx_class_count[i] = hf.itx_class_count[i]
        if tasks:
            b.itx_tasks = d["itx_tasks"]
            b.n_itx_tasks[0], b.n_itx_tasks[1] = hf.n_itx_tasks
        b.intra, b.n_intra = d["intra"], hf.n_intra
        b.intra_cellmap = self._cellmap
        b.intra_itx = d["intra_itx"]
        for i in range(19):
            b.intra_itx_class_count[i] = hf.intra_itx_class_count[i]
        if tasks:
            b.intra_itx_tasks = d["intra_itx_tasks"]
            b.n_intra_itx_tasks[0], b.n_intra_itx_tasks[1] = hf.n_intra_itx_tasks
        b.intra_res = C.pointer(self.res)
        b.intra_levels_recorded = 1 if hf.n_levels else 0
        return b

    def set_levels_recorded(self, on):
        """Tell the library whether the descriptors' recorder-assigned dependency levels are to be
        used (HostFrame.record_levels() ran) or the device works the levels out itself."""
        for st in self._sets:
            st["batch"].intra_levels_recorded = 1 if (on and st["hf"].n_levels) else 0

    def use(self, k):
        """Make descriptor set k the frame to reconstruct next (its arrays still have to be shipped)."""
        st = self._sets[k % len(self._sets)]
        self._cur = st
        self.hf, self._host, self._off = st["hf"], st["host"], st["off"]
        self._dev = {name: self._arena + o for name, o in st["off"].items()}
        self.arena_bytes, self.batch = st["bytes"], st["batch"]

    def upload_descriptors(self):
        for name, arr in self._host.items():
            if arr.nbytes:
                self.L.dav1d_cuda_upload(self.ctx, self._dev[name], arr.ctypes.data, arr.nbytes)

    def cellmap_is_clear(self):
        """The count part of the cell map (one byte per 4x4 cell of the three planes, rows padded to
        a word; the 16-bit level map follows it) is back at zero once a frame is complete."""
        hf = self.hf
        cells = 0
        for pl in range(3):
            w4 = (hf.bw4 + self.dst.ss_hor) >> self.dst.ss_hor if pl else hf.bw4
            h4 = (hf.bh4 + self.dst.ss_ver) >> self.dst.ss_ver if pl else hf.bh4
            cells += ((w4 + 3) & ~3) * h4
        buf = np.ones(cells, dtype=np.uint8)
        self.L.dav1d_cuda_download(self.ctx, buf.ctypes.data, self._cellmap, cells)
        self.L.dav1d_cuda_synchronize(self.ctx)
        return not buf.any()

    def upload_picture(self, pic, planes):
        for pl, a in enumerate(planes):
            self.L.dav1d_cuda_picture_upload(self.ctx, C.byref(pic), pl, a.ctypes.data, a.strides[0])

    def download_picture(self, pic=None):
        pic = pic or self.dst
        hf = self.hf
        out = []
        for pl in range(1 if hf.no_chroma else 3):
            a = np.zeros(hf.plane_shape(pl), dtype=np.uint16 if hf.hbd else np.uint8)
            self.L.dav1d_cuda_picture_download(self.ctx, C.byref(pic), pl, a.ctypes.data, a.strides[0])
            out.append(a)
        self.L.dav1d_cuda_synchronize(self.ctx)
        return out

    # ---- end-to-end path: host buffers in pinned memory
    def alloc_pinned(self, share=None):
        """share: a dict that keeps ONE pinned mirror per descriptor set for all the streams that
        decode it (the bench's streams cycle through the same few sets); such mirrors are not freed
        by close()."""
        if self._pinned:
            return
        L = self.L
        # pinned mirrors of the descriptor sets
        self._own_pinned = share is None
        for st in self._sets:
            key = id(st["hf"])
            if share is not None and key in share:
                st["pinned"] = share[key]
                continue
            st["pinned"] = L.dav1d_cuda_host_alloc(st["bytes"])
            for name, arr in st["host"].items():
                if arr.nbytes:
                    C.memmove(st["pinned"] + st["off"][name], arr.ctypes.data, arr.nbytes)
            if share is not None:
                share[key] = st["pinned"]
        self._pinned = True
        # pinned host frame with the device picture's layout (plane offsets and strides): the
        # reconstructed frame comes back with ONE device->host copy
        hf = self.hf
        npl = 1 if hf.no_chroma else 3
        bpp = 2 if hf.hbd else 1
        base = self.dst.p[0].data
        last = self.dst.p[npl - 1]
        self.pinned_out_bytes = (last.data - base) + (last.h - 1) * last.stride + last.w * bpp
        self._pinned_out = L.dav1d_cuda_host_alloc(self.pinned_out_bytes)

    def upload_descriptors_pinned(self):
        self.L.dav1d_cuda_upload(self.ctx, self._arena, self._cur["pinned"], self.arena_bytes)

    def download_pinned(self):
        self.L.dav1d_cuda_download(self.ctx, self._pinned_out, self.dst.p[0].data, self.pinned_out_bytes)

    def pinned_planes(self):
        """numpy views of the last downloaded frame (after a synchronize)."""
        hf = self.hf
        out = []
        bpp = 2 if hf.hbd else 1
        base = self.dst.p[0].data
        for pl in range(1 if hf.no_chroma else 3):
            hh, ww = hf.plane_shape(pl)
            p = self.dst.p[pl]
            a = np.frombuffer((C.c_char * (p.stride * hh)).from_address(self._pinned_out + (p.data - base)),
                              dtype=np.uint16 if hf.hbd else np.uint8)
            out.append(a.reshape(hh, p.stride // bpp)[:, :ww])
        return out

    # ---- per-launch-class timing (CUDA events on this context's stream)
    def run_class(self, name):
        """Launch one launch class of this frame exactly as the submit does (async on the context's stream)."""
        bit = {"mc_put": 1, "mc_compound": 2, "warp": 4, "itx": 8, "intra": 16}[name]
        r = self.L.dav1d_cuda_recon_submit_phases(self.ctx, C.byref(self.batch), bit)
        if r:
            raise RuntimeError(f"dav1d_cuda_recon_submit_phases: {r}")

    def time_classes(self, reps=5, flush_mb=256):
        return time_classes([self], reps=reps, flush_mb=flush_mb)

    def submit(self):
        r = self.L.dav1d_cuda_recon_submit(self.ctx, C.byref(self.batch))
        if r:
            raise RuntimeError(f"dav1d_cuda_recon_submit: {r}")

    def build_graph(self):
        g = C.c_void_p()
        n = self.L.dav1d_cuda_recon_graph_build(self.ctx, C.byref(self.batch), C.byref(g))
        if n < 0:
            raise RuntimeError(f"dav1d_cuda_recon_graph_build: {n}")
        self.graph = g
        self.graph_nodes = n
        return n

    def launch_graph(self):
        r = self.L.dav1d_cuda_recon_graph_launch(self.ctx, self.graph)
        if r:
            raise RuntimeError(f"dav1d_cuda_recon_graph_launch: {r}")

    def close(self):
        L = self.L
        L.dav1d_cuda_synchronize(self.ctx)
        if self.graph:
            L.dav1d_cuda_recon_graph_free(self.graph)
            self.graph = None
        if self._arena:
            L.dav1d_cuda_free(self._arena)
            self._arena = None
        L.dav1d_cuda_free(self._cellmap)
        self._dev = {}
        if self._pinned:
            if self._own_pinned:
                for st in self._sets:
                    L.dav1d_cuda_host_free(st["pinned"])
            L.dav1d_cuda_host_free(self._pinned_out)
        self._pinned, self._pinned_out = None, None
        for pic in [self.dst, self.res] + self.refs:
            L.dav1d_cuda_picture_free(self.ctx, C.byref(pic))


CLASSES = ("mc_put", "mc_compound", "warp", "itx", "intra")


def time_classes(dfs, reps=3, flush_mb=0):
    """Average duration (ms) of every launch class of ONE frame, measured with CUDA events on the
    launching stream.  With several frames the launches rotate over them, so the data of a frame has
    left the L2 (working set of all frames >> L2) by the time it is touched again while the code stays
    warm; with a single frame an explicit L2 flush (memset) is issued between repetitions."""
    L = B.lib()
    out = {}
    flush = L.dav1d_cuda_malloc(flush_mb << 20) if flush_mb else None
    evs = [(L.dav1d_cuda_event_create(), L.dav1d_cuda_event_create()) for _ in dfs]
    for name in CLASSES:
        tot, n = 0.0, 0
        for rep in range(reps + 1):
            for df, (e0, e1) in zip(dfs, evs):
                if flush:
                    L.dav1d_cuda_memset(df.ctx, flush, 0, flush_mb << 20)
                L.dav1d_cuda_event_record(df.ctx, e0)
                df.run_class(name)
                L.dav1d_cuda_event_record(df.ctx, e1)
                ms = L.dav1d_cuda_event_elapsed_ms(e0, e1)
                if rep:                      # first round = warm-up
                    tot += ms
                    n += 1
        out[name] = tot / max(n, 1)
    for e0, e1 in evs:
        L.dav1d_cuda_event_destroy(e0)
        L.dav1d_cuda_event_destroy(e1)
    if flush:
        L.dav1d_cuda_free(flush)
    return out


class MultiFrame:
    """Frames of several independent streams submitted together (dav1d_cuda_recon_group_submit):
    one intra executor launch for the whole group.  graph=True captures the launches once
    (dav1d_cuda_recon_graph_build_multi) and replays them."""

    def __init__(self, ctx, dfs, phase_mask=31, graph=False):
        self.L = B.lib()
        self.ctx = ctx
        self.dfs = dfs
        self.phase_mask = phase_mask
        self.arr = (C.POINTER(B.ReconBatch) * len(dfs))(*[C.pointer(df.batch) for df in dfs])
        self.graph = None
        if graph:
            g = C.c_void_p()
            n = self.L.dav1d_cuda_recon_graph_build_multi_phases(ctx, self.arr, len(dfs), phase_mask, C.byref(g))
            if n < 0:
                raise RuntimeError(f"dav1d_cuda_recon_graph_build_multi: {n}")
            self.graph, self.graph_nodes = g, n

    def launch(self):
        if self.graph:
            r = self.L.dav1d_cuda_recon_graph_launch(self.ctx, self.graph)
        else:
            r = self.L.dav1d_cuda_recon_group_submit_phases(self.ctx, self.arr, len(self.dfs), self.phase_mask)
        if r:
            raise RuntimeError(f"group submission failed: {r}")

    def close(self):
        if self.graph:
            self.L.dav1d_cuda_recon_graph_free(self.graph)
            self.graph = None


def open_context(device=0, stream=None):
    L = B.lib()
    ctx = C.c_void_p()
    r = L.dav1d_cuda_open(C.byref(ctx), device, stream)
    if r:
        B.check_error()
        raise RuntimeError(f"dav1d_cuda_open: {r}")
    return ctx
