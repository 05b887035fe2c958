"""The product's recorder - dav1d_cuda_record_b_intra(), the transcription of dav1d_recon_b_intra()
(src/recon_tmpl.c:1195-1596) into descriptor emission - run over the Av1Block-style records of synthetic
all-intra frames: the descriptors it emits are, field for field, the ones the generator recorded
independently, and those reproduce the reference driver's pixels bit for bit
(tests/test_reference_driver.py).  CPU only."""
import ctypes as C

import numpy as np
import pytest

import _d1pkg

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import binding as B  # noqa: E402
from dav1d_mirror_b200 import frame as F  # noqa: E402
import test_reference_driver as R  # noqa: E402

BLK = np.dtype([("bx4", "u2"), ("by4", "u2"), ("w4", "u1"), ("h4", "u1"), ("intra", "u1"), ("has_chroma", "u1"),
                ("skip", "u1"), ("tile", "u1"), ("edge_tr", "u1"), ("edge_bl", "u1"), ("y_mode", "u1"), ("uv_mode", "u1"),
                ("y_angle", "i1"), ("uv_angle", "i1"), ("tx", "u1"), ("uvtx", "u1"), ("pal_sz", "u1", 2),
                ("cfl_alpha", "i1", 2), ("tile_rect", "u2", 4), ("pad0", "u2"), ("pal_off", "u4", 3),
                ("pal_idx_off", "u4", 2), ("first_op", "u4"), ("n_ops", "u4"), ("sm_flags", "u1"), ("pad", "u1", 3),
                ("mvx", "i2", 2), ("mvy", "i2", 2), ("ref", "u1", 2), ("comp_kind", "u1"), ("filter2d", "u1"),
                ("mask_sign", "u1"), ("max_ytx", "u1"), ("tx_split", "u1"), ("jnt_weight", "u1"), ("first_tx", "u4"),
                ("n_tx", "u4"), ("warp_matrix", "i4", 6), ("warp_abcd", "i2", 4)])
OP = np.dtype([("x4", "u2"), ("y4", "u2"), ("tile_x4_start", "u2"), ("tile_y4_start", "u2"), ("tile_x4_end", "u2"),
               ("tile_y4_end", "u2"), ("plane", "u1"), ("tw4", "u1"), ("th4", "u1"), ("mode", "u1"),
               ("angle_delta", "i1"), ("edge_flags", "u1"), ("flags", "u2"), ("eob", "i2"), ("tx", "u1"), ("txtp", "u1"),
               ("coef_off", "u4"), ("aux", "u4"), ("reserved", "u4"), ("cw4", "u1"), ("ch4", "u1"), ("pad", "u2")])


def record_intra_block(L, r, s, ops):
    """One intra block record through dav1d_cuda_record_b_intra into the recorder `r`."""
    r.tile_col_start, r.tile_row_start, r.tile_col_end, r.tile_row_end = (int(v) for v in s["tile_rect"])
    b = B.BlockIntra()
    b.bx4, b.by4, b.bw4, b.bh4 = int(s["bx4"]), int(s["by4"]), int(s["w4"]), int(s["h4"])
    b.y_mode, b.uv_mode, b.y_angle, b.uv_angle = int(s["y_mode"]), int(s["uv_mode"]), int(s["y_angle"]), int(s["uv_angle"])
    b.tx, b.uvtx, b.skip, b.sm_flags = int(s["tx"]), int(s["uvtx"]), int(s["skip"]), int(s["sm_flags"])
    for k in range(2):
        b.pal_sz[k], b.cfl_alpha[k], b.pal_idx_off[k] = int(s["pal_sz"][k]), int(s["cfl_alpha"][k]), int(s["pal_idx_off"][k])
    for k in range(3):
        b.pal_off[k] = int(s["pal_off"][k])
    # block-level EdgeFlags (src/intra_edge.h:27-32): bit 0 of the record = luma, bit 1 = the chroma layout
    b.edge_flags = ((1 if s["edge_tr"] & 1 else 0) | (8 if s["edge_bl"] & 1 else 0) |
                    (6 if s["edge_tr"] & 2 else 0) | (48 if s["edge_bl"] & 2 else 0))
    mine = ops[s["first_op"]:s["first_op"] + s["n_ops"]]
    # the block's cbi / cf entries in consumption order: every transform block of a non-skip block
    txs = [(int(o["coef_off"]), int(o["eob"]), int(o["txtp"]), int(o["cw4"]), int(o["ch4"]))
           for o in mine if o["mode"] != 15 and not s["skip"]]
    arr = (B.TxCoef * max(len(txs), 1))()
    for k, (co, eob, txtp, cw4, ch4) in enumerate(txs):
        arr[k].coef_off, arr[k].eob, arr[k].txtp, arr[k].cw4, arr[k].ch4 = co, eob, txtp, cw4, ch4
    n = L.dav1d_cuda_record_b_intra(C.byref(r), C.byref(b), arr, len(txs))
    assert n == s["n_ops"], (n, int(s["n_ops"]), s)


def record_frame(hf):
    assert BLK.itemsize == 120 and OP.itemsize == 40
    L = pkg.lib()
    blocks = np.frombuffer(hf.blocks.tobytes(), dtype=BLK)
    ops = np.frombuffer(hf.intra.tobytes(), dtype=OP)
    out = np.zeros(len(ops) + 16, dtype=OP)
    r = B.Recorder()
    r.bw4, r.bh4 = hf.bw4, hf.bh4
    r.layout = 0 if hf.no_chroma else 1 if hf.ss_ver else 2 if hf.ss_hor else 3
    r.intra_edge_filter = hf.params.edge_filter
    r.intra, r.n_intra, r.cap_intra = out.ctypes.data, 0, len(out)
    for s in blocks:
        record_intra_block(L, r, s, ops)
    return ops, out[:r.n_intra]


@pytest.mark.parametrize("name", [n for n in R.CASES if not n.startswith(("inter_", "obmc_", "scaled_", "wedge_", "ii_", "warp_", "ibc_", "sub8x8_"))])
def test_recorder_emits_the_generators_descriptors(name):
    hf, _ = R.make(name)
    want, got = record_frame(hf)
    assert len(want) == len(got)
    # The top-right / bottom-left bits only reach the predictor through have_top / have_left
    # (ipred_prepare_tmpl.c:139-170); without the neighbour they carry no meaning (the reference sets
    # bottom-left for every transform block above the last row, the generator only where pixels exist).
    want, got = want.copy(), got.copy()
    for a in (want, got):
        a["edge_flags"] &= np.where(a["y4"] > a["tile_y4_start"], 0xff, 0xfe).astype(np.uint8)
        a["edge_flags"] &= np.where(a["x4"] > a["tile_x4_start"], 0xff, 0xf7).astype(np.uint8)
    for f in OP.names:
        if f in ("reserved", "pad"):
            continue
        bad = np.nonzero(want[f] != got[f])[0]
        assert bad.size == 0, f"{name}: field {f} differs at operation {bad[0]}: {want[bad[0]]} vs {got[bad[0]]}"


def test_recorder_rejects_what_it_cannot_record():
    L = pkg.lib()
    out = np.zeros(4, dtype=OP)
    r = B.Recorder()
    r.bw4 = r.bh4 = 16
    r.layout, r.tile_col_end, r.tile_row_end = 1, 16, 16
    r.intra, r.cap_intra = out.ctypes.data, 1
    b = B.BlockIntra()
    b.bw4 = b.bh4 = 4
    b.tx, b.uvtx, b.skip = 2, 1, 1               # 16x16 luma transform, 8x8 chroma: three operations
    assert L.dav1d_cuda_record_b_intra(C.byref(r), C.byref(b), None, 0) == -28 and r.n_intra == 0   # -ENOSPC, nothing kept
    r.cap_intra = 4
    assert L.dav1d_cuda_record_b_intra(C.byref(r), C.byref(b), None, 0) == 3
    b.skip = 0
    assert L.dav1d_cuda_record_b_intra(C.byref(r), C.byref(b), None, 0) == -22                      # cbi entries missing


@pytest.mark.parametrize("name", ["420_10b_cfl_pal_filter", "422_12b_cfl_filter", "420_10b_tiles_2x2", "444_8b_cfl_pal"])
def test_recorded_descriptors_reproduce_the_reference_driver(ref, name):
    """End to end on the CPU: Av1Block records -> product recorder -> descriptors -> sequential replay through the
    reference DSP tables == dav1d_recon_b_intra on the same records."""
    import refframe
    hf, init = R.make(name)
    want = refframe.run_reference_driver(ref, hf, [p.copy() for p in init])
    _, got_ops = record_frame(hf)
    hf.intra = np.frombuffer(got_ops.tobytes(), dtype=np.uint8).copy()
    got = refframe.run_oracle(ref, hf, [p.copy() for p in init], [])
    assert all(np.array_equal(a, b) for a, b in zip(want, got))


# ---------------------------------------------------------------- inter half
MC = np.dtype([("x", "u2"), ("y", "u2"), ("w", "u1"), ("h", "u1"), ("plane", "u1"), ("kind", "u1"),
               ("src", [("x", "i4"), ("y", "i4"), ("ref", "u1"), ("filter_2d", "u1"), ("mx", "u1"), ("my", "u1")], 2),
               ("weight", "u1"), ("mask_ss", "u1"), ("aux16", "u2"), ("aux_off", "u4")])
MCS = np.dtype([("x", "u2"), ("y", "u2"), ("w", "u1"), ("h", "u1"), ("plane", "u1"), ("kind", "u1"),
                ("src", [("pos_x", "i4"), ("pos_y", "i4"), ("step_x", "i4"), ("step_y", "i4"), ("ref", "u1"),
                         ("filter_2d", "u1"), ("pad", "u2")], 2),
                ("weight", "u1"), ("mask_ss", "u1"), ("aux16", "u2"), ("aux_off", "u4")])
ITX = np.dtype([("coef_off", "u4"), ("x", "u2"), ("y", "u2"), ("eob", "i2"), ("plane", "u1"), ("tx", "u1"), ("txtp", "u1"),
                ("cw4", "u1"), ("ch4", "u1"), ("pad", "u1")])
TXR = np.dtype([("coef_off", "u4"), ("eob", "i2"), ("txtp", "u1"), ("cw4", "u1"), ("ch4", "u1"), ("tx", "u1"),
                ("plane", "u1"), ("pad", "u1")])
WARP = np.dtype([("x", "u2"), ("y", "u2"), ("sx", "i4"), ("sy", "i4"), ("mx", "i4"), ("my", "i4"), ("abcd", "i2", 4),
                 ("plane", "u1"), ("ref", "u1"), ("pad", "u2")])
# generator's compound kind (enum Dav1dCudaMcKind; 255 = a warped single-reference block) -> enum CompInterType
COMP_TYPE = {0: 0, 1: 2, 2: 1, 4: 3, 255: 0, 254: 0}        # 254: an intrabc block of a key frame


# generator's compound kind 3 (DAV1D_CUDA_MC_MASK on a real-block frame) = COMP_INTER_WEDGE
COMP_TYPE[3] = 4


def record_inter_frame(hf, mask_tab=None):
    """Every block record of a mixed frame through the product recorder: intra blocks through
    dav1d_cuda_record_b_intra (when the frame has inter-intra blocks, whose intra-class operations interleave
    with theirs in decode order), inter blocks through dav1d_cuda_record_b_inter.  `mask_tab`: the reference's
    wedge / inter-intra tables (refframe.reference_mask_tab) the caller of the recorder picks the masks from."""
    assert MC.itemsize == 40 and MCS.itemsize == 56 and ITX.itemsize == 16 and TXR.itemsize == 12 and WARP.itemsize == 32
    L = pkg.lib()
    blocks = np.frombuffer(hf.blocks.tobytes(), dtype=BLK)
    txr = np.frombuffer(hf.tx_recs.tobytes(), dtype=TXR)
    ops = np.frombuffer(hf.intra.tobytes(), dtype=OP)
    ops_out = np.zeros(len(ops) + 16, dtype=OP)
    ri = B.Recorder()
    ri.bw4, ri.bh4 = hf.bw4, hf.bh4
    ri.layout = 0 if hf.no_chroma else 1 if hf.ss_ver else 2 if hf.ss_hor else 3
    ri.intra_edge_filter = hf.params.edge_filter
    ri.intra, ri.n_intra, ri.cap_intra = ops_out.ctypes.data, 0, len(ops_out)
    masks = np.zeros(max(hf.masks.nbytes, 1) + 64, dtype=np.uint8)
    lay_c = 0 if (hf.no_chroma or not hf.ss_hor) else 2 if hf.ss_ver else 1
    l2 = {2: 0, 4: 1, 8: 2}
    r = B.InterRecorder()
    r.bw4, r.bh4, r.w, r.h = hf.bw4, hf.bh4, hf.w, hf.h
    r.layout = 0 if hf.no_chroma else 1 if hf.ss_ver else 2 if hf.ss_hor else 3
    for i in range(7):
        r.ref_w[i], r.ref_h[i] = hf.params.ref_w[i], hf.params.ref_h[i]
        for j in range(7):
            r.jnt_weights[i][j] = 1 + (i * 7 + j * 3 + 4) % 15         # the frame's table (generator + reference harness)
    above = np.zeros(hf.bw4 + 1, dtype=np.uint64)
    left = np.zeros(hf.bh4 + 1, dtype=np.uint64)
    above.view(np.int8)[4::8] = -1
    left.view(np.int8)[4::8] = -1
    r.above, r.left = above.ctypes.data, left.ctypes.data
    cap = 1 << 15
    out = {"put": np.zeros(cap, dtype=MC), "comp0": np.zeros(cap, dtype=MC), "comp1": np.zeros(cap, dtype=MC),
           "obmc0": np.zeros(cap, dtype=MC), "obmc1": np.zeros(cap, dtype=MC), "itx": np.zeros(cap, dtype=ITX)}
    sc = [np.zeros(cap, dtype=MCS) for _ in range(4)]
    r.put, r.cap_put = out["put"].ctypes.data, cap
    for k in range(2):
        r.comp[k], r.cap_comp[k] = out[f"comp{k}"].ctypes.data, cap
        r.obmc[k], r.cap_obmc[k] = out[f"obmc{k}"].ctypes.data, cap
    for k in range(4):
        r.scaled[k], r.cap_scaled[k] = sc[k].ctypes.data, cap
    r.itx, r.cap_itx = out["itx"].ctypes.data, cap
    r.masks, r.cap_masks = masks.ctypes.data, len(masks)
    r.intra = C.pointer(ri)
    warps = np.zeros(cap, dtype=WARP)
    r.warp, r.cap_warp = warps.ctypes.data, cap
    r.intrabc = int(np.any((blocks["intra"] == 0) & (blocks["comp_kind"] == 254)))      # IS_KEY_OR_INTRA(f->frame_hdr)
    for s in blocks:
        r.tile_col_start, r.tile_row_start = int(s["tile_rect"][0]), int(s["tile_rect"][1])
        if s["intra"]:
            assert L.dav1d_cuda_record_nb_intra(C.byref(r), int(s["bx4"]), int(s["by4"]), int(s["w4"]), int(s["h4"])) == 0
            record_intra_block(L, ri, s, ops)
            continue
        ri.tile_col_start, ri.tile_row_start, ri.tile_col_end, ri.tile_row_end = (int(v) for v in s["tile_rect"])
        b = B.BlockInter()
        b.bx4, b.by4, b.bw4, b.bh4 = int(s["bx4"]), int(s["by4"]), int(s["w4"]), int(s["h4"])
        b.comp_type, b.motion_mode = COMP_TYPE[int(s["comp_kind"])], int(s["pad"][0])
        if s["comp_kind"] == 255:                                      # MM_WARP with a valid model in t->warpmv
            b.motion_mode, b.warp = 2, 1
            for k in range(6):
                b.warp_matrix[k] = int(s["warp_matrix"][k])
            for k in range(4):
                b.warp_abcd[k] = int(s["warp_abcd"][k])
        for k in range(2):
            b.mvx[k], b.mvy[k], b.ref[k] = int(s["mvx"][k]), int(s["mvy"][k]), int(s["ref"][k])
        b.filter2d, b.mask_sign, b.skip = int(s["filter2d"]), int(s["mask_sign"]), int(s["skip"])
        b.max_ytx, b.uvtx = int(s["max_ytx"]), int(s["uvtx"])
        b.tx_split[0] = 1 if s["tx_split"] else 0
        wi, hi, wedge_idx = l2.get(int(s["w4"])), l2.get(int(s["h4"])), int(s["pad"][1])
        if b.comp_type == 4:                                           # the driver's WEDGE_MASK picks (recon_tmpl.c:1861-1866)
            b.wedge_mask[0] = mask_tab.base + mask_tab.wedge[0][wi][hi][0][wedge_idx]
            b.wedge_mask[1] = b.wedge_mask[2] = mask_tab.base + mask_tab.wedge[lay_c][wi][hi][b.mask_sign][wedge_idx]
        if s["pad"][2]:                                                # inter-intra: II_MASK of each plane is in the byte pool
            b.interintra_type, b.interintra_mode = int(s["pad"][2]) & 3, int(s["pad"][2]) >> 2
            for pl in range(1 if hf.no_chroma else 3):                 # where the generator's operations say it put them
                o = ops[ri.n_intra + pl]
                assert o["mode"] == 16 and o["plane"] == pl, o
                b.ii_mask_off[pl] = int(o["coef_off"])
        mine = txr[s["first_tx"]:s["first_tx"] + s["n_tx"]]
        arr = (B.TxCoef * max(len(mine), 1))()
        for k, t in enumerate(mine):
            arr[k].coef_off, arr[k].eob, arr[k].txtp, arr[k].cw4, arr[k].ch4 = (int(t["coef_off"]), int(t["eob"]),
                                                                                 int(t["txtp"]), int(t["cw4"]), int(t["ch4"]))
        n = L.dav1d_cuda_record_b_inter(C.byref(r), C.byref(b), arr, len(mine))
        assert n >= 0, (n, s)
    got = {"put": out["put"][:r.n_put],
           "comp": np.concatenate([out["comp0"][:r.n_comp[0]], out["comp1"][:r.n_comp[1]]]),
           "obmc": np.concatenate([out["obmc0"][:r.n_obmc[0]], out["obmc1"][:r.n_obmc[1]]]),
           "scaled": np.concatenate([sc[k][:r.n_scaled[k]] for k in range(4)]),
           "itx": out["itx"][:r.n_itx], "intra": ops_out[:ri.n_intra], "masks": masks[:r.masks_bytes],
           "warp": warps[:r.n_warp]}
    return got, [r.n_scaled[k] for k in range(4)], r.masks_bytes


def _is_seg_chroma(hf, d):
    """A MASK-kind chroma descriptor that reads the luma-derived segmentation mask (second wave) rather than a
    wedge table: its aux_off is the one of a W_MASK luma descriptor."""
    seg = getattr(hf, "_seg_offs", None)
    if seg is None:
        comp = np.frombuffer(hf.mc_comp.tobytes(), dtype=MC)
        seg = hf._seg_offs = set(int(c["aux_off"]) for c in comp if c["kind"] == 4)
    return int(d["aux_off"]) in seg


def _normalise_edge_bits(a):
    # top-right / bottom-left only reach the predictor through have_top / have_left (see above)
    a = a.copy()
    a["edge_flags"] &= np.where(a["y4"] > a["tile_y4_start"], 0xff, 0xfe).astype(np.uint8)
    a["edge_flags"] &= np.where(a["x4"] > a["tile_x4_start"], 0xff, 0xf7).astype(np.uint8)
    return a


@pytest.mark.parametrize("name", [n for n in R.CASES if n.startswith(("inter_", "obmc_", "scaled_", "wedge_", "ii_", "warp_", "ibc_", "sub8x8_"))])
def test_inter_recorder_emits_the_generators_descriptors(name):
    """dav1d_cuda_record_b_inter over the Av1Block-style records == the descriptor arrays the generator wrote
    for the same blocks (which reproduce dav1d_recon_b_inter's pixels bit for bit, tests/test_reference_driver.py).
    On the wedge_ / ii_ frames that includes the mask pool (the reference's own wedge tables) and the intra-class
    operations of the inter-intra blocks, interleaved in decode order with the intra blocks'."""
    hf, _ = R.make(name)
    mask_tab = None
    if name.startswith(("wedge_", "ii_")):
        import refdsp
        import refframe
        mask_tab = refframe.reference_mask_tab(refdsp.RefDSP())
    got, n_scaled, masks_bytes = record_inter_frame(hf, mask_tab)
    want_ops = _normalise_edge_bits(np.frombuffer(hf.intra.tobytes(), dtype=OP))
    got_ops = _normalise_edge_bits(got["intra"])
    assert len(want_ops) == len(got_ops), (len(want_ops), len(got_ops))
    for f in OP.names:
        if f in ("reserved", "pad"):
            continue
        bad = np.nonzero(want_ops[f] != got_ops[f])[0]
        assert bad.size == 0, f"{name}: field {f} differs at operation {bad[0]}: {want_ops[bad[0]]} vs {got_ops[bad[0]]}"
    if mask_tab is not None:
        # segmentation masks are only allotted (the device writes them); wedge masks are the bytes themselves
        wedge = np.zeros(masks_bytes, dtype=bool)
        for d in np.frombuffer(hf.mc_comp.tobytes(), dtype=MC):
            if d["kind"] == 3 and d["plane"] == 0 or (d["kind"] == 3 and not _is_seg_chroma(hf, d)):
                wedge[d["aux_off"]:d["aux_off"] + int(d["w"]) * int(d["h"])] = True
        assert np.array_equal(got["masks"][wedge], np.asarray(hf.masks)[:masks_bytes][wedge])
    want = {"put": np.frombuffer(hf.mc_put.tobytes(), dtype=MC), "comp": np.frombuffer(hf.mc_comp.tobytes(), dtype=MC),
            "obmc": np.frombuffer(hf.mc_obmc.tobytes(), dtype=MC), "scaled": np.frombuffer(hf.mc_scaled.tobytes(), dtype=MCS),
            "itx": np.frombuffer(hf.itx.tobytes(), dtype=ITX)}
    assert tuple(n_scaled) == tuple(hf.n_mc_scaled)
    assert masks_bytes == hf.masks.nbytes
    assert got["warp"].tobytes() == np.frombuffer(hf.warp.tobytes(), dtype=WARP).tobytes(), name   # decode order
    for key in ("put", "comp", "obmc", "scaled"):
        w, g = want[key].copy(), got[key].copy()
        assert len(w) == len(g), (key, len(w), len(g))
        # the second source of a single-reference prediction is never read
        single = np.isin(w["kind"], (0, 6, 7))
        for a in (w, g):
            a["src"][single, 1] = 0
        assert w.tobytes() == g.tobytes(), (name, key, next(i for i in range(len(w)) if w[i] != g[i]))
    # the generator hands `itx` over sorted by size class; same multiset
    w, g = np.sort(want["itx"], order=list(ITX.names)), np.sort(got["itx"], order=list(ITX.names))
    assert w.tobytes() == g.tobytes(), name


def test_inter_recorder_refuses_what_it_does_not_transcribe():
    L = pkg.lib()
    r = B.InterRecorder()
    r.bw4 = r.bh4 = 16
    r.w = r.h = 64
    r.layout = 1
    nb = np.zeros(17, dtype=np.uint64)
    r.above = r.left = nb.ctypes.data
    put = np.zeros(8, dtype=MC)
    r.put, r.cap_put = put.ctypes.data, 2
    b = B.BlockInter()
    b.bw4 = b.bh4 = 4
    b.skip = 1
    b.motion_mode, b.warp = 2, 1                                       # MM_WARP without a warp array
    assert L.dav1d_cuda_record_b_inter(C.byref(r), C.byref(b), None, 0) == -22
    b.motion_mode, b.warp, b.interintra_type = 0, 0, 1
    assert L.dav1d_cuda_record_b_inter(C.byref(r), C.byref(b), None, 0) == -22    # inter-intra without an intra recorder
    b.interintra_type, b.comp_type = 0, 4                              # COMP_INTER_WEDGE
    assert L.dav1d_cuda_record_b_inter(C.byref(r), C.byref(b), None, 0) == -22    # ... without the block's masks
    b.bw4 = 16
    assert L.dav1d_cuda_record_b_inter(C.byref(r), C.byref(b), None, 0) == -22    # no wedge masks beyond 32x32
    b.bw4 = 4
    b.comp_type = 0
    assert L.dav1d_cuda_record_b_inter(C.byref(r), C.byref(b), None, 0) == -28 and r.n_put == 0   # three planes, room for two
    r.cap_put = 8
    assert L.dav1d_cuda_record_b_inter(C.byref(r), C.byref(b), None, 0) == 3 and r.n_put == 3
    b.skip = 0
    assert L.dav1d_cuda_record_b_inter(C.byref(r), C.byref(b), None, 0) == -22 and r.n_put == 3  # cbi entries missing
