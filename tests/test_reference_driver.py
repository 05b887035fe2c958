"""Frame parity against the reference's OWN reconstruction drivers.

oracle/ref_recon.c calls dav1d_recon_b_intra_{8,16}bpc (src/recon_tmpl.c:1195-1596) and
dav1d_recon_b_inter_{8,16}bpc (:1598-2036), compiled where they lie under /root/reference, block by block
on Av1Block-style records of synthetic frames (generator option real_blocks: block contexts, chroma
ownership of 4-pixel blocks, tile resets as src/decode.c does them).  Everything the drivers decide -
predictor and angle, edge flags per transform block, CfL / palette order, smooth-neighbour flags from the
above / left contexts, tile edges, the superblock-row edge backup; for inter blocks the source position
and sub-pel phase per plane, the emu_edge decision, the compound combine and its mask hand-over to
chroma, the transform tree and the cbi / cf consumption order - is the reference's code; the descriptors
the CUDA path consumes are recorded independently by the generator.  Inter coverage: single-reference
and compound (avg, distance-weighted avg, segmentation mask, wedge) blocks with residuals, OBMC (obmc(),
recon_tmpl.c:1071-1132, over refmvs rows the harness fills the way decode.c does after every block),
inter-intra blocks (smooth and wedge blends; the generator copies the masks out of the reference's own
tables) and references of another size; warped and intrabc blocks stay on the descriptor-replay oracle
(tests/test_frame.py).

 * CPU: the descriptor-driven oracle (oracle/ref_frame.c, the bench's CPU arm) reproduces the
   reference driver bit for bit - i.e. the descriptors mean what recon_tmpl.c means;
 * GPU: libdav1d_cuda.so reconstructs the same frames bit for bit."""
import hashlib
import json
import os

import numpy as np
import pytest

import _d1pkg
import refframe

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_driver_md5.json")

CASES = {
    # name: (w, h, bdmax, seed, kwargs) - all blocks intra, Av1Block semantics
    "luma_8b": (256, 256, 0xff, 5, {"no_chroma": 1}),
    "420_10b_cfl_pal_filter": (384, 256, 0x3ff, 6, {"p_palette": 0.1, "p_cfl": 0.6, "p_filter_intra": 0.2}),
    "444_8b_cfl_pal": (256, 192, 0xff, 7, {"ss_hor": 0, "ss_ver": 0, "p_cfl": 0.4, "p_palette": 0.1}),
    "422_12b_cfl_filter": (256, 192, 0xfff, 8, {"ss_hor": 1, "ss_ver": 0, "p_cfl": 0.5, "p_filter_intra": 0.1}),
    "420_10b_tiles_2x2": (384, 256, 0x3ff, 9, {"tile_cols": 2, "tile_rows": 2, "p_cfl": 0.5}),
    "420_8b_ragged": (328, 200, 0xff, 10, {"p_cfl": 0.5, "p_palette": 0.1}),
    "420_12b_no_edge_filter_split": (320, 192, 0xfff, 11, {"edge_filter": 0, "p_tx_split": 1.0, "p_residual": 1.0}),
    "444_10b_tiles_3x2_no_residual": (448, 256, 0x3ff, 12, {"ss_hor": 0, "ss_ver": 0, "tile_cols": 3, "tile_rows": 2,
                                                            "p_residual": 0.0, "p_palette": 0.15}),
    # frames with inter blocks (dav1d_recon_b_inter): put / avg / w_avg / segmentation-mask compound + residual trees
    "inter_420_8b_put_only": (256, 192, 0xff, 31, {"p_intra": 0.0, "p_avg": 0, "p_w_avg": 0, "p_seg": 0}),
    "inter_420_10b_mixed": (320, 256, 0x3ff, 32, {"p_intra": 0.3, "p_avg": 0.2, "p_w_avg": 0.15, "p_seg": 0.15, "p_cfl": 0.4}),
    "inter_444_12b": (256, 192, 0xfff, 33, {"ss_hor": 0, "ss_ver": 0, "p_intra": 0.2, "p_avg": 0.2, "p_w_avg": 0.2, "p_seg": 0.2}),
    "inter_422_10b": (256, 192, 0x3ff, 34, {"ss_hor": 1, "ss_ver": 0, "p_intra": 0.3, "p_avg": 0.2, "p_w_avg": 0.1, "p_seg": 0.2}),
    "inter_420_8b_long_vectors_ragged": (200, 136, 0xff, 35, {"p_intra": 0.1, "mv_range": 300, "p_avg": 0.2, "p_w_avg": 0.1,
                                                            "p_seg": 0.1}),
    "inter_420_10b_tiles_2x2": (384, 256, 0x3ff, 36, {"tile_cols": 2, "tile_rows": 2, "p_intra": 0.4, "p_avg": 0.2, "p_seg": 0.1}),
    "inter_luma_8b": (256, 256, 0xff, 37, {"no_chroma": 1, "p_intra": 0.3, "p_avg": 0.3, "p_seg": 0.2}),
    # overlapped block motion compensation: obmc() reads the ACTUAL above / left neighbours from the refmvs rows
    "obmc_420_8b": (256, 192, 0xff, 41, {"p_intra": 0.2, "p_avg": 0, "p_w_avg": 0, "p_seg": 0, "p_obmc": 0.8}),
    "obmc_420_10b_mixed": (320, 256, 0x3ff, 42, {"p_intra": 0.3, "p_avg": 0.2, "p_w_avg": 0.15, "p_seg": 0.15, "p_cfl": 0.4,
                                                 "p_obmc": 0.5}),
    "obmc_444_12b": (256, 192, 0xfff, 43, {"ss_hor": 0, "ss_ver": 0, "p_intra": 0.2, "p_avg": 0.2, "p_obmc": 0.6}),
    "obmc_422_10b": (256, 192, 0x3ff, 44, {"ss_hor": 1, "ss_ver": 0, "p_intra": 0.3, "p_avg": 0.2, "p_obmc": 0.6}),
    "obmc_420_8b_long_vectors_ragged": (200, 136, 0xff, 45, {"p_intra": 0.1, "mv_range": 300, "p_avg": 0.2, "p_obmc": 0.6}),
    "obmc_420_10b_tiles_2x2": (384, 256, 0x3ff, 46, {"tile_cols": 2, "tile_rows": 2, "p_intra": 0.4, "p_avg": 0.2, "p_obmc": 0.7}),
    # wedge compounds and inter-intra blocks (smooth and wedge blends) with the reference's own mask tables
    # (src/wedge.c through WEDGE_MASK / II_MASK); "masks": the generator is handed those tables
    "wedge_420_8b": (256, 192, 0xff, 101, {"masks": 1, "p_intra": 0.2, "p_avg": 0.1, "p_seg": 0.1, "p_wedge": 0.5}),
    "ii_420_10b": (320, 256, 0x3ff, 102, {"masks": 1, "p_intra": 0.3, "p_avg": 0.1, "p_ii": 0.7, "p_cfl": 0.4}),
    "wedge_ii_444_12b": (256, 192, 0xfff, 103, {"masks": 1, "ss_hor": 0, "ss_ver": 0, "p_intra": 0.2, "p_wedge": 0.3, "p_ii": 0.5}),
    "wedge_ii_422_10b": (256, 192, 0x3ff, 104, {"masks": 1, "ss_hor": 1, "ss_ver": 0, "p_intra": 0.3, "p_wedge": 0.3, "p_ii": 0.5}),
    "wedge_ii_obmc_420_8b_long_vectors_ragged": (200, 136, 0xff, 105, {"masks": 1, "p_intra": 0.1, "mv_range": 300, "p_wedge": 0.3,
                                                                       "p_ii": 0.4, "p_obmc": 0.3}),
    "wedge_ii_luma_8b": (256, 256, 0xff, 107, {"masks": 1, "no_chroma": 1, "p_intra": 0.3, "p_wedge": 0.3, "p_ii": 0.5}),
    # locally warped blocks (MM_WARP): warp_affine() with models whose shear parameters come from the reference's
    # dav1d_get_shear_params; "warps": the generator is handed those models
    "warp_420_10b": (320, 256, 0x3ff, 111, {"warps": 1, "p_intra": 0.2, "p_avg": 0.1, "p_warp": 0.6}),
    "warp_444_8b": (256, 192, 0xff, 112, {"warps": 1, "ss_hor": 0, "ss_ver": 0, "p_intra": 0.1, "p_warp": 0.5, "p_obmc": 0.3}),
    "warp_422_12b_long_vectors_ragged": (264, 200, 0xfff, 113, {"warps": 1, "ss_hor": 1, "ss_ver": 0, "p_intra": 0.1,
                                                               "mv_range": 300, "p_warp": 0.6}),
    "warp_luma_10b_tiles_2x2": (384, 256, 0x3ff, 114, {"warps": 1, "no_chroma": 1, "tile_cols": 2, "tile_rows": 2,
                                                      "p_intra": 0.3, "p_warp": 0.5}),
    # intrabc blocks of key frames: dav1d_recon_b_inter's IS_KEY_OR_INTRA branch, mc() from the picture being decoded
    # (integer and half-pel chroma vectors, sources past the right frame edge, residual trees)
    "ibc_420_10b_cfl": (320, 256, 0x3ff, 121, {"p_ibc": 0.4, "p_cfl": 0.3}),
    "ibc_444_8b": (256, 192, 0xff, 122, {"p_ibc": 0.5, "ss_hor": 0, "ss_ver": 0}),
    "ibc_422_12b_pal_ragged": (264, 200, 0xfff, 123, {"p_ibc": 0.5, "ss_hor": 1, "ss_ver": 0, "p_palette": 0.1}),
    "ibc_420_10b_tiles_2x2": (384, 256, 0x3ff, 124, {"p_ibc": 0.5, "tile_cols": 2, "tile_rows": 2}),
    "ibc_luma_8b": (256, 256, 0xff, 125, {"p_ibc": 0.5, "no_chroma": 1}),
    # 8x4 / 4x8 / 4x4 blocks in 4:2:0: the chroma of the block at the odd position of its 8x8 is predicted part by part
    # with the vectors of its partners (recon_tmpl.c:1685-1751) or, next to an intra partner, with its own
    "sub8x8_420_10b_all_inter": (320, 256, 0x3ff, 131, {"p_sub8x8": 0.6, "p_intra": 0.0}),
    "sub8x8_420_8b_mixed_cfl": (256, 192, 0xff, 132, {"p_sub8x8": 0.5, "p_intra": 0.3, "p_cfl": 0.3}),
    "sub8x8_420_12b_long_vectors_ragged": (264, 200, 0xfff, 133, {"p_sub8x8": 0.7, "p_intra": 0.2, "mv_range": 300}),
    "sub8x8_420_10b_tiles_2x2_obmc": (384, 256, 0x3ff, 134, {"p_sub8x8": 0.5, "p_intra": 0.3, "tile_cols": 2, "tile_rows": 2,
                                                            "p_obmc": 0.4}),
    "sub8x8_420_8b_scaled": (256, 192, 0xff, 135, {"p_sub8x8": 0.6, "p_intra": 0.1, "ref_w": [384, 0], "ref_h": [288, 0]}),
    # 4:2:2: 4x4 blocks (chroma 2x4 parts with the left partner's vector) and 8x4 blocks
    "sub8x8_422_10b_all_inter": (320, 256, 0x3ff, 141, {"p_sub8x8": 0.6, "p_intra": 0.0, "ss_hor": 1, "ss_ver": 0}),
    "sub8x8_422_8b_mixed_obmc_ragged": (264, 200, 0xff, 142, {"p_sub8x8": 0.6, "p_intra": 0.3, "ss_hor": 1, "ss_ver": 0, "p_obmc": 0.3}),
    # 4-px-wide / -high intrabc blocks: the odd block predicts the chroma of its 8x8 with its own vector
    "ibc_sub8x8_420_10b": (320, 256, 0x3ff, 151, {"p_ibc": 0.6, "p_sub8x8": 0.6}),
    "ibc_sub8x8_422_8b": (256, 192, 0xff, 152, {"p_ibc": 0.6, "p_sub8x8": 0.7, "ss_hor": 1, "ss_ver": 0}),
    "ibc_sub8x8_420_12b_tiles_2x2_cfl": (384, 256, 0xfff, 153, {"p_ibc": 0.6, "p_sub8x8": 0.6, "tile_cols": 2, "tile_rows": 2,
                                                               "p_cfl": 0.3}),
    # references of another size: the scaled branch of mc() with f->svc as decode.c:3517-3524 sets it
    "scaled_420_10b_half_and_same": (320, 256, 0x3ff, 51, {"ref_w": [160, 0], "ref_h": [128, 0], "p_intra": 0.2, "p_avg": 0.2,
                                                           "p_w_avg": 0.1, "p_seg": 0.15, "p_obmc": 0.3}),
    "scaled_444_8b_up": (256, 192, 0xff, 52, {"ref_w": [512, 320], "ref_h": [384, 288], "ss_hor": 0, "ss_ver": 0,
                                              "p_intra": 0.1, "p_avg": 0.2, "p_seg": 0.15, "p_obmc": 0.3}),
    "scaled_422_12b_odd_long_vectors": (264, 200, 0xfff, 53, {"ref_w": [200, 376], "ref_h": [120, 312], "ss_hor": 1, "ss_ver": 0,
                                                              "mv_range": 400, "p_intra": 0.1, "p_avg": 0.2, "p_seg": 0.2}),
}


def make(name):
    w, h, bd, seed, kw = CASES[name]
    kw = dict(kw)
    kw.setdefault("p_intra", 1.0)
    kw.setdefault("p_wedge", 0.0)
    kw.setdefault("p_warp", 0.0)
    if kw.pop("masks", 0):
        import refdsp
        kw["mask_tab"] = refframe.reference_mask_tab(refdsp.RefDSP())
    if kw.pop("warps", 0):
        import refdsp
        kw["warp_tab"] = refframe.reference_warp_tab(refdsp.RefDSP())
    hf = F.HostFrame(w, h, bd, seed, real_blocks=1, **kw)
    init = F.random_planes(hf, seed * 10 + 5)
    return hf, init


def refs_of(hf, name):
    seed = CASES[name][3]
    return [F.random_planes(hf, seed * 10 + k, ref=k) for k in range(2)] if hf.params.p_intra < 1.0 else []


def md5_planes(planes):
    m = hashlib.md5()
    for p in planes:
        m.update(np.ascontiguousarray(p).tobytes())
    return m.hexdigest()


@pytest.mark.parametrize("name", list(CASES))
def test_descriptors_mean_what_the_reference_driver_means(ref, name):
    hf, init = make(name)
    refs = refs_of(hf, name)
    want = refframe.run_reference_driver(ref, hf, [p.copy() for p in init], refs)
    got = refframe.run_oracle(ref, hf, [p.copy() for p in init], refs)
    for pl, (a, b) in enumerate(zip(want, got)):
        bad = np.argwhere(a != b)
        assert bad.size == 0, f"{name}: plane {pl}: {len(bad)} pixels differ, first at (y,x)={bad[0]}"
    with open(GOLDEN) as f:
        assert md5_planes(want) == json.load(f)[name], name      # drift of generator / reference build


def random_frames(ref, n, first=0):
    """The random sweep: frame k's parameters are a function of k alone (one generator stream, drawn in order)."""
    rng = np.random.default_rng(20261019)
    for k in range(n):
        lay = [(1, 1), (1, 0), (0, 0)][rng.integers(3)]
        kw = dict(ss_hor=lay[0], ss_ver=lay[1], p_cfl=float(rng.choice([0, 0.5])), p_palette=float(rng.choice([0, 0.15])),
                  p_intra=float(rng.choice([1.0, 0.5, 0.2, 0.0])), p_wedge=0.0, p_warp=0.0,
                  p_avg=float(rng.choice([0, 0.2])), p_w_avg=float(rng.choice([0, 0.2])), p_seg=float(rng.choice([0, 0.2])),
                  p_obmc=float(rng.choice([0, 0.5])), p_ii=float(rng.choice([0, 0.4])), mv_range=int(rng.choice([16, 128, 400])),
                  p_filter_intra=float(rng.choice([0, 0.2])), tile_cols=int(rng.integers(1, 4)),
                  tile_rows=int(rng.integers(1, 3)), p_tx_split=float(rng.choice([0, 0.5, 1.0])),
                  p_residual=float(rng.choice([0.3, 0.6, 1.0])), edge_filter=int(rng.integers(2)))
        w, h = int(rng.integers(8, 60)) * 8, int(rng.integers(8, 40)) * 8
        bd = [0xff, 0x3ff, 0xfff][rng.integers(3)]
        kw["p_wedge"] = float(rng.choice([0, 0.3]))
        kw["p_warp"] = float(rng.choice([0, 0.3]))
        kw["p_sub8x8"] = float(rng.choice([0, 0.5]))          # takes effect in 4:2:0 / 4:2:2
        if k < first:
            continue
        hf = F.HostFrame(w, h, bd, 500 + k, real_blocks=1, mask_tab=refframe.reference_mask_tab(ref),
                         warp_tab=refframe.reference_warp_tab(ref), **kw)
        init = F.random_planes(hf, 9000 + k)
        refs = [F.random_planes(hf, 9100 + 2 * k + j) for j in range(2)]
        yield k, hf, init, refs, (w, h, hex(bd), kw)


def test_random_frames_against_the_reference_drivers(ref):
    for k, hf, init, refs, what in random_frames(ref, 48):
        want = refframe.run_reference_driver(ref, hf, [p.copy() for p in init], refs)
        got = refframe.run_oracle(ref, hf, [p.copy() for p in init], refs)
        assert all(np.array_equal(a, b) for a, b in zip(want, got)), (k, what)


@pytest.mark.gpu
def test_cuda_random_frames_against_the_reference_drivers(ref):
    """The same sweep on the product path: CUDA == dav1d_recon_b_intra / dav1d_recon_b_inter, frame after frame."""
    import test_frame
    for k, hf, init, refs, what in random_frames(ref, 32):
        want = refframe.run_reference_driver(ref, hf, [p.copy() for p in init], refs)
        if k & 1:
            hf.record_levels()
        got = test_frame.run_gpu(hf, refs, init, use_graph=False)
        for pl, (a, b) in enumerate(zip(want, got)):
            bad = np.argwhere(a != b)
            assert bad.size == 0, f"frame {k} {what}: plane {pl}: {len(bad)} pixels differ, first {bad[0]}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_frame_equals_the_reference_driver(ref, name):
    import test_frame
    hf, init = make(name)
    refs = refs_of(hf, name)
    want = refframe.run_reference_driver(ref, hf, [p.copy() for p in init], refs)
    for record in (False, True):
        if record:
            hf.record_levels()
        got = test_frame.run_gpu(hf, refs, init, use_graph=False)
        for pl, (a, b) in enumerate(zip(want, got)):
            bad = np.argwhere(a != b)
            assert bad.size == 0, f"{name} (recorded levels: {record}): plane {pl}: {len(bad)} pixels differ, first {bad[0]}"
