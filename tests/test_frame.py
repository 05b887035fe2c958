"""Frame-level parity: the batched CUDA reconstruction (MC -> inter residual ->
intra executor) against the sequential oracle that replays the same
descriptors in decode order through the reference's C DSP tables."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import _d1pkg

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402
from dav1d_mirror_b200 import binding as B  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "frame_md5.json")

CASES = {
    # name: (w, h, bdmax, seed, kwargs)
    "420_8b_small": (256, 192, 0xff, 1, {}),
    "420_10b_small": (320, 256, 0x3ff, 2, {}),
    "420_12b_small": (192, 256, 0xfff, 3, {}),
    "444_10b_small": (256, 128, 0x3ff, 4, {"ss_hor": 0, "ss_ver": 0}),
    "luma_8b_intra_only": (256, 256, 0xff, 5, {"no_chroma": 1, "p_intra": 1.0}),
    "420_10b_intra_heavy": (384, 256, 0x3ff, 6, {"p_intra": 0.8, "p_palette": 0.1, "p_cfl": 0.6,
                                                 "p_filter_intra": 0.2}),
    "420_8b_inter_only": (384, 320, 0xff, 7, {"p_intra": 0.0, "mv_range": 300}),
    "422_10b_intra_heavy": (256, 192, 0x3ff, 9, {"ss_hor": 1, "ss_ver": 0, "p_intra": 0.8, "p_cfl": 0.6,
                                                 "p_palette": 0.1}),
    "422_8b_small": (192, 128, 0xff, 13, {"ss_hor": 1, "ss_ver": 0}),
    # OBMC: neighbour predictions blended onto single-reference blocks (obmc(), recon_tmpl.c:1071-1131)
    "420_10b_obmc": (320, 256, 0x3ff, 14, {"p_obmc": 0.6, "p_intra": 0.1, "p_avg": 0.05, "p_w_avg": 0.05,
                                           "p_wedge": 0.02, "p_seg": 0.02}),
    "444_8b_obmc": (256, 192, 0xff, 15, {"p_obmc": 0.5, "ss_hor": 0, "ss_ver": 0}),
    # inter-intra: whole-block intra prediction blended onto an inter prediction, residuals in the wavefront
    "420_10b_interintra": (320, 256, 0x3ff, 16, {"p_ii": 0.5, "p_intra": 0.2, "p_obmc": 0.2}),
    "444_8b_interintra": (256, 192, 0xff, 17, {"p_ii": 0.6, "ss_hor": 0, "ss_ver": 0}),
    "422_12b_interintra": (256, 192, 0xfff, 18, {"p_ii": 0.6, "ss_hor": 1, "ss_ver": 0, "p_intra": 0.3}),
    # intrabc: blocks copied (bilinear, half-pel in subsampled chroma) from the decoded part of the current picture
    "420_10b_intrabc": (384, 320, 0x3ff, 19, {"p_ibc": 0.5, "p_intra": 0.7}),
    "444_8b_intrabc": (256, 256, 0xff, 20, {"p_ibc": 0.6, "p_intra": 1.0, "ss_hor": 0, "ss_ver": 0}),
    "422_12b_intrabc": (256, 256, 0xfff, 21, {"p_ibc": 0.5, "p_intra": 0.8, "ss_hor": 1, "ss_ver": 0}),
    # ragged picture sizes (not multiples of the 64x64 superblock / of 8 in chroma)
    "420_10b_ragged": (328, 200, 0x3ff, 8, {}),
    "420_8b_ragged": (200, 120, 0xff, 10, {"p_intra": 0.6}),
    # several tiles: nothing is predicted across a tile edge (have_left / have_top / the w, h arguments of
    # dav1d_prepare_intra_edges come from the tile, recon_tmpl.c:1283-1287; OBMC and intrabc stay inside it)
    "420_10b_tiles_2x2": (384, 256, 0x3ff, 22, {"tile_cols": 2, "tile_rows": 2, "p_intra": 0.6, "p_cfl": 0.5}),
    "420_8b_tiles_3x2_obmc_ibc": (448, 256, 0xff, 23, {"tile_cols": 3, "tile_rows": 2, "p_intra": 0.5, "p_obmc": 0.4,
                                                       "p_ibc": 0.3, "p_ii": 0.2}),
    "444_12b_tiles_2x1": (256, 128, 0xfff, 24, {"tile_cols": 2, "tile_rows": 1, "ss_hor": 0, "ss_ver": 0, "p_intra": 0.7}),
    # dense coefficient blocks (the reference's layout, cw4 = ch4 = 0) through the batched path
    "420_10b_dense_coefs": (256, 192, 0x3ff, 12, {"dense_coefs": 1}),
    # references of another size: the scaled branch of mc() (recon_tmpl.c:1010-1065) as Dav1dCudaMcScaledDesc -
    # reference 0 half the frame's size and reference 1 the frame's own (compounds mix both kinds), both
    # references larger (2x / 1.25x1.5), odd ratios with long vectors that leave the reference
    "420_10b_scaled_half_and_same": (320, 256, 0x3ff, 25, {"ref_w": [160, 0], "ref_h": [128, 0], "p_obmc": 0.3,
                                                           "p_intra": 0.15}),
    "444_8b_scaled_up": (256, 192, 0xff, 26, {"ref_w": [512, 320], "ref_h": [384, 288], "ss_hor": 0, "ss_ver": 0,
                                              "p_intra": 0.1, "p_obmc": 0.3, "p_avg": 0.2, "p_seg": 0.15}),
    "422_12b_scaled_odd_long_vectors": (264, 200, 0xfff, 27, {"ref_w": [200, 376], "ref_h": [120, 312], "ss_hor": 1,
                                                              "ss_ver": 0, "mv_range": 400, "p_intra": 0.1,
                                                              "p_wedge": 0.2, "p_seg": 0.2}),
}


def oracle_planes(ref, hf, seed):
    import refframe
    refs = [F.random_planes(hf, seed * 100 + r, ref=r) for r in range(2)]
    init = F.random_planes(hf, seed * 100 + 50)
    out = refframe.run_oracle(ref, hf, [p.copy() for p in init], refs)
    return refs, init, out


def md5_planes(planes):
    m = hashlib.md5()
    for p in planes:
        m.update(np.ascontiguousarray(p).tobytes())
    return m.hexdigest()


def test_generator_and_oracle_match_golden(ref):
    """CPU: generator + reference-driven oracle reproduce the committed checksums
    (made by tools/make_golden.py with the reference compiled in this container)."""
    with open(GOLDEN) as f:
        gold = json.load(f)
    for name, (w, h, bd, seed, kw) in CASES.items():
        hf = F.HostFrame(w, h, bd, seed, **kw)
        _, _, out = oracle_planes(ref, hf, seed)
        assert md5_planes(out) == gold[name], name


def run_gpu(hf, refs, init, use_graph=False, tasks=True):
    ctx = F.open_context(0)
    df = F.DeviceFrame(ctx, hf, n_refs=len(refs), tasks=tasks)
    try:
        df.upload_descriptors()
        for r, planes in enumerate(refs):
            df.upload_picture(df.refs[r], planes)
        df.upload_picture(df.dst, init)
        if use_graph:
            df.build_graph()
            df.launch_graph()
        else:
            df.submit()
        out = df.download_picture()
        pkg.check_error()
        assert df.cellmap_is_clear(), "the cell map must be back at zero after a frame"
    finally:
        df.close()
        pkg.lib().dav1d_cuda_close(ctx)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_frame_parity_small(ref, name):
    w, h, bd, seed, kw = CASES[name]
    hf = F.HostFrame(w, h, bd, seed, **kw)
    refs, init, want = oracle_planes(ref, hf, seed)
    # explicit transform tasks (the benchmark's configuration); implicit per-size transform runs.
    # First with the dependency levels worked out on the device, then with the recorder's pass
    # (dav1d_cuda_intra_levels) over the same descriptors.
    for tasks in (True, False):
        if not tasks:
            hf.record_levels()
        got = run_gpu(hf, refs, init, use_graph=(seed % 2 == 0), tasks=tasks)
        for pl, (a, b) in enumerate(zip(want, got)):
            bad = np.argwhere(a != b)
            assert bad.size == 0, (f"{name} tasks={tasks}: plane {pl} first "
                                   f"mismatch at (y,x)={bad[0]} ref={a[tuple(bad[0])]} got={b[tuple(bad[0])]} "
                                   f"n={len(bad)}")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bd,seed,kw", [
    (704, 480, 0x3ff, 901, {"p_intra": 0.1, "mv_range": 48}),
    (704, 480, 0xff, 902, {"p_intra": 0.1, "mv_range": 48}),
    (640, 360, 0xfff, 903, {"p_intra": 0.0, "ss_hor": 0, "ss_ver": 0, "p_avg": 0.3, "p_w_avg": 0.2, "p_seg": 0.2}),
    (328, 200, 0x3ff, 904, {"p_intra": 0.0, "ss_hor": 1, "ss_ver": 0, "mv_range": 400}),
    (256, 192, 0xff, 905, {"p_intra": 0.0, "no_chroma": 1, "p_obmc": 0.3}),
])
def test_mc_window_staging_tma_and_cp_async(ref, w, h, bd, seed, kw):
    """The 32x32-tile MC kernels stage the reference windows with cp.async.bulk.tensor (pictures from
    dav1d_cuda_picture_alloc carry tensor maps; windows at the picture border keep the clamped path) - mode 2:
    all of them, 1 (default): the single-reference ones - or, switched off, with per-lane cp.async: the
    reference's pixels in every mode, on frames whose windows are mostly interior
    (short vectors), mostly at the border (long vectors on a small frame) and on odd plane sizes."""
    L = pkg.lib()
    hf = F.HostFrame(w, h, bd, seed, **kw)
    refs, init, want = oracle_planes(ref, hf, seed)
    assert L.dav1d_cuda_get_mc_tma() == 1
    try:
        for on in (2, 1, 0):
            L.dav1d_cuda_set_mc_tma(on)
            got = run_gpu(hf, refs, init)
            for pl, (a, b) in enumerate(zip(want, got)):
                bad = np.argwhere(a != b)
                assert bad.size == 0, (f"tma={on} plane {pl}: first mismatch at (y,x)={bad[0]} n={len(bad)}")
    finally:
        L.dav1d_cuda_set_mc_tma(1)


@pytest.mark.gpu
def test_allocated_pictures_carry_tensor_maps():
    L = pkg.lib()
    ctx = F.open_context(0)
    pic = B.Picture()
    assert L.dav1d_cuda_picture_alloc(ctx, C.byref(pic), 320, 200, 1, 1, 0x3ff) == 0
    assert pic.tma and pic.tma % 128 == 0
    L.dav1d_cuda_picture_free(ctx, C.byref(pic))
    assert not pic.tma
    L.dav1d_cuda_close(ctx)


def big_coefs(hf, seed):
    """Some coefficients beyond int16 (legal at 10 / 12 bit: |c| <= cf_max = 2^(bitdepth+7) - 1, recon_tmpl.c:594),
    the sentinel value itself and the int16 limits."""
    rng = np.random.default_rng(seed)
    cf = hf.cf.view(np.int32)
    nz = np.flatnonzero(cf)
    pick = rng.choice(nz, size=min(200, len(nz)), replace=False)
    cf_max = (128 << (12 if hf.bdmax > 0x3ff else 10)) - 1
    vals = rng.integers(32768, cf_max + 1, size=len(pick)) * rng.choice([-1, 1], size=len(pick))
    vals[:6] = [-32768, 32767, -32767, 32768, -32769, cf_max]
    cf[pick] = vals
    return len(pick)


def test_pack_coefs_round_trip():
    hf = F.HostFrame(256, 192, 0x3ff, 77)
    n_big = big_coefs(hf, 1)
    k = hf.pack_coefs()
    cf = hf.cf.view(np.int32)
    c16 = hf.cf16.view(np.int16)[:len(cf)]
    esc = hf.cf_esc.view(np.int32).reshape(-1, 2)[:k]
    assert 0 < k <= n_big and np.all(np.diff(esc[:, 0]) > 0)
    back = c16.astype(np.int32)
    assert np.all(back[esc[:, 0]] == -32768)
    back[esc[:, 0]] = esc[:, 1]
    assert np.array_equal(back, cf)
    small = np.zeros(4, dtype=np.int32)
    small[:] = [1, 70000, -70000, 3]
    out = np.zeros(4, dtype=np.int16)
    e = np.zeros((1, 2), dtype=np.int32)
    assert pkg.lib().dav1d_cuda_pack_coefs(small.ctypes.data, 4, out.ctypes.data, e.ctypes.data, 1) == -28   # -ENOSPC


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["420_10b_small", "420_12b_small", "422_10b_intra_heavy", "420_10b_dense_coefs"])
def test_compact_coefficient_stream(ref, name):
    """High bit depth with the int16 coefficient stream + escape list (Dav1dCudaReconBatch.cf_int16): same
    pixels as the reference computes from the int32 stream, including coefficients that do not fit int16."""
    w, h, bd, seed, kw = CASES[name]
    hf = F.HostFrame(w, h, bd, seed, **kw)
    big_coefs(hf, seed)
    refs, init, want = oracle_planes(ref, hf, seed)
    assert hf.pack_coefs() > 0
    for tasks in (True, False):
        got = run_gpu(hf, refs, init, tasks=tasks)
        for pl, (a, b) in enumerate(zip(want, got)):
            bad = np.argwhere(a != b)
            assert bad.size == 0, f"{name} tasks={tasks}: plane {pl}: {len(bad)} pixels differ, first {bad[0]}"


@pytest.mark.gpu
def test_config2_itx_1080p_8bit(ref):
    """BASELINE config 2: batched itxfm_add over all 19 sizes x valid types, 8-bit, 1080p-worth."""
    for tx in range(19):
        hf = F.HostFrame(1920, 1080 - 1080 % 8, 0xff, 100 + tx, no_chroma=1, only_tx=tx)
        refs, init, want = oracle_planes(ref, hf, 100 + tx)
        got = run_gpu(hf, refs, init)
        assert np.array_equal(want[0], got[0]), f"tx={tx}"


@pytest.mark.gpu
def test_config3_mc_1080p_8bit(ref):
    """BASELINE config 3: MC put/prep 8-tap + avg/w_avg/mask + warp over a 1080p frame, random MVs, 8-bit."""
    hf = F.HostFrame(1920, 1080, 0xff, 300, p_intra=0.0, p_residual=0.0)
    refs, init, want = oracle_planes(ref, hf, 300)
    got = run_gpu(hf, refs, init)
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("bd", [0x3ff, 0xfff])
def test_config4_full_4k(ref, bd):
    """BASELINE config 4 (+ the 12-bit spot check of config 5): full synthetic 4K reconstruction."""
    hf = F.HostFrame(3840, 2160, bd, 400 + bd)
    refs, init, want = oracle_planes(ref, hf, 400)
    for rep in range(3):      # the executor's schedule is timing dependent: repeat
        got = run_gpu(hf, refs, init, use_graph=(rep == 1))
        for pl, (a, b) in enumerate(zip(want, got)):
            bad = np.argwhere(a != b)
            assert bad.size == 0, f"rep {rep} plane {pl}: {len(bad)} mismatches, first at {bad[0]}"


MULTI_SPECS = {
    "small": [(256, 192, 0x3ff, 21, {}), (256, 192, 0x3ff, 22, {"p_intra": 0.7}), (256, 192, 0xfff, 23, {}),
              (256, 192, 0x3ff, 24, {"p_intra": 0.0})],
    "1080p": [(1920, 1080, 0x3ff, 31, {"p_intra": 0.6}), (1920, 1080, 0x3ff, 32, {}),
              (1280, 720, 0x3ff, 33, {"p_intra": 1.0})],
    "8bit": [(640, 368, 0xff, 41, {"p_intra": 0.8, "p_ibc": 0.3}), (640, 368, 0xff, 42, {"p_obmc": 0.5, "p_ii": 0.3})],
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(MULTI_SPECS))
def test_multi_frame_group(ref, case):
    """Several independent streams in one submission (dav1d_cuda_recon_group_submit, then the same
    launches replayed from a captured graph): every frame must match its own sequential oracle."""
    specs = MULTI_SPECS[case]
    hfs = [F.HostFrame(w, h, bd, seed, **kw) for (w, h, bd, seed, kw) in specs]
    want = [oracle_planes(ref, hf, sp[3]) for hf, sp in zip(hfs, specs)]
    ctx = F.open_context(0)
    dfs = []
    for hf, (refs, init, _) in zip(hfs, want):
        df = F.DeviceFrame(ctx, hf)
        df.upload_descriptors()
        for r, planes in enumerate(refs):
            df.upload_picture(df.refs[r], planes)
        df.upload_picture(df.dst, init)
        dfs.append(df)
    for graph in (False, True):
        for df, (_, init, _) in zip(dfs, want):
            df.upload_picture(df.dst, init)
        mf = F.MultiFrame(ctx, dfs, graph=graph)
        mf.launch()
        for i, (df, (_, _, exp)) in enumerate(zip(dfs, want)):
            got = df.download_picture()
            for pl, (a, b) in enumerate(zip(exp, got)):
                assert np.array_equal(a, b), f"graph={graph} frame {i} plane {pl}"
        pkg.check_error()
        mf.close()
    for df in dfs:
        df.close()
    pkg.lib().dav1d_cuda_close(ctx)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["420_10b_small", "420_8b_ragged"])
def test_end_to_end_pinned_path(ref, name):
    """The end-to-end path of bench.py: descriptors + coefficients from ONE pinned arena (one H2D
    copy), group graph, the reconstructed picture back with ONE D2H copy - bit-exact vs the oracle,
    twice in a row (the arena is re-shipped every step)."""
    w, h, bd, seed, kw = CASES[name]
    hf = F.HostFrame(w, h, bd, seed, **kw)
    refs, init, want = oracle_planes(ref, hf, seed)
    ctx = F.open_context(0)
    df = F.DeviceFrame(ctx, hf, n_refs=len(refs))
    try:
        for r, planes in enumerate(refs):
            df.upload_picture(df.refs[r], planes)
        df.alloc_pinned()
        mf = F.MultiFrame(ctx, [df])
        for _ in range(2):
            df.upload_picture(df.dst, init)
            df.upload_descriptors_pinned()
            mf.launch()
            df.download_pinned()
            pkg.lib().dav1d_cuda_synchronize(ctx)
            for pl, (a, b) in enumerate(zip(want, df.pinned_planes())):
                assert np.array_equal(a, b), f"plane {pl}"
        pkg.check_error()
        mf.close()
    finally:
        df.close()
        pkg.lib().dav1d_cuda_close(ctx)


@pytest.mark.gpu
def test_empty_batch_is_a_no_op(ref):
    """A frame without any descriptor (and a group that contains one): nothing is launched, the
    destination picture is untouched, no error is raised."""
    hf = F.HostFrame(128, 128, 0x3ff, 77, p_intra=0.0)
    init = F.random_planes(hf, 5)
    ctx = F.open_context(0)
    df = F.DeviceFrame(ctx, hf)
    try:
        df.upload_picture(df.dst, init)
        b = df.batch
        b.n_mc_put_tiles = 0
        b.n_mc_put_small = 0
        b.n_mc_comp_tiles[0] = b.n_mc_comp_tiles[1] = 0
        b.n_mc_comp_small[0] = b.n_mc_comp_small[1] = 0
        b.n_warp = 0
        b.n_itx_tasks[0] = b.n_itx_tasks[1] = 0
        for i in range(19):
            b.itx_class_count[i] = 0
        b.n_intra = 0
        l0 = pkg.lib().dav1d_cuda_launch_count()
        df.submit()
        got = df.download_picture()
        assert pkg.lib().dav1d_cuda_launch_count() == l0
        for a, g in zip(init, got):
            assert np.array_equal(a, g)
        mf = F.MultiFrame(ctx, [df])
        mf.launch()
        got = df.download_picture()
        for a, g in zip(init, got):
            assert np.array_equal(a, g)
        pkg.check_error()
        mf.close()
    finally:
        df.close()
        pkg.lib().dav1d_cuda_close(ctx)


@pytest.mark.gpu
def test_inconsistent_descriptors_are_reported_not_hidden():
    """Two intra-class operations that both claim to be the (first) writer of the same cells: whatever
    reads those cells can never learn their level.  The launch must terminate, dav1d_cuda_synchronize()
    must return an error (-EIO) and the sticky error must be set - never a silently wrong frame."""
    import ctypes as C
    L = pkg.lib()
    hf = F.HostFrame(256, 192, 0x3ff, 31, p_intra=1.0)
    d = hf.intra.reshape(-1, 40)
    # duplicate a regular prediction in the middle of the frame (decode order kept for everything else)
    k = next(i for i in range(len(d) // 3, len(d)) if d[i, 15] <= 12 and d[i, 0] > 8 and d[i, 2] > 8)
    hf.intra = np.concatenate([d[:k + 1], d[k:k + 1], d[k + 1:]]).reshape(-1).copy()
    hf.n_intra += 1
    hf.intra_itx = hf.intra_itx[:0]            # keep it simple: no residual pre-pass lists
    hf.intra_itx_class_count = [0] * 19
    hf.intra_itx_tasks = hf.intra_itx_tasks[:4] * 0
    hf.n_intra_itx_tasks = (0, 0)
    ctx = F.open_context(0)
    df = F.DeviceFrame(ctx, hf, n_refs=0)
    try:
        df.upload_descriptors()
        df.upload_picture(df.dst, F.random_planes(hf, 1))
        df.submit()
        r = L.dav1d_cuda_synchronize(ctx)
        assert r != 0 and L.dav1d_cuda_last_error() != 0
        L.dav1d_cuda_clear_error()
        assert L.dav1d_cuda_synchronize(ctx) == 0      # the status word was consumed; the context is usable again
    finally:
        df.close()
        L.dav1d_cuda_close(ctx)
        L.dav1d_cuda_clear_error()
