"""Motion-compensation parity, per DSP function, CUDA table vs the reference's C
templates - the sweeps of tests/checkasm/mc.c (sizes incl. non-pow2 heights,
all 10 filters, every sub-pel path, 8/10/12 bit, guard band around dst)."""
import numpy as np
import pytest

import _d1pkg

PAD = 8


def mc_h_next(h):
    if h in (4, 8, 16):
        return (h * 3) >> 1
    if h in (6, 12, 24):
        return (h & (h - 1)) * 2
    return h * 2


def bd_list(hbd):
    return (0x3ff, 0xfff) if hbd else (0xff,)


def pdt(hbd):
    return np.uint16 if hbd else np.uint8


def padded(rng, w, h, hbd, bdmax, fill=None):
    """dst rect with a guard band (checkasm PIXEL_RECT, checkasm.h:398-405)."""
    a = np.full((h + 2 * PAD, w + 2 * PAD + 32), 0x99 if not hbd else 0x9999, dtype=pdt(hbd))
    if fill is not None:
        a[PAD:PAD + h, PAD:PAD + w] = fill
    return a


def dptr(a):
    return a.ctypes.data + (PAD * a.shape[1] + PAD) * a.itemsize, a.shape[1] * a.itemsize


def mct_input(rng, n, hbd, bdmax):
    """generate_mct_input (mc.c:114-122): worst-case +/- pattern in the top-left corner."""
    pat = np.array([-1, 0, -1, 0, 0, -1, 0, -1])
    sign = -int(rng.integers(0, 2))
    buf = rng.integers(0, bdmax + 1, size=(n, n)).astype(np.int64)
    buf[:8, :8] = (pat[None, :] ^ pat[:, None] ^ sign) & bdmax
    return buf.astype(pdt(hbd))


def call(fn, args, hbd, bdmax):
    if hbd:
        args = list(args) + [bdmax]
    fn(*args)


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_mc_put(ref, cuda, hbd):
    rng = np.random.default_rng(10 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    n = 0
    for filt in range(10):
        w = 2
        while w <= 128:
            for mxy in range(4):
                h = 2 if w <= 32 else w // 4
                h_max = max(min(w * 4, 128), 32)
                while h <= h_max:
                    mx = int(rng.integers(1, 16)) if mxy & 1 else 0
                    my = int(rng.integers(1, 16)) if mxy & 2 else 0
                    bdmax = int(rng.choice(bd_list(hbd)))
                    src = rng.integers(0, bdmax + 1, size=(135, 135)).astype(pdt(hbd))
                    sp = src.ctypes.data + (135 * 3 + 3) * src.itemsize
                    outs = []
                    for T in (R, G):
                        d = padded(rng, w, h, hbd, bdmax)
                        p, s = dptr(d)
                        call(T.mc[filt], [p, s, sp, 135 * src.itemsize, w, h, mx, my], hbd, bdmax)
                        outs.append(d)
                    assert np.array_equal(outs[0], outs[1]), f"mc filt={filt} w={w} h={h} mx={mx} my={my} bd={bdmax}"
                    n += 1
                    h = mc_h_next(h)
            w <<= 1
    assert n > 1000
    _d1pkg.load_pkg().check_error()


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_mc_prep(ref, cuda, hbd):
    rng = np.random.default_rng(20 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    for filt in range(10):
        w = 4
        while w <= 128:
            for mxy in range(4):
                h = max(w // 4, 4)
                while h <= min(w * 4, 128):
                    mx = int(rng.integers(1, 16)) if mxy & 1 else 0
                    my = int(rng.integers(1, 16)) if mxy & 2 else 0
                    bdmax = int(rng.choice(bd_list(hbd)))
                    src = mct_input(rng, 135, hbd, bdmax)
                    sp = src.ctypes.data + (135 * 3 + 3) * src.itemsize
                    outs = []
                    for T in (R, G):
                        t = np.full(128 * 128 + 64, -12345, dtype=np.int16)
                        call(T.mct[filt], [t.ctypes.data, sp, 135 * src.itemsize, w, h, mx, my], hbd, bdmax)
                        outs.append(t)
                    assert np.array_equal(outs[0], outs[1]), f"mct filt={filt} w={w} h={h} mx={mx} my={my} bd={bdmax}"
                    h <<= 1
            w <<= 1
    _d1pkg.load_pkg().check_error()


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
@pytest.mark.parametrize("prep", [False, True])
def test_mc_scaled(ref, cuda, hbd, prep):
    rng = np.random.default_rng(30 + hbd + 2 * prep)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    for filt in range(10):
        w = 4 if prep else 2
        while w <= 128:
            for p in range(3):
                if prep:
                    h, h_max = max(w // 4, 4), min(w * 4, 128)
                else:
                    h, h_max = (2 if w <= 32 else w // 4), max(min(w * 4, 128), 32)
                while h <= h_max:
                    mx, my = int(rng.integers(0, 1024)), int(rng.integers(0, 1024))
                    dx = int(rng.integers(1, 2049))
                    dy = int(rng.integers(1, 2049)) if p == 0 else p << 10
                    bdmax = int(rng.choice(bd_list(hbd)))
                    src = rng.integers(0, bdmax + 1, size=(263, 263)).astype(pdt(hbd))
                    sp = src.ctypes.data + (263 * 3 + 3) * src.itemsize
                    outs = []
                    for T in (R, G):
                        if prep:
                            t = np.full(128 * 128 + 64, -12345, dtype=np.int16)
                            call(T.mct_scaled[filt], [t.ctypes.data, sp, 263 * src.itemsize, w, h, mx, my, dx, dy],
                                 hbd, bdmax)
                            outs.append(t)
                        else:
                            d = padded(rng, w, h, hbd, bdmax)
                            pp, s = dptr(d)
                            call(T.mc_scaled[filt], [pp, s, sp, 263 * src.itemsize, w, h, mx, my, dx, dy], hbd, bdmax)
                            outs.append(d)
                    assert np.array_equal(outs[0], outs[1]), \
                        f"scaled prep={prep} filt={filt} w={w} h={h} mx={mx} my={my} dx={dx} dy={dy} bd={bdmax}"
                    h = mc_h_next(h)
            w <<= 1
    _d1pkg.load_pkg().check_error()


def init_tmp(R, rng, hbd, bdmax):
    """init_tmp (mc.c:278-287): real prep output of the sharp filter on the worst-case input."""
    tmps = []
    for _ in range(2):
        src = mct_input(rng, 135, hbd, bdmax)
        t = np.zeros(128 * 128, dtype=np.int16)
        call(R.mct[5], [t.ctypes.data, src.ctypes.data + (135 * 3 + 3) * src.itemsize, 135 * src.itemsize,
                        128, 128, 8, 8], hbd, bdmax)
        tmps.append(t)
    return tmps


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_compound_avg_wavg_mask_wmask(ref, cuda, hbd):
    rng = np.random.default_rng(40 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    for bdmax in bd_list(hbd):
        t1, t2 = init_tmp(R, rng, hbd, bdmax)
        w = 4
        while w <= 128:
            h = max(w // 4, 4)
            while h <= min(w * 4, 128):
                weight = int(rng.integers(1, 16))
                mask = rng.integers(0, 65, size=w * h).astype(np.uint8)
                for name, extra in (("avg", []), ("w_avg", [weight]), ("mask", [mask.ctypes.data])):
                    outs = []
                    for T in (R, G):
                        d = padded(rng, w, h, hbd, bdmax)
                        p, s = dptr(d)
                        call(getattr(T, name), [p, s, t1.ctypes.data, t2.ctypes.data, w, h] + extra, hbd, bdmax)
                        outs.append(d)
                    assert np.array_equal(outs[0], outs[1]), f"{name} w={w} h={h} bd={bdmax}"
                if w >= 4 and h >= 4:
                    for ss in range(3):
                        for sign in (0, 1):
                            outs = []
                            for T in (R, G):
                                d = padded(rng, w, h, hbd, bdmax)
                                m = np.full(w * h + 64, 0xAA, dtype=np.uint8)
                                p, s = dptr(d)
                                call(T.w_mask[ss], [p, s, t1.ctypes.data, t2.ctypes.data, w, h, m.ctypes.data, sign],
                                     hbd, bdmax)
                                outs.append((d, m))
                            assert np.array_equal(outs[0][0], outs[1][0]), f"w_mask dst ss={ss} w={w} h={h}"
                            assert np.array_equal(outs[0][1], outs[1][1]), f"w_mask mask ss={ss} w={w} h={h}"
                h <<= 1
            w <<= 1
    _d1pkg.load_pkg().check_error()


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_blend(ref, cuda, hbd):
    rng = np.random.default_rng(50 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]

    def one(name, w, h, with_mask):
        bdmax = int(rng.choice(bd_list(hbd)))
        tmp = rng.integers(0, bdmax + 1, size=128 * 128).astype(pdt(hbd))
        mask = rng.integers(0, 65, size=32 * 32).astype(np.uint8)
        fill = rng.integers(0, bdmax + 1, size=(h, w)).astype(pdt(hbd))
        outs = []
        for T in (R, G):
            d = padded(rng, w, h, hbd, bdmax, fill)
            p, s = dptr(d)
            args = [p, s, tmp.ctypes.data, w, h] + ([mask.ctypes.data] if with_mask else [])
            getattr(T, name)(*args)
            outs.append(d)
        assert np.array_equal(outs[0], outs[1]), f"{name} w={w} h={h}"

    for w in (4, 8, 16, 32):
        h = max(w // 2, 4)
        while h <= min(w * 2, 32):
            one("blend", w, h, True)
            h <<= 1
    for w in (2, 4, 8, 16, 32):
        h = 2
        while h <= (64 if w == 2 else 128):
            one("blend_v", w, h, False)
            h <<= 1
    for w in (2, 4, 8, 16, 32, 64, 128):
        h = 4 if w == 128 else 2
        while h <= 32:
            one("blend_h", w, h, False)
            h <<= 1
    _d1pkg.load_pkg().check_error()


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_warp8x8(ref, cuda, hbd):
    rng = np.random.default_rng(60 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    for it in range(200):
        bdmax = int(rng.choice(bd_list(hbd)))
        mx = int(rng.integers(0, 0x2000)) - 0xa00
        my = int(rng.integers(0, 0x2000)) - 0xa00
        abcd = (rng.integers(0, 0x2000, size=4) - 0xa00).astype(np.int16)
        src = rng.integers(0, bdmax + 1, size=(15, 15)).astype(pdt(hbd))
        sp = src.ctypes.data + (15 * 3 + 3) * src.itemsize
        outs, outt = [], []
        for T in (R, G):
            d = padded(rng, 8, 8, hbd, bdmax)
            p, s = dptr(d)
            call(T.warp8x8, [p, s, sp, 15 * src.itemsize, abcd.ctypes.data, mx, my], hbd, bdmax)
            outs.append(d)
            t = np.full(8 * 24, -777, dtype=np.int16)
            call(T.warp8x8t, [t.ctypes.data, 24 if it & 1 else 8, sp, 15 * src.itemsize, abcd.ctypes.data, mx, my],
                 hbd, bdmax)
            outt.append(t)
        assert np.array_equal(outs[0], outs[1]), f"warp8x8 it={it}"
        assert np.array_equal(outt[0], outt[1]), f"warp8x8t it={it}"
    _d1pkg.load_pkg().check_error()


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_emu_edge(ref, cuda, hbd):
    rng = np.random.default_rng(70 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    bdmax = 0xfff if hbd else 0xff
    src = rng.integers(0, bdmax + 1, size=(160, 160)).astype(pdt(hbd))

    def off(edge2, b):
        # random_offset_for_edge (mc.c:650-675); edge2 bit0 = have "before" side, bit1 = have "after" side
        i = 160 if edge2 else 1 + int(rng.integers(0, b - 2))
        if edge2 == 3:
            pos = int(rng.integers(0, i - b + 1))
        elif edge2 == 1:      # HAVE_LEFT / HAVE_TOP only
            pos = (i - b) + 1 + int(rng.integers(0, b - 1))
        elif edge2 == 2:      # HAVE_RIGHT / HAVE_BOTTOM only
            pos = -(1 + int(rng.integers(0, b - 1)))
        else:
            pos = -(1 + int(rng.integers(0, b - i - 1)))
        return pos, i

    w = 4
    while w <= 128:
        h = max(w // 4, 4)
        while h <= min(w * 4, 128):
            for edge in range(15):
                bw, bh = w + int(rng.integers(0, 8)), h + int(rng.integers(0, 8))
                x, iw = off(((edge >> 2) & 1) | (((edge >> 3) & 1) << 1), bw)
                y, ih = off((edge & 1) | (((edge >> 1) & 1) << 1), bh)
                outs = []
                for T in (R, G):
                    d = np.full((135, 192), 0x55, dtype=pdt(hbd))
                    T.emu_edge(bw, bh, iw, ih, x, y, d.ctypes.data, 192 * d.itemsize, src.ctypes.data,
                               160 * src.itemsize)
                    outs.append(d)
                assert np.array_equal(outs[0], outs[1]), f"emu_edge w={bw} h={bh} iw={iw} ih={ih} x={x} y={y}"
            h <<= 1
        w <<= 1
    _d1pkg.load_pkg().check_error()


def _c_div(a, b):
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def _resize_params(src_w, dst_w):
    """f->resize_step / f->resize_start as decode.c:3365-3369,3576-3583 compute them."""
    dx = ((src_w << 14) + (dst_w >> 1)) // dst_w
    err = dst_w * dx - (src_w << 14)
    x0 = _c_div(-((dst_w - src_w) << 13) + (dst_w >> 1), dst_w) + 128 - _c_div(err, 2)
    return dx, x0 & 0x3fff


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_resize(ref, cuda, hbd):
    """mc.resize (mc_tmpl.c:877-903), the sweep of checkasm's check_resize (mc.c:723-771): random source widths
    16..512, denominators 9..16 (dst_w = w_den * src_w >> 3), 64 rows; plus downscale-free 1:1 and odd strides."""
    rng = np.random.default_rng(90 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    for it in range(40):
        bdmax = (0x3ff if rng.integers(2) else 0xfff) if hbd else 0xff
        w_den = 9 + int(rng.integers(8))
        src_w = 16 + int(rng.integers(512 - 16 + 1))
        if it == 0:
            w_den, src_w = 8, 64                      # 1:1
        dst_w = w_den * src_w >> 3
        dx, mx0 = _resize_params(src_w, dst_w)
        h = 64 if it % 3 else 1 + int(rng.integers(70))
        sstride = 512 + 8 * int(rng.integers(3))
        src = rng.integers(0, bdmax + 1, size=(h, sstride)).astype(pdt(hbd))
        outs = []
        for T in (R, G):
            d = np.full((h, 1040), 0x55, dtype=pdt(hbd))
            args = [d.ctypes.data, 1040 * d.itemsize, src.ctypes.data, sstride * src.itemsize, dst_w, h, src_w, dx, mx0]
            T.resize(*(args + ([bdmax] if hbd else [])))
            outs.append(d)
        assert np.array_equal(outs[0], outs[1]), f"resize src_w={src_w} dst_w={dst_w} dx={dx} mx0={mx0} h={h}"
    _d1pkg.load_pkg().check_error()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,ss_hor,ss_ver,bd,den", [(320, 200, 1, 1, 0x3ff, 12), (264, 136, 1, 0, 0xfff, 9),
                                                     (200, 96, 0, 0, 0xff, 16), (512, 64, 1, 1, 0xff, 11)])
def test_resize_frame(ref, w, h, ss_hor, ss_ver, bd, den):
    """dav1d_cuda_resize_frame == mc.resize of the reference over every row of every plane with
    f->resize_step[] / f->resize_start[] (what dav1d_filter_sbrow_resize, recon_tmpl.c:2104-2137, does one
    superblock row at a time): a frame of width w upscaled to (w * den + 4) / 8."""
    import ctypes as C
    pkg = _d1pkg.load_pkg()
    from dav1d_mirror_b200 import binding as B
    from dav1d_mirror_b200 import frame as F
    L = pkg.lib()
    hbd = bd > 0xff
    rng = np.random.default_rng(w * 7 + den)
    out_w = (w * den + 4) >> 3
    ctx = F.open_context(0)
    src, dst = B.Picture(), B.Picture()
    assert L.dav1d_cuda_picture_alloc(ctx, C.byref(src), w, h, ss_hor, ss_ver, bd) == 0
    assert L.dav1d_cuda_picture_alloc(ctx, C.byref(dst), out_w, h, ss_hor, ss_ver, bd) == 0
    in_cw, out_cw = (w + ss_hor) >> ss_hor, (out_w + ss_hor) >> ss_hor
    step, start = zip(_resize_params(w, out_w), _resize_params(in_cw, out_cw))
    R = ref.bpc[hbd]
    want, planes = [], []
    for pl in range(3):
        pw, ph = (w, h) if pl == 0 else (in_cw, (h + ss_ver) >> ss_ver)
        ow = out_w if pl == 0 else out_cw
        p = rng.integers(0, bd + 1, size=(ph, pw)).astype(pdt(hbd))
        planes.append(p)
        assert L.dav1d_cuda_picture_upload(ctx, C.byref(src), pl, p.ctypes.data, p.strides[0]) == 0
        d = np.zeros((ph, ow), dtype=pdt(hbd))
        args = [d.ctypes.data, d.strides[0], p.ctypes.data, p.strides[0], ow, ph, pw, step[pl > 0], start[pl > 0]]
        R.resize(*(args + ([bd] if hbd else [])))
        want.append(d)
    st, sa = (C.c_int32 * 2)(*step), (C.c_int32 * 2)(*start)
    assert L.dav1d_cuda_resize_frame(ctx, C.byref(dst), C.byref(src), st, sa) == 0
    for pl in range(3):
        got = np.zeros_like(want[pl])
        assert L.dav1d_cuda_picture_download(ctx, C.byref(dst), pl, got.ctypes.data, got.strides[0]) == 0
        L.dav1d_cuda_synchronize(ctx)
        assert np.array_equal(got, want[pl]), f"plane {pl}"
    assert L.dav1d_cuda_resize_frame(ctx, C.byref(src), C.byref(src), st, sa) == -22      # in place
    L.dav1d_cuda_picture_free(ctx, C.byref(src))
    L.dav1d_cuda_picture_free(ctx, C.byref(dst))
    L.dav1d_cuda_close(ctx)
    pkg.check_error()
