"""Intra-prediction parity per DSP function (sweeps of tests/checkasm/ipred.c)
plus the on-device dav1d_prepare_intra_edges against the reference's."""
import ctypes as C

import numpy as np
import pytest

import _d1pkg
from test_mc import padded, dptr, pdt, bd_list, call

Z_ANGLES = [3, 6, 9, 14, 17, 20, 23, 26, 29, 32, 36, 39, 42, 45, 48, 51, 54, 58, 61, 64, 67, 70, 73, 76, 81, 84, 87]
Z1, Z2, Z3, FILTER = 6, 7, 8, 13


def gen_z2_max_wh(rng, sz):
    n = int(rng.integers(0, 1 << 31))
    if n & (1 << 17):
        return (n & (sz - 1)) + 1
    if n & (1 << 16):
        return 65536
    return (n & 65535) + 1


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_intra_pred_all_modes(ref, cuda, hbd):
    rng = np.random.default_rng(100 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    n = 0
    for mode in range(14):
        wmax = 32 if mode == FILTER else 64
        w = 4
        while w <= wmax:
            h = max(w // 4, 4)
            while h <= min(w * 4, wmax):
                iters = 8 if Z1 <= mode <= Z3 else 2
                for it in range(iters):
                    a = maxw = maxh = 0
                    if Z1 <= mode <= Z3:
                        a = (90 * (mode - Z1) + Z_ANGLES[int(rng.integers(0, 27))]) | (int(rng.integers(0, 4)) << 9)
                        if mode == Z2:
                            maxw, maxh = gen_z2_max_wh(rng, w), gen_z2_max_wh(rng, h)
                    elif mode == FILTER:
                        a = int(rng.integers(0, 5)) | (int(rng.integers(0, 8)) << 9)
                    bdmax = int(rng.choice(bd_list(hbd)))
                    tl = np.zeros(257, dtype=pdt(hbd))
                    tl[128 - 2 * h:128 + 2 * w + 1] = rng.integers(0, bdmax + 1, size=2 * h + 2 * w + 1)
                    if it == 1:   # flat / extreme edges
                        tl[:] = bdmax if rng.integers(0, 2) else 0
                    tp = tl.ctypes.data + 128 * tl.itemsize
                    outs = []
                    for T in (R, G):
                        d = padded(rng, w, h, hbd, bdmax)
                        p, s = dptr(d)
                        call(T.intra_pred[mode], [p, s, tp, w, h, a, maxw, maxh], hbd, bdmax)
                        outs.append(d)
                    assert np.array_equal(outs[0], outs[1]), \
                        f"intra_pred mode={mode} w={w} h={h} a={a & 511} flags={a >> 9} maxw={maxw} maxh={maxh} bd={bdmax}"
                    n += 1
                h <<= 1
            w <<= 1
    assert n > 300
    _d1pkg.load_pkg().check_error()


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_cfl_ac(ref, cuda, hbd):
    rng = np.random.default_rng(110 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    for li, (ss_hor, ss_ver) in enumerate(((1, 1), (1, 0), (0, 0))):
        h_step, v_step = 2 >> ss_hor, 2 >> ss_ver
        w = 4
        while w <= (32 >> ss_hor):
            h = max(w // 4, 4)
            while h <= min(w * 4, 32 >> ss_ver):
                w_pad = max((w >> 2) - h_step, 0)
                while w_pad >= 0:
                    h_pad = max((h >> 2) - v_step, 0)
                    while h_pad >= 0:
                        bdmax = int(rng.choice(bd_list(hbd)))
                        luma = rng.integers(0, bdmax + 1, size=(32, 32)).astype(pdt(hbd))
                        outs = []
                        for T in (R, G):
                            ac = np.full(32 * 32, -999, dtype=np.int16)
                            T.cfl_ac[li](ac.ctypes.data, luma.ctypes.data, 32 * luma.itemsize, w_pad, h_pad, w, h)
                            outs.append(ac)
                        assert np.array_equal(outs[0], outs[1]), f"cfl_ac layout={li} w={w} h={h} pad={w_pad},{h_pad}"
                        h_pad -= v_step
                    w_pad -= h_step
                h <<= 1
            w <<= 1
    _d1pkg.load_pkg().check_error()


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_cfl_pred_and_pal_pred(ref, cuda, hbd):
    rng = np.random.default_rng(120 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    for mode in (0, 3, 4, 5):
        w = 4
        while w <= 32:
            h = max(w // 4, 4)
            while h <= min(w * 4, 32):
                bdmax = int(rng.choice(bd_list(hbd)))
                alpha = (int(rng.integers(0, 16)) + 1) * (1 - 2 * int(rng.integers(0, 2)))
                tl = np.zeros(257, dtype=pdt(hbd))
                tl[128 - 2 * h:128 + 2 * w + 1] = rng.integers(0, bdmax + 1, size=2 * h + 2 * w + 1)
                ac = rng.integers(0, (bdmax << 3) + 1, size=w * h).astype(np.int64)
                ac = (ac - (ac.sum() + (w * h >> 1)) // (w * h)).astype(np.int16)
                outs = []
                for T in (R, G):
                    d = padded(rng, w, h, hbd, bdmax)
                    p, s = dptr(d)
                    call(T.cfl_pred[mode], [p, s, tl.ctypes.data + 128 * tl.itemsize, w, h, ac.ctypes.data, alpha],
                         hbd, bdmax)
                    outs.append(d)
                assert np.array_equal(outs[0], outs[1]), f"cfl_pred mode={mode} w={w} h={h} alpha={alpha}"
                h <<= 1
            w <<= 1
    w = 4
    while w <= 64:
        h = max(w // 4, 4)
        while h <= min(w * 4, 64):
            bdmax = int(rng.choice(bd_list(hbd)))
            pal = rng.integers(0, bdmax + 1, size=8).astype(pdt(hbd))
            idx = (rng.integers(0, 256, size=w * h // 2) & 0x77).astype(np.uint8)
            outs = []
            for T in (R, G):
                d = padded(rng, w, h, hbd, bdmax)
                p, s = dptr(d)
                T.pal_pred(p, s, pal.ctypes.data, idx.ctypes.data, w, h)
                outs.append(d)
            assert np.array_equal(outs[0], outs[1]), f"pal_pred w={w} h={h}"
            h <<= 1
        w <<= 1
    _d1pkg.load_pkg().check_error()


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_prepare_intra_edges(ref, hbd):
    """dav1d_cuda_prepare_intra_edges vs dav1d_prepare_intra_edges on a random frame:
    all availability combinations, tile/frame-edge clipping, sb-row top backup."""
    rng = np.random.default_rng(130 + hbd)
    L = _d1pkg.load_pkg().lib()
    P, SZ, I = C.c_void_p, C.c_ssize_t, C.c_int
    fn = L.dav1d_cuda_prepare_intra_edges_16bpc if hbd else L.dav1d_cuda_prepare_intra_edges_8bpc
    fn.argtypes = [I, I, I, I, I, I, I, P, SZ, P, I, C.POINTER(I), I, I, I, P] + ([I] if hbd else [])
    fn.restype = I
    rfn = ref.prepare_intra_edges[hbd]
    FW4, FH4 = 40, 36          # frame in 4-px units
    stride_px = FW4 * 4 + 32
    for it in range(1500):
        bdmax = int(rng.choice(bd_list(hbd)))
        frame = rng.integers(0, bdmax + 1, size=(FH4 * 4 + 8, stride_px)).astype(pdt(hbd))
        sbrow = rng.integers(0, bdmax + 1, size=stride_px).astype(pdt(hbd))
        tw = int(rng.choice([1, 2, 4, 8, 16]))
        th = int(rng.choice([1, 2, 4, 8, 16]))
        if max(tw, th) > 4 * min(tw, th):
            continue
        x = int(rng.integers(0, FW4 - 1))
        y = int(rng.integers(0, FH4 - 1))
        w = int(rng.integers(x + 1, FW4 + 1))
        h = int(rng.integers(y + 1, FH4 + 1))
        have_left = int(x > 0 and rng.integers(0, 4) > 0)
        have_top = int(y > 0 and rng.integers(0, 4) > 0)
        edge_flags = int(rng.choice([0, 1, 8, 9]))
        mode = int(rng.integers(0, 14))
        angle_in = int(rng.integers(-3, 4)) if 1 <= mode <= 8 else 0
        filt = int(rng.integers(0, 2))
        use_sb = bool(have_top and rng.integers(0, 3) == 0)
        dstp = frame.ctypes.data + ((y * 4 + 4) * stride_px + x * 4 + 4) * frame.itemsize
        sbp = (sbrow.ctypes.data + 4 * sbrow.itemsize) if use_sb else None
        res = []
        for f in (rfn, fn):
            edge = np.full(257 + 64, 0x5a, dtype=pdt(hbd))
            ang = C.c_int(angle_in)
            args = [x, have_left, y, have_top, w, h, edge_flags, dstp, stride_px * frame.itemsize, sbp,
                    mode, C.byref(ang), tw, th, filt, edge.ctypes.data + 160 * edge.itemsize]
            if hbd:
                args.append(bdmax)
            m = f(*args)
            res.append((m, ang.value, edge))
        assert res[0][0] == res[1][0] and res[0][1] == res[1][1], (it, res[0][:2], res[1][:2])
        assert np.array_equal(res[0][2], res[1][2]), \
            f"edge mismatch it={it} mode={mode}->{res[0][0]} x={x} y={y} w={w} h={h} tw={tw} th={th} hl={have_left} ht={have_top} ef={edge_flags} sb={use_sb}"
    _d1pkg.load_pkg().check_error()
