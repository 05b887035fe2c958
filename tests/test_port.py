"""CPU: the numpy restatement oracle/port/recon_np.py against the reference's own C templates
(oracle/_ref) - every function it covers, 8 / 10 / 12 bit, bit-exact."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "port"))
import recon_np as P  # noqa: E402
import refdsp  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    return refdsp.RefDSP()


def pdt(hbd):
    return np.uint16 if hbd else np.uint8


def call(fn, args, hbd, bdmax):
    fn(*(args + ([bdmax] if hbd else [])))


@pytest.mark.parametrize("bdmax", [0xff, 0x3ff, 0xfff])
def test_mc_put_prep(ref, bdmax):
    hbd = bdmax > 0xff
    R = ref.bpc[hbd]
    rng = np.random.default_rng(bdmax)
    n = 0
    for filt in range(10):
        for w in (2, 4, 8, 16, 32):
            for h in (2, 4, 8, 24, 32):
                for mx, my in ((0, 0), (int(rng.integers(1, 16)), 0), (0, int(rng.integers(1, 16))),
                               (int(rng.integers(1, 16)), int(rng.integers(1, 16)))):
                    src = rng.integers(0, bdmax + 1, size=(h + 7, w + 7)).astype(pdt(hbd))
                    sp = src.ctypes.data + (3 * (w + 7) + 3) * src.itemsize
                    d = np.zeros((h, w), pdt(hbd))
                    call(R.mc[filt], [d.ctypes.data, d.strides[0], sp, src.strides[0], w, h, mx, my], hbd, bdmax)
                    assert np.array_equal(d, P.mc_put(src, w, h, mx, my, filt, bdmax)), (filt, w, h, mx, my)
                    if w >= 4 and h >= 4:
                        t = np.zeros((h, w), np.int16)
                        call(R.mct[filt], [t.ctypes.data, sp, src.strides[0], w, h, mx, my], hbd, bdmax)
                        assert np.array_equal(t, P.mc_prep(src, w, h, mx, my, filt, bdmax)), (filt, w, h, mx, my)
                    n += 1
    assert n == 1000


@pytest.mark.parametrize("bdmax", [0xff, 0x3ff, 0xfff])
def test_compound_and_blend(ref, bdmax):
    hbd = bdmax > 0xff
    R = ref.bpc[hbd]
    rng = np.random.default_rng(100 + bdmax)
    ib = P.inter_bits(bdmax)
    for w, h in ((4, 4), (8, 16), (16, 8), (32, 32), (64, 16)):
        lo, hi = -P.prep_bias(bdmax), (bdmax << ib) - P.prep_bias(bdmax)
        t1 = rng.integers(lo, hi + 1, size=(h, w)).astype(np.int16)
        t2 = rng.integers(lo, hi + 1, size=(h, w)).astype(np.int16)
        d = np.zeros((h, w), pdt(hbd))
        call(R.avg, [d.ctypes.data, d.strides[0], t1.ctypes.data, t2.ctypes.data, w, h], hbd, bdmax)
        assert np.array_equal(d, P.avg(t1, t2, bdmax))
        wt = int(rng.integers(1, 16))
        call(R.w_avg, [d.ctypes.data, d.strides[0], t1.ctypes.data, t2.ctypes.data, w, h, wt], hbd, bdmax)
        assert np.array_equal(d, P.w_avg(t1, t2, wt, bdmax))
        m = rng.integers(0, 65, size=(h, w)).astype(np.uint8)
        call(R.mask, [d.ctypes.data, d.strides[0], t1.ctypes.data, t2.ctypes.data, w, h, m.ctypes.data], hbd, bdmax)
        assert np.array_equal(d, P.mask(t1, t2, m.astype(np.int64), bdmax))
    for w, h in ((4, 4), (8, 16), (16, 8), (32, 32)):
        dst = rng.integers(0, bdmax + 1, size=(h, w)).astype(pdt(hbd))
        tmp = rng.integers(0, bdmax + 1, size=(h, w)).astype(pdt(hbd))
        m = rng.integers(0, 65, size=(h, w)).astype(np.uint8)
        d = dst.copy()
        R.blend(d.ctypes.data, d.strides[0], tmp.ctypes.data, w, h, m.ctypes.data)
        assert np.array_equal(d, P.blend(dst, tmp, m.astype(np.int64)))
        d = dst.copy()
        R.blend_v(d.ctypes.data, d.strides[0], tmp.ctypes.data, w, h)
        assert np.array_equal(d, P.blend_v(dst, tmp))
        d = dst.copy()
        R.blend_h(d.ctypes.data, d.strides[0], tmp.ctypes.data, w, h)
        assert np.array_equal(d, P.blend_h(dst, tmp))


@pytest.mark.parametrize("bdmax", [0xff, 0x3ff, 0xfff])
def test_intra_predictors(ref, bdmax):
    hbd = bdmax > 0xff
    R = ref.bpc[hbd]
    rng = np.random.default_rng(200 + bdmax)
    table = {0: 0, 1: 1, 2: 2, "dc_left": 3, "dc_top": 4, "dc_128": 5, 9: 9, 10: 10, 11: 11, 12: 12}
    for w in (4, 8, 16, 32, 64):
        for h in (4, 8, 16, 32, 64):
            if max(w, h) > 4 * min(w, h):
                continue
            edge = rng.integers(0, bdmax + 1, size=(2 * 64 + 2 * 64 + 1 + 64,)).astype(pdt(hbd))
            c = 160                                   # edge[c] = top-left, edge[c + 1 ..] = top, edge[c - 1 - y] = left[y]
            top, left, tl = edge[c + 1:c + 1 + w], edge[c - h:c][::-1], edge[c]
            for mode, idx in table.items():
                d = np.zeros((h, w), pdt(hbd))
                call(R.intra_pred[idx], [d.ctypes.data, d.strides[0], edge.ctypes.data + c * edge.itemsize, w, h, 0, 0, 0],
                     hbd, bdmax)
                assert np.array_equal(d, P.ipred(mode, top, left, tl, w, h, bdmax)), (mode, w, h)


@pytest.mark.parametrize("bdmax", [0xff, 0x3ff, 0xfff])
def test_itxfm_add_dct_identity(ref, bdmax):
    hbd = bdmax > 0xff
    R = ref.bpc[hbd]
    rng = np.random.default_rng(300 + bdmax)
    sizes = {0: (4, 4), 1: (8, 8), 2: (16, 16), 5: (4, 8), 6: (8, 4), 7: (8, 16), 8: (16, 8), 13: (4, 16), 14: (16, 4)}
    types = {0: (False, False), 9: (True, True), 10: (True, False), 11: (False, True)}   # txtp: (row id, col id)
    cdt = np.int32 if hbd else np.int16
    for tx, (w, h) in sizes.items():
        for txtp, (rid, cid) in types.items():
            for amp in (bdmax, bdmax * 16):
                coef = rng.integers(-amp, amp + 1, size=(w * h,)).astype(cdt)
                dst = rng.integers(0, bdmax + 1, size=(h, w)).astype(pdt(hbd))
                want = P.itxfm_add(dst, coef, w, h, rid, cid, bdmax)
                d, c = dst.copy(), coef.copy()
                call(R.itxfm_add[tx][txtp], [d.ctypes.data, d.strides[0], c.ctypes.data, w * h - 1], hbd, bdmax)
                assert np.array_equal(d, want), (tx, txtp, amp)
                assert not c.any()                    # the reference zeroes its input (itx_tmpl.c:89)


@pytest.mark.parametrize("bdmax", [0xff, 0x3ff, 0xfff])
def test_w_mask_and_warp(ref, bdmax):
    hbd = bdmax > 0xff
    R = ref.bpc[hbd]
    rng = np.random.default_rng(400 + bdmax)
    ib = P.inter_bits(bdmax)
    lo, hi = -P.prep_bias(bdmax), (bdmax << ib) - P.prep_bias(bdmax)
    for lay, (sh, sv) in enumerate(((0, 0), (1, 0), (1, 1))):          # w_mask[0..2] = 444, 422, 420
        for w, h in ((8, 8), (16, 32), (32, 16), (64, 64)):
            for sign in (0, 1):
                t1 = rng.integers(lo, hi + 1, size=(h, w)).astype(np.int16)
                t2 = (t1 + rng.integers(-300, 301, size=(h, w))).clip(lo, hi).astype(np.int16)
                d = np.zeros((h, w), pdt(hbd))
                m = np.zeros((h >> sv, w >> sh), np.uint8)
                call(R.w_mask[lay], [d.ctypes.data, d.strides[0], t1.ctypes.data, t2.ctypes.data, w, h, m.ctypes.data,
                                     sign], hbd, bdmax)
                px, mk = P.w_mask(t1, t2, sign, sh, sv, bdmax)
                assert np.array_equal(d, px) and np.array_equal(m, mk), (lay, w, h, sign)
    for _ in range(20):
        src = rng.integers(0, bdmax + 1, size=(15, 15)).astype(pdt(hbd))
        sp = src.ctypes.data + (3 * 15 + 3) * src.itemsize
        abcd = (rng.integers(0, 0x2000, size=4) - 0xa00).astype(np.int16)
        mx, my = (int(rng.integers(0, 0x2000)) - 0xa00) & ~0x3f, (int(rng.integers(0, 0x2000)) - 0xa00) & ~0x3f
        d = np.zeros((8, 8), pdt(hbd))
        call(R.warp8x8, [d.ctypes.data, d.strides[0], sp, src.strides[0], abcd.ctypes.data, mx, my], hbd, bdmax)
        assert np.array_equal(d, P.warp8x8(src, [int(v) for v in abcd], mx, my, bdmax))
        t = np.zeros((8, 8), np.int16)
        call(R.warp8x8t, [t.ctypes.data, 8, sp, src.strides[0], abcd.ctypes.data, mx, my], hbd, bdmax)
        assert np.array_equal(t, P.warp8x8(src, [int(v) for v in abcd], mx, my, bdmax, prep=True))


@pytest.mark.parametrize("bdmax", [0xff, 0x3ff, 0xfff])
def test_cfl_and_palette(ref, bdmax):
    hbd = bdmax > 0xff
    R = ref.bpc[hbd]
    rng = np.random.default_rng(500 + bdmax)
    for idx, (sh, sv) in enumerate(((1, 1), (1, 0), (0, 0))):          # cfl_ac[0..2] = 420, 422, 444
        for cw, ch in ((4, 4), (8, 8), (16, 8), (8, 16), (32, 32)):
            w_pad, h_pad = int(rng.integers(0, cw // 4)), int(rng.integers(0, ch // 4))
            luma = rng.integers(0, bdmax + 1, size=(ch << sv, cw << sh)).astype(pdt(hbd))
            ac = np.zeros((ch, cw), np.int16)
            R.cfl_ac[idx](ac.ctypes.data, luma.ctypes.data, luma.strides[0], w_pad, h_pad, cw, ch)
            want = P.cfl_ac(luma, w_pad, h_pad, cw, ch, sh, sv)
            assert np.array_equal(ac, want), (idx, cw, ch, w_pad, h_pad)
            # cfl_pred[0] = DC variant: dc from the edges as in ipred (dc_gen), then the ac term
            edge = rng.integers(0, bdmax + 1, size=(400,)).astype(pdt(hbd))
            c = 160
            top, left = edge[c + 1:c + 1 + cw], edge[c - ch:c][::-1]
            alpha = int(rng.integers(1, 17)) * (1 if rng.integers(2) else -1)
            d = np.zeros((ch, cw), pdt(hbd))
            call(R.cfl_pred[0], [d.ctypes.data, d.strides[0], edge.ctypes.data + c * edge.itemsize, cw, ch,
                                 ac.ctypes.data, alpha], hbd, bdmax)
            dc = int(P.ipred(0, top, left, 0, cw, ch, bdmax)[0, 0])
            assert np.array_equal(d, P.cfl_pred(dc, want, alpha, bdmax)), (idx, cw, ch)
    for w, h in ((4, 4), (8, 16), (64, 64)):
        pal = rng.integers(0, bdmax + 1, size=(8,)).astype(pdt(hbd))
        idx = (rng.integers(0, 256, size=(w * h // 2,)) & 0x77).astype(np.uint8)
        d = np.zeros((h, w), pdt(hbd))
        R.pal_pred(d.ctypes.data, d.strides[0], pal.ctypes.data, idx.ctypes.data, w, h)
        assert np.array_equal(d, P.pal_pred(pal.astype(np.int64), idx, w, h))
