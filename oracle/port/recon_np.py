"""Plain numpy RESTATEMENT of a subset of dav1d's pixel-reconstruction DSP - TEST INFRASTRUCTURE ONLY.

Independent of the CUDA sources and of the compiled reference: written from the algorithm in
the reference's C templates (file:line cited per function).  It is pinned against
oracle/_ref/libdav1d_ref.so (the reference itself, compiled in place) by
tests/test_port.py; where the two exist side by side the compiled reference is the checker
the parity tests use, this file documents the arithmetic in executable form and is the
checker of last resort where /root/reference cannot be compiled.

Covered: put / prep 8-tap + bilinear (all four (mx, my) paths), avg, w_avg, mask, w_mask
(444 / 422 / 420), blend, blend_v, blend_h, warp8x8 / warp8x8t; intra DC family, V, H, Paeth,
smooth / smooth_v / smooth_h, cfl_ac (three layouts), cfl_pred, pal_pred; itxfm_add for the
4/8/16-point DCT and identity transforms (9 block sizes, 4 type combinations).
NOT covered (the compiled reference is the only checker): scaled MC, emu_edge, ADST / flipADST /
WHT / 32- and 64-point transforms, directional + filter-intra prediction, edge preparation.
"""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _table(name, dtype, shape):
    """AV1 constant tables (values only) from the generated include the kernels use
    (dav1d-mirror_b200/csrc/tables_data.inc, made by tools/gen_tables.py from tables.c:443-823)."""
    src = open(os.path.join(ROOT, "dav1d-mirror_b200", "csrc", "tables_data.inc")).read()
    m = re.search(r"#define D1_TBL_%s \{(.*?)\}" % name, src, re.S)
    vals = [int(v) for v in re.findall(r"-?\d+", m.group(1))]
    return np.array(vals, dtype=dtype).reshape(shape)


SUBPEL = _table("SUBPEL_FILTERS", np.int64, (6, 15, 8))
SM_WEIGHTS = _table("SM_WEIGHTS", np.int64, (128,))
OBMC_MASKS = _table("OBMC_MASKS", np.int64, (64,))
WARP_FILTER = _table("WARP_FILTER", np.int64, (193, 8))


def inter_bits(bdmax):           # get_intermediate_bits(), include/common/bitdepth.h:60-76
    return 2 if bdmax > 1023 else 4


def prep_bias(bdmax):            # PREP_BIAS: 0 for 8 bpc, 8192 for 16 bpc (mc_tmpl.c:39-49)
    return 8192 if bdmax > 0xff else 0


def _taps(filter_2d, vertical, frac, dim):
    """8 taps of one direction.  filter_2d = enum Filter2d (levels.h:184-196): index 9 is bilinear;
    otherwise h type = filter_2d's horizontal filter, v type = vertical one (mc_tmpl.c:330-384);
    blocks of at most 4 samples in that direction use the 4-tap sets 3 / 4 (mc_tmpl.c:99-107)."""
    if filter_2d == 9:
        t = np.zeros(8, np.int64)
        t[3], t[4] = 16 - frac, frac
        return t
    h_type = [0, 0, 0, 2, 2, 2, 1, 1, 1][filter_2d]     # REGULAR=0, SMOOTH=1, SHARP=2 per Filter2d row
    v_type = filter_2d % 3
    ty = v_type if vertical else h_type
    s = ty if dim > 4 else 3 + (ty & 1)
    return SUBPEL[s, frac - 1]


def _filt_h(win, taps):          # win[rows, w + 7] -> [rows, w]
    w = win.shape[1] - 7
    return sum(taps[k] * win[:, k:k + w] for k in range(8))


def _filt_v(win, taps):          # win[h + 7, w] -> [h, w]
    h = win.shape[0] - 7
    return sum(taps[k] * win[k:k + h, :] for k in range(8))


def mc_put(src, w, h, mx, my, filter_2d, bdmax):
    """put_8tap_c / put_bilin_c (mc_tmpl.c:113-171, 395-450).  src: int array with the block's
    top-left at [3, 3] (rows/cols -3..+4 present).  Returns the w x h pixels."""
    s = src.astype(np.int64)
    ib = inter_bits(bdmax)
    bs = 4 if filter_2d == 9 else 6                   # bilinear taps sum to 16, 8-tap to 64
    blk = s[3:3 + h, 3:3 + w]
    if mx and my:
        mid = (_filt_h(s[0:h + 7, 0:w + 7], _taps(filter_2d, False, mx, w)) + ((1 << (bs - ib)) >> 1)) >> (bs - ib)
        out = (_filt_v(mid, _taps(filter_2d, True, my, h)) + ((1 << (bs + ib)) >> 1)) >> (bs + ib)
    elif mx:
        # folded two-stage rounding: (sum + 32 + ((1 << (6 - ib)) >> 1)) >> 6  (mc_tmpl.c:120,153; bilinear :431-433)
        rnd = (1 << (bs - 1)) + ((1 << (bs - ib)) >> 1)
        out = (_filt_h(s[3:3 + h, 0:w + 7], _taps(filter_2d, False, mx, w)) + rnd) >> bs
    elif my:
        out = (_filt_v(s[0:h + 7, 3:3 + w], _taps(filter_2d, True, my, h)) + ((1 << bs) >> 1)) >> bs
    else:
        out = blk
    return np.clip(out, 0, bdmax)


def mc_prep(src, w, h, mx, my, filter_2d, bdmax):
    """prep_8tap_c / prep_bilin_c (mc_tmpl.c:223-282, 493-546): int16 intermediates."""
    s = src.astype(np.int64)
    ib = inter_bits(bdmax)
    bs = 4 if filter_2d == 9 else 6
    if mx and my:
        mid = (_filt_h(s[0:h + 7, 0:w + 7], _taps(filter_2d, False, mx, w)) + ((1 << (bs - ib)) >> 1)) >> (bs - ib)
        out = (_filt_v(mid, _taps(filter_2d, True, my, h)) + ((1 << bs) >> 1)) >> bs
    elif mx:
        out = (_filt_h(s[3:3 + h, 0:w + 7], _taps(filter_2d, False, mx, w)) + ((1 << (bs - ib)) >> 1)) >> (bs - ib)
    elif my:
        out = (_filt_v(s[0:h + 7, 3:3 + w], _taps(filter_2d, True, my, h)) + ((1 << (bs - ib)) >> 1)) >> (bs - ib)
    else:
        out = s[3:3 + h, 3:3 + w] << ib
    return out - prep_bias(bdmax)


def avg(t1, t2, bdmax):                      # avg_c, mc_tmpl.c:587-602
    ib = inter_bits(bdmax)
    return np.clip((t1.astype(np.int64) + t2 + (1 << ib) + 2 * prep_bias(bdmax)) >> (ib + 1), 0, bdmax)


def w_avg(t1, t2, weight, bdmax):            # w_avg_c, mc_tmpl.c:604-620
    ib = inter_bits(bdmax)
    s = t1.astype(np.int64) * weight + t2.astype(np.int64) * (16 - weight)
    return np.clip((s + (8 << ib) + 16 * prep_bias(bdmax)) >> (ib + 4), 0, bdmax)


def mask(t1, t2, m, bdmax):                  # mask_c, mc_tmpl.c:622-639
    ib = inter_bits(bdmax)
    s = t1.astype(np.int64) * m + t2.astype(np.int64) * (64 - m)
    return np.clip((s + (32 << ib) + 64 * prep_bias(bdmax)) >> (ib + 6), 0, bdmax)


def blend(dst, tmp, m):                      # blend_c, mc_tmpl.c:642-653
    return (dst.astype(np.int64) * (64 - m) + tmp.astype(np.int64) * m + 32) >> 6


def blend_v(dst, tmp):                       # blend_v_c, mc_tmpl.c:655-666: first 3w/4 columns
    h, w = dst.shape
    out = dst.astype(np.int64).copy()
    n = (w * 3) >> 2
    out[:, :n] = blend(dst[:, :n], tmp[:, :n], OBMC_MASKS[w:w + n][None, :])
    return out


def blend_h(dst, tmp):                       # blend_h_c, mc_tmpl.c:668-681: first 3h/4 rows
    h, w = dst.shape
    out = dst.astype(np.int64).copy()
    n = (h * 3) >> 2
    out[:n] = blend(dst[:n], tmp[:n], OBMC_MASKS[h:h + n][:, None])
    return out


def w_mask(t1, t2, sign, ss_hor, ss_ver, bdmax):
    """w_mask_c (mc_tmpl.c:683-728): blend with a mask derived from |t1 - t2|; returns (pixels, mask at the
    chroma resolution of the layout: 444 m, 422 (m + n + 1 - sign) >> 1, 420 (sum of four + 2 - sign) >> 2)."""
    ib = inter_bits(bdmax)
    bitdepth = {0xff: 8, 0x3ff: 10, 0xfff: 12}[bdmax]
    a, b = t1.astype(np.int64), t2.astype(np.int64)
    mask_sh = bitdepth + ib - 4
    m = np.minimum(38 + ((np.abs(a - b) + (1 << (mask_sh - 5))) >> mask_sh), 64)
    px = np.clip((a * m + b * (64 - m) + (32 << ib) + 64 * prep_bias(bdmax)) >> (ib + 6), 0, bdmax)
    if not ss_hor:
        return px, m
    pair = m[:, 0::2] + m[:, 1::2]
    if not ss_ver:
        return px, (pair + 1 - sign) >> 1
    return px, (pair[0::2] + pair[1::2] + 2 - sign) >> 2


def warp8x8(src, abcd, mx, my, bdmax, prep=False):
    """warp_affine_8x8_c / warp_affine_8x8t_c (mc_tmpl.c:758-825).  src: int array with the block's
    top-left at [3, 3] (15 x 15 window).  Per-pixel filters from dav1d_mc_warp_filter[64 + ((t + 512) >> 10)]."""
    s = src.astype(np.int64)
    ib = inter_bits(bdmax)
    mid = np.zeros((15, 8), np.int64)
    for y in range(15):
        for x in range(8):
            f = WARP_FILTER[64 + ((mx + y * abcd[1] + x * abcd[0] + 512) >> 10)]
            mid[y, x] = (int((f * s[y, x:x + 8]).sum()) + ((1 << (7 - ib)) >> 1)) >> (7 - ib)
    out = np.zeros((8, 8), np.int64)
    sh = 7 if prep else 7 + ib
    for y in range(8):
        for x in range(8):
            f = WARP_FILTER[64 + ((my + y * abcd[3] + x * abcd[2] + 512) >> 10)]
            out[y, x] = (int((f * mid[y:y + 8, x]).sum()) + ((1 << sh) >> 1)) >> sh
    return out - prep_bias(bdmax) if prep else np.clip(out, 0, bdmax)


# ------------------------------------------------------------------ intra prediction
def _dc_gen(s, w, h, hbd):
    """dc_gen (ipred_tmpl.c:140-166): (sum + (w + h) / 2) / (w + h) with the reference's
    multiply-shift division for the 1:2 and 1:4 shapes."""
    dc = int(s) + ((w + h) >> 1)
    dc >>= int(np.log2(min(w, h))) + 1 if w == h else int(np.log2(min(w, h)))
    if w != h:
        r = max(w, h) // min(w, h)
        m12, m14, sh = (0xAAAB, 0x6667, 17) if hbd else (0x5556, 0x3334, 16)
        dc = (dc * (m12 if r == 2 else m14)) >> sh
    return dc


def ipred(mode, top, left, topleft, w, h, bdmax):
    """intra_pred[mode] for DC (0), V (1), H (2), SMOOTH (9), SMOOTH_V (10), SMOOTH_H (11), PAETH (12),
    LEFT_DC (13... here 'dc_left'), TOP_DC ('dc_top'), DC_128 ('dc_128')  (ipred_tmpl.c:86-310).
    top[w], left[h] (left[0] = the row-0 neighbour), topleft scalar."""
    top = top.astype(np.int64)
    left = left.astype(np.int64)
    hbd = bdmax > 0xff
    if mode == 0:
        return np.full((h, w), _dc_gen(top.sum() + left.sum(), w, h, hbd), np.int64)
    if mode == "dc_top":
        return np.full((h, w), (int(top.sum()) + (w >> 1)) >> int(np.log2(w)), np.int64)
    if mode == "dc_left":
        return np.full((h, w), (int(left.sum()) + (h >> 1)) >> int(np.log2(h)), np.int64)
    if mode == "dc_128":
        return np.full((h, w), (bdmax + 1) >> 1, np.int64)
    if mode == 1:
        return np.tile(top[None, :], (h, 1))
    if mode == 2:
        return np.tile(left[:, None], (1, w))
    if mode == 12:                               # paeth, ipred_tmpl.c:249-268 (tie-break order: left, top, topleft)
        T, Lf, tl = top[None, :], left[:, None], int(topleft)
        base = Lf + T - tl
        ld, td, tld = np.abs(Lf - base), np.abs(T - base), np.abs(tl - base)
        return np.where((ld <= td) & (ld <= tld), Lf, np.where(td <= tld, T, tl))
    wh, wv = SM_WEIGHTS[w:2 * w], SM_WEIGHTS[h:2 * h]
    right, bottom = int(top[w - 1]), int(left[h - 1])
    if mode == 9:                                # smooth, ipred_tmpl.c:270-291
        p = (wv[:, None] * top[None, :] + (256 - wv[:, None]) * bottom +
             wh[None, :] * left[:, None] + (256 - wh[None, :]) * right)
        return (p + 256) >> 9
    if mode == 10:                               # smooth_v, :293-309
        return (wv[:, None] * top[None, :] + (256 - wv[:, None]) * bottom + 128) >> 8
    if mode == 11:                               # smooth_h, :311-327
        return (wh[None, :] * left[:, None] + (256 - wh[None, :]) * right + 128) >> 8
    raise ValueError(mode)


def cfl_ac(luma, w_pad, h_pad, cw, ch, ss_hor, ss_ver):
    """cfl_ac_c (ipred_tmpl.c:657-703): subsampled luma * 8 over the un-padded area, edge replication into
    the padded part, minus the rounded mean.  luma: co-located luma block ((ch << ss_ver) x (cw << ss_hor))."""
    y = luma.astype(np.int64)
    if ss_hor:
        y = y[:, 0::2] + y[:, 1::2]
    if ss_ver:
        y = y[0::2] + y[1::2]
    ac = y[:ch, :cw] << (1 + (not ss_ver) + (not ss_hor))
    vw, vh = cw - 4 * w_pad, ch - 4 * h_pad
    ac[:, vw:] = ac[:, vw - 1:vw]
    ac[vh:] = ac[vh - 1:vh]
    log2sz = int(np.log2(cw)) + int(np.log2(ch))
    return ac - ((int(ac.sum()) + ((1 << log2sz) >> 1)) >> log2sz)


def cfl_pred(dc, ac, alpha, bdmax):
    """cfl_pred (ipred_tmpl.c:71-84): dc + sign(alpha * ac) * ((|alpha * ac| + 32) >> 6)."""
    diff = alpha * ac.astype(np.int64)
    return np.clip(dc + np.sign(diff) * ((np.abs(diff) + 32) >> 6), 0, bdmax)


def pal_pred(pal, idx, w, h):
    """pal_pred_c (ipred_tmpl.c:717-730): two 3-bit indices per byte, low nibble first."""
    i = idx[:w * h // 2].astype(np.int64)
    out = np.empty(w * h, np.int64)
    out[0::2], out[1::2] = pal[i & 7], pal[i >> 4]
    return out.reshape(h, w)


# ------------------------------------------------------------------ inverse transforms (subset)
def _clip(v, lo, hi):
    return np.clip(v, lo, hi)


def _dct4(c, lo, hi):
    """inv_dct4_1d_internal_c, itx_1d.c:65-86; c: [4, n] (n independent vectors)."""
    in0, in1, in2, in3 = c
    t0 = ((in0 + in2) * 181 + 128) >> 8
    t1 = ((in0 - in2) * 181 + 128) >> 8
    t2 = ((in1 * 1567 - in3 * (3784 - 4096) + 2048) >> 12) - in3
    t3 = ((in1 * (3784 - 4096) + in3 * 1567 + 2048) >> 12) + in1
    return np.stack([_clip(t0 + t3, lo, hi), _clip(t1 + t2, lo, hi), _clip(t1 - t2, lo, hi), _clip(t0 - t3, lo, hi)])


def _dct8(c, lo, hi):
    """inv_dct8_1d_internal_c, itx_1d.c:93-135."""
    e = _dct4(c[0::2], lo, hi)
    in1, in3, in5, in7 = c[1], c[3], c[5], c[7]
    t4a = ((in1 * 799 - in7 * (4017 - 4096) + 2048) >> 12) - in7
    t5a = (in5 * 1703 - in3 * 1138 + 1024) >> 11
    t6a = (in5 * 1138 + in3 * 1703 + 1024) >> 11
    t7a = ((in1 * (4017 - 4096) + in7 * 799 + 2048) >> 12) + in1
    t4, t5a_ = _clip(t4a + t5a, lo, hi), _clip(t4a - t5a, lo, hi)
    t7, t6a_ = _clip(t7a + t6a, lo, hi), _clip(t7a - t6a, lo, hi)
    t5 = ((t6a_ - t5a_) * 181 + 128) >> 8
    t6 = ((t6a_ + t5a_) * 181 + 128) >> 8
    return np.stack([_clip(e[0] + t7, lo, hi), _clip(e[1] + t6, lo, hi), _clip(e[2] + t5, lo, hi),
                     _clip(e[3] + t4, lo, hi), _clip(e[3] - t4, lo, hi), _clip(e[2] - t5, lo, hi),
                     _clip(e[1] - t6, lo, hi), _clip(e[0] - t7, lo, hi)])


def _dct16(c, lo, hi):
    """inv_dct16_1d_internal_c, itx_1d.c:142-232."""
    e = _dct8(c[0::2], lo, hi)
    in1, in3, in5, in7, in9, in11, in13, in15 = c[1::2]
    t8a = ((in1 * 401 - in15 * (4076 - 4096) + 2048) >> 12) - in15
    t9a = (in9 * 1583 - in7 * 1299 + 1024) >> 11
    t10a = ((in5 * 1931 - in11 * (3612 - 4096) + 2048) >> 12) - in11
    t11a = ((in13 * (3920 - 4096) - in3 * 1189 + 2048) >> 12) + in13
    t12a = ((in13 * 1189 + in3 * (3920 - 4096) + 2048) >> 12) + in3
    t13a = ((in5 * (3612 - 4096) + in11 * 1931 + 2048) >> 12) + in5
    t14a = (in9 * 1299 + in7 * 1583 + 1024) >> 11
    t15a = ((in1 * (4076 - 4096) + in15 * 401 + 2048) >> 12) + in1
    t8, t9 = _clip(t8a + t9a, lo, hi), _clip(t8a - t9a, lo, hi)
    t10, t11 = _clip(t11a - t10a, lo, hi), _clip(t11a + t10a, lo, hi)
    t12, t13 = _clip(t12a + t13a, lo, hi), _clip(t12a - t13a, lo, hi)
    t14, t15 = _clip(t15a - t14a, lo, hi), _clip(t15a + t14a, lo, hi)
    t9a = ((t14 * 1567 - t9 * (3784 - 4096) + 2048) >> 12) - t9
    t14a = ((t14 * (3784 - 4096) + t9 * 1567 + 2048) >> 12) + t14
    t10a = ((-(t13 * (3784 - 4096) + t10 * 1567) + 2048) >> 12) - t13
    t13a = ((t13 * 1567 - t10 * (3784 - 4096) + 2048) >> 12) - t10
    t8a, t9 = _clip(t8 + t11, lo, hi), _clip(t9a + t10a, lo, hi)
    t10, t11a = _clip(t9a - t10a, lo, hi), _clip(t8 - t11, lo, hi)
    t12a, t13 = _clip(t15 - t12, lo, hi), _clip(t14a - t13a, lo, hi)
    t14, t15a = _clip(t14a + t13a, lo, hi), _clip(t15 + t12, lo, hi)
    t10a = ((t13 - t10) * 181 + 128) >> 8
    t13a = ((t13 + t10) * 181 + 128) >> 8
    t11 = ((t12a - t11a) * 181 + 128) >> 8
    t12 = ((t12a + t11a) * 181 + 128) >> 8
    o = [t15a, t14, t13a, t12, t11, t10a, t9, t8a]
    return np.stack([_clip(e[i] + o[i], lo, hi) for i in range(8)] +
                    [_clip(e[7 - i] - o[7 - i], lo, hi) for i in range(8)])


def _identity(c, n):
    """inv_identity{4,8,16}_1d_c, itx_1d.c:930-962."""
    if n == 4:
        return c + ((c * 1697 + 2048) >> 12)
    if n == 8:
        return c * 2
    return 2 * c + ((c * 1697 + 1024) >> 11)


_TX1D = {4: _dct4, 8: _dct8, 16: _dct16}


def itxfm_add(dst, coef, w, h, row_identity, col_identity, bdmax):
    """inv_txfm_add_c (itx_tmpl.c:40-140) for w, h in {4, 8, 16}, DCT or identity per direction, without
    the dc-only shortcut (that shortcut is bit-identical by construction only for DCT_DCT with eob 0;
    callers pass eob > 0 cases).  coef: column-major h x w block as the reference stores it
    (coef[y + x * h]); dst: [h, w] pixels.  Returns dst + residual, clipped."""
    c = np.asarray(coef, np.int64).reshape(w, h).T.copy()          # c[y, x]
    hbd = bdmax > 0xff
    row_lo = -((bdmax + 1) << 7) if hbd else -32768                 # itx_tmpl.c:50-62
    col_lo = -((bdmax + 1) << 5) if hbd else -32768
    row_hi, col_hi = ~row_lo, ~col_lo
    is_rect2 = w * 2 == h or h * 2 == w
    shift = {(4, 4): 0, (4, 8): 0, (8, 4): 0, (16, 16): 2}.get((w, h), 1)   # itx_tmpl.c:142-160
    if is_rect2:
        c = (c * 181 + 128) >> 8
    rows = c.T                                                      # [w, h]: element k of every row
    rows = _identity(rows, w) if row_identity else _TX1D[w](rows, row_lo, row_hi)
    t = np.clip((rows + ((1 << shift) >> 1)) >> shift, col_lo, col_hi)      # [w(x), h(y)]
    cols = t.T                                                      # [h, w]: element k of every column
    cols = _identity(cols, h) if col_identity else _TX1D[h](cols, col_lo, col_hi)
    return np.clip(dst.astype(np.int64) + ((cols + 8) >> 4), 0, bdmax)
