"""Coefficient generation for inverse-transform tests, following the
reference's checkasm strategy (tests/checkasm/itx.c:131-240): a
double-precision forward transform of a random residual in
[-bitdepth_max, bitdepth_max], scaled per size and rounded, then truncated in
scan order at a random eob that confines the non-zero coefficients to the
top-left (8*subsh)^2 corner."""
import math

import numpy as np

import _d1pkg

pkg = _d1pkg.load_pkg()
TX_DIMS = pkg.TX_DIMS

DCT, ADST, FLIPADST, IDENTITY, WHT = range(5)
# txtp -> (first index used by ftx for BOTH passes, second) as in itx.c:46-64
ITX_1D_TYPES = [(DCT, DCT), (DCT, ADST), (ADST, DCT), (ADST, ADST), (DCT, FLIPADST),
                (FLIPADST, DCT), (FLIPADST, FLIPADST), (FLIPADST, ADST), (ADST, FLIPADST),
                (IDENTITY, IDENTITY), (IDENTITY, DCT), (DCT, IDENTITY), (IDENTITY, ADST),
                (ADST, IDENTITY), (IDENTITY, FLIPADST), (FLIPADST, IDENTITY), (WHT, WHT)]
NAMES = ["dct", "adst", "flipadst", "identity", "wht"]
TX_CLASS = [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1, 2, 1, 2, 1, 0]  # tables.c:305-323 (2D=0,H=1,V=2)

SCALING = [4.0, 4.0 * math.sqrt(0.5), 2.0, 2.0 * math.sqrt(0.5), 1.0, 0.5 * math.sqrt(0.5),
           0.25, 0.125 * math.sqrt(0.5), 0.0625]


def valid_types(tx):
    """Populated itxfm_add slots per size (reference src/itx_tmpl.c:248-268)."""
    w, h = TX_DIMS[tx]
    m = max(w, h)
    if m == 64:
        return [0]
    if m == 32:
        return [0, 9]
    if w == 16 and h == 16:
        return [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]
    t = list(range(16))
    if tx == 0:
        t.append(16)
    return t


def _basis(kind, n):
    if kind == DCT:
        j = np.arange(n)[None, :]
        i = np.arange(n)[:, None]
        m = np.cos(math.pi * (2 * j + 1) * i / (2.0 * n))
        m[0] *= math.sqrt(0.5)
        return m
    if kind in (ADST, FLIPADST):
        j = np.arange(n)[None, :]
        i = np.arange(n)[:, None]
        if n == 4:
            return np.sin(math.pi * (j + 1) * (2 * i + 1) / 9.0)
        return np.sin(math.pi * (2 * j + 1) * (2 * i + 1) / (4.0 * n))
    if kind == IDENTITY:
        return np.eye(n)
    # forward WHT4 (itx.c:113-125)
    m = np.zeros((4, 4))
    for k in range(4):
        e = np.zeros(4)
        e[k] = 1.0
        t0 = e[0] + e[1]
        t3 = e[3] - e[2]
        t4 = (t0 - t3) * 0.5
        t1 = t4 - e[1]
        t2 = t4 - e[2]
        m[:, k] = [t0 - t2, t2, t3 + t1, t1]
    return m


def ftx(rng, tx, txtp, subsh, bitdepth_max, scan):
    """Returns (coef[sw*sh] int64 column-major as the decoder stores them, eob)."""
    w, h = TX_DIMS[tx]
    sw, sh = min(w, 32), min(h, 32)
    kind = ITX_1D_TYPES[txtp][0]
    scale = SCALING[int(math.log2(w * h)) - 4]
    resid = rng.integers(-bitdepth_max, bitdepth_max + 1, size=(h, w)).astype(np.float64)
    t = (resid @ _basis(kind, w).T) * scale          # rows transformed: [h][w]
    out = _basis(kind, h) @ t                        # columns transformed: out[y][x] -> stored [x*h + y]
    # reference layout: out[i*h + y] for column i; buf[y*sw + x] = out[y*w + x] with (y<sh, x<sw)
    flat = out.T.reshape(-1)                         # flat[i*h + y]
    buf = np.zeros(sw * sh, dtype=np.int64)
    for y in range(sh):
        buf[y * sw:(y + 1) * sw] = np.floor(flat[y * w:y * w + sw] + 0.5).astype(np.int64)
    return _copy_subcoefs(rng, buf, txtp, sw, sh, subsh, scan)


def _copy_subcoefs(rng, coeff, txtp, sw, sh, subsh, scan):
    cls = TX_CLASS[txtp]
    sub_high = subsh * 8 - 1 if subsh > 0 else 0
    sub_low = sub_high - 8 if subsh > 1 else 0
    eob = 0
    n = 0
    while n < sw * sh:
        if cls == 0:
            rc = int(scan[n]); rcx, rcy = rc % sh, rc // sh
        elif cls == 1:
            rcx, rcy = n % sh, n // sh
        else:
            rcx, rcy = n // sw, n % sw
        if rcx > sub_high or rcy > sub_high:
            break
        elif not eob and (rcx > sub_low or rcy > sub_low):
            eob = n
        n += 1
    if eob:
        eob += int(rng.integers(0, 1 << 30)) % (n - eob - 1)
    if cls == 0:
        coeff[scan[eob + 1:sw * sh].astype(np.int64)] = 0
    elif cls == 1:
        coeff[eob + 1:] = 0
    else:
        rcx, rcy = eob // sw, eob % sw
        while rcx < sh:
            rcy += 1
            while rcy < sw:
                coeff[rcy * sh + rcx] = 0
                rcy += 1
            rcx += 1
            rcy = -1
    return coeff, eob
