"""Inverse-transform parity: every populated itxfm_add slot (19 sizes x valid
types x eob classes x 8/10/12 bit) of the CUDA DSP table against the
reference's C template, bit-exact on dst and on the (zeroed) coefficient
buffer - the contract of tests/checkasm/itx.c:243-301."""
import ctypes as C

import numpy as np
import pytest

import itxgen
from itxgen import TX_DIMS, valid_types

SUBSH_ITERS = [2, 2, 3, 5, 5]


def _cases(bpcs):
    for tx in range(19):
        w, h = TX_DIMS[tx]
        lmax = max(int(np.log2(w)), int(np.log2(h))) - 2
        for txtp in valid_types(tx):
            for subsh in range(SUBSH_ITERS[lmax]):
                for bpc in bpcs:
                    yield tx, txtp, subsh, bpc


def _run_one(ref, cuda, rng, tx, txtp, subsh, bpc):
    hbd = bpc > 8
    w, h = TX_DIMS[tx]
    sw, sh = min(w, 32), min(h, 32)
    bdmax = (1 << bpc) - 1
    cdt = np.int32 if hbd else np.int16
    pdt = np.uint16 if hbd else np.uint8
    coef, eob = itxgen.ftx(rng, tx, txtp, subsh, bdmax, ref.scan(tx, sw * sh))
    cbuf = np.zeros(32 * 32, dtype=cdt)
    cbuf[:sw * sh] = coef.astype(cdt)
    cbuf[sw * sh:] = rng.integers(-1000, 1000, size=32 * 32 - sw * sh).astype(cdt)
    stride_px = 64 + 32
    dst = rng.integers(0, bdmax + 1, size=(h + 16, stride_px)).astype(pdt)
    outs = []
    for tbl in (ref.bpc[hbd], cuda.bpc[hbd]):
        d = dst.copy()
        c = cbuf.copy()
        fn = tbl.itxfm_add[tx][txtp]
        assert fn is not None, (tx, txtp)
        args = [d.ctypes.data + (8 * stride_px + 8) * d.itemsize, stride_px * d.itemsize,
                c.ctypes.data, eob]
        if hbd:
            args.append(bdmax)
        fn(*args)
        outs.append((d, c))
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("bpcs", [(8,), (10, 12)])
def test_itxfm_add_all_slots(ref, cuda, bpcs):
    rng = np.random.default_rng(1234 + bpcs[0])
    n = 0
    for tx, txtp, subsh, bpc in _cases(bpcs):
        (rd, rc), (gd, gc) = _run_one(ref, cuda, rng, tx, txtp, subsh, bpc)
        assert np.array_equal(rd, gd), f"dst mismatch tx={tx} txtp={txtp} subsh={subsh} bpc={bpc}"
        assert np.array_equal(rc, gc), f"coef mismatch tx={tx} txtp={txtp} subsh={subsh} bpc={bpc}"
        n += 1
    assert n > 400
    import _d1pkg
    _d1pkg.load_pkg().check_error()


@pytest.mark.gpu
def test_itx_table_population(ref, cuda):
    """The CUDA init overrides exactly the slots the reference populates (itx_tmpl.c:248-268)."""
    for hbd in (False, True):
        for tx in range(19):
            for txtp in range(17):
                assert (ref.bpc[hbd].itxfm_add[tx][txtp] is None) == \
                       (cuda.bpc[hbd].itxfm_add[tx][txtp] is None), (tx, txtp)


def test_itx_generator_and_reference_smoke(ref):
    """CPU-only: the generator yields in-range coefficients and the reference runs on them."""
    rng = np.random.default_rng(7)
    for tx in (0, 2, 4, 9, 13):
        w, h = TX_DIMS[tx]
        sw, sh = min(w, 32), min(h, 32)
        coef, eob = itxgen.ftx(rng, tx, 0, 1, 255, ref.scan(tx, sw * sh))
        assert 0 <= eob < sw * sh
        assert np.abs(coef).max() < 32768
        cbuf = coef.astype(np.int16)
        dst = np.full((h, w), 128, dtype=np.uint8)
        ref.bpc[False].itxfm_add[tx][0](dst.ctypes.data, w, cbuf.ctypes.data, eob)
        assert not cbuf.any()
