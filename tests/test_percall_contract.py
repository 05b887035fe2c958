"""Per-call (DSP-table) surface: the parts of the reference contract that are not about arithmetic.

 * negative strides - dav1d's --negstride option (tools/dav1d_cli_parse.c:89,151) hands the DSP
   functions pictures whose rows go DOWN in memory: `ptr` is the first row, `stride` < 0;
 * re-entrancy - the tables are shared by up to DAV1D_MAX_THREADS (256) worker threads that call
   concurrently with disjoint outputs (SURVEY.md section 8b "Threading").
Every result is compared with the reference's C templates on the same inputs."""
import threading

import numpy as np
import pytest

from test_mc import call, pdt


def flipped(a):
    """(pointer to row 0, negative stride) of a copy of `a` stored bottom-up; returns the store too."""
    store = np.ascontiguousarray(a[::-1])
    row = store.shape[1] * store.itemsize
    return store, store.ctypes.data + (store.shape[0] - 1) * row, -row


@pytest.mark.gpu
@pytest.mark.parametrize("hbd", [False, True])
def test_negative_strides(ref, cuda, hbd):
    rng = np.random.default_rng(77 + hbd)
    R, G = ref.bpc[hbd], cuda.bpc[hbd]
    bdmax = 0x3ff if hbd else 0xff
    dt = pdt(hbd)
    w, h = 16, 8
    src = rng.integers(0, bdmax + 1, size=(h + 7, w + 16)).astype(dt)
    dst0 = rng.integers(0, bdmax + 1, size=(h, w + 8)).astype(dt)

    def both(fn_of, args_of, outs_of, has_bd=True):
        res = []
        for T in (R, G):
            keep = []
            args = args_of(keep)
            if has_bd:
                call(fn_of(T), args, hbd, bdmax)
            else:
                fn_of(T)(*args)             # no bitdepth_max argument (blend, src/mc.h:93-96)
            res.append([o[::-1].copy() for o in outs_of(keep)])
        for a, b in zip(*res):
            assert np.array_equal(a, b)

    # mc put 8-tap (hv): source and destination both bottom-up
    def mc_args(keep):
        s, sp, ss = flipped(src)
        d, dp, ds = flipped(dst0)
        keep += [s, d]
        return [dp, ds, sp + 3 * ss + 3 * s.itemsize, ss, w, h, 5, 9]
    both(lambda T: T.mc[0], mc_args, lambda keep: [keep[1]])

    # blend (dst read-modify-write), tmp dense
    tmp = rng.integers(0, bdmax + 1, size=(h, w)).astype(dt)
    mask = rng.integers(0, 65, size=(h, w)).astype(np.uint8)

    def blend_args(keep):
        d, dp, ds = flipped(dst0)
        keep += [d]
        return [dp, ds, tmp.ctypes.data, w, h, mask.ctypes.data]
    both(lambda T: T.blend, blend_args, lambda keep: [keep[0]], has_bd=False)

    # itxfm_add 8x8 DCT_DCT
    coef_dt = np.int32 if hbd else np.int16
    coef0 = np.zeros(64, dtype=coef_dt)
    coef0[:10] = rng.integers(-200, 200, size=10)

    def itx_args(keep):
        d, dp, ds = flipped(dst0[:8, :8].copy())
        cf = coef0.copy()
        keep += [d, cf]
        return [dp, ds, cf.ctypes.data, 9]
    both(lambda T: T.itxfm_add[1][0], itx_args, lambda keep: [keep[0], keep[1][None, :][::-1]])

    # intra prediction (PAETH), 8x8
    edge = rng.integers(0, bdmax + 1, size=64).astype(dt)

    def ipred_args(keep):
        d, dp, ds = flipped(dst0[:8, :8].copy())
        keep += [d]
        return [dp, ds, edge.ctypes.data + 32 * edge.itemsize, 8, 8, 0, 8, 8]
    both(lambda T: T.intra_pred[12], ipred_args, lambda keep: [keep[0]])


@pytest.mark.gpu
def test_concurrent_calls_from_16_threads(ref, cuda):
    """16 threads call different functions of the 8- and 16-bit tables at the same time, each on
    its own buffers; every result must equal the reference's."""
    errors = []

    def worker(k):
        try:
            hbd = bool(k & 1)
            rng = np.random.default_rng(1000 + k)
            R, G = ref.bpc[hbd], cuda.bpc[hbd]
            bdmax = 0x3ff if hbd else 0xff
            dt = pdt(hbd)
            for it in range(12):
                w = int(rng.choice([4, 8, 16, 32]))
                h = int(rng.choice([4, 8, 16, 32]))
                src = rng.integers(0, bdmax + 1, size=(h + 7, w + 7)).astype(dt)
                sp = src.ctypes.data + (3 * (w + 7) + 3) * src.itemsize
                # (the two tables must see the same sub-pel phase: draw it once)
                mxy = (int(rng.integers(0, 16)), int(rng.integers(0, 16)))
                outs = []
                for T in (R, G):
                    d = np.zeros((h, w), dtype=dt)
                    call(T.mc[(k + it) % 10], [d.ctypes.data, w * d.itemsize, sp, (w + 7) * src.itemsize, w, h,
                                               mxy[0], mxy[1]], hbd, bdmax)
                    outs.append(d)
                if not np.array_equal(outs[0], outs[1]):
                    errors.append(f"thread {k} iteration {it}: mc mismatch")
                # an intra predictor on the same thread
                edge = rng.integers(0, bdmax + 1, size=160).astype(dt)
                outs = []
                for T in (R, G):
                    d = np.zeros((h, w), dtype=dt)
                    call(T.intra_pred[9 + (k % 4)], [d.ctypes.data, w * d.itemsize,
                                                      edge.ctypes.data + 80 * edge.itemsize, w, h, 0, w, h], hbd, bdmax)
                    outs.append(d)
                if not np.array_equal(outs[0], outs[1]):
                    errors.append(f"thread {k} iteration {it}: ipred mismatch")
        except Exception as e:       # noqa: BLE001
            errors.append(f"thread {k}: {e!r}")

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(16)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:5]
