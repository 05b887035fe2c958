"""ctypes access to oracle/_ref/libdav1d_ref.so = the reference's own C
templates (src/{mc,itx,ipred,ipred_prepare}_tmpl.c, itx_1d.c, tables.c)
compiled in place by oracle/Makefile.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libdav1d_ref.so")

import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()


def _fn(ptr, restype, *argtypes):
    return C.CFUNCTYPE(restype, *argtypes)(ptr) if ptr else None


# signatures of the DSP function pointers (reference src/mc.h, src/itx.h, src/ipred.h)
P, SZ, I = C.c_void_p, C.c_ssize_t, C.c_int


def sig_mc(hbd): return [P, SZ, P, SZ, I, I, I, I] + ([I] if hbd else [])
def sig_mc_scaled(hbd): return [P, SZ, P, SZ, I, I, I, I, I, I] + ([I] if hbd else [])
def sig_mct(hbd): return [P, P, SZ, I, I, I, I] + ([I] if hbd else [])
def sig_mct_scaled(hbd): return [P, P, SZ, I, I, I, I, I, I] + ([I] if hbd else [])
def sig_avg(hbd): return [P, SZ, P, P, I, I] + ([I] if hbd else [])
def sig_w_avg(hbd): return [P, SZ, P, P, I, I, I] + ([I] if hbd else [])
def sig_mask(hbd): return [P, SZ, P, P, I, I, P] + ([I] if hbd else [])
def sig_w_mask(hbd): return [P, SZ, P, P, I, I, P, I] + ([I] if hbd else [])
def sig_blend(hbd): return [P, SZ, P, I, I, P]
def sig_blend_dir(hbd): return [P, SZ, P, I, I]
def sig_warp(hbd): return [P, SZ, P, SZ, P, I, I] + ([I] if hbd else [])
def sig_emu_edge(hbd): return [C.c_ssize_t] * 6 + [P, SZ, P, SZ]
def sig_resize(hbd): return [P, SZ, P, SZ, I, I, I, I, I] + ([I] if hbd else [])
def sig_itx(hbd): return [P, SZ, P, I] + ([I] if hbd else [])
def sig_ipred(hbd): return [P, SZ, P, I, I, I, I, I] + ([I] if hbd else [])
def sig_cfl_ac(hbd): return [P, P, SZ, I, I, I, I]
def sig_cfl_pred(hbd): return [P, SZ, P, I, I, P, I] + ([I] if hbd else [])
def sig_pal_pred(hbd): return [P, SZ, P, P, I, I]


class DSPTables:
    """Typed callables over the three DSP structs for one bit-depth class."""

    def __init__(self, mc, itx, ipred, hbd):
        self.hbd = hbd
        self.raw_mc, self.raw_itx, self.raw_ipred = mc, itx, ipred
        f = lambda p, s: _fn(p, None, *s(hbd))
        self.mc = [f(mc.mc[i], sig_mc) for i in range(10)]
        self.mc_scaled = [f(mc.mc_scaled[i], sig_mc_scaled) for i in range(10)]
        self.mct = [f(mc.mct[i], sig_mct) for i in range(10)]
        self.mct_scaled = [f(mc.mct_scaled[i], sig_mct_scaled) for i in range(10)]
        self.avg = f(mc.avg, sig_avg)
        self.w_avg = f(mc.w_avg, sig_w_avg)
        self.mask = f(mc.mask, sig_mask)
        self.w_mask = [f(mc.w_mask[i], sig_w_mask) for i in range(3)]
        self.blend = f(mc.blend, sig_blend)
        self.blend_v = f(mc.blend_v, sig_blend_dir)
        self.blend_h = f(mc.blend_h, sig_blend_dir)
        self.warp8x8 = f(mc.warp8x8, sig_warp)
        self.warp8x8t = f(mc.warp8x8t, sig_warp)
        self.emu_edge = f(mc.emu_edge, sig_emu_edge)
        self.resize = f(mc.resize, sig_resize)
        self.itxfm_add = [[f(itx.itxfm_add[t][k], sig_itx) for k in range(17)] for t in range(19)]
        self.intra_pred = [f(ipred.intra_pred[i], sig_ipred) for i in range(14)]
        self.cfl_ac = [f(ipred.cfl_ac[i], sig_cfl_ac) for i in range(3)]
        self.cfl_pred = [f(ipred.cfl_pred[i], sig_cfl_pred) for i in range(6)]
        self.pal_pred = f(ipred.pal_pred, sig_pal_pred)


class RefDSP:
    def __init__(self):
        if not os.path.exists(REF_SO):
            raise RuntimeError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(REF_SO)
        self.lib = L
        L.oracle_ref_table.restype = C.c_void_p
        L.oracle_ref_table.argtypes = [C.c_char_p, C.POINTER(C.c_size_t)]
        L.oracle_ref_scan.restype = C.c_void_p
        L.oracle_ref_scan.argtypes = [C.c_int]
        self.bpc = {}
        for hbd, sfx in ((False, "8bpc"), (True, "16bpc")):
            mc, itx, ip = pkg.MCDSPContext(), pkg.InvTxfmDSPContext(), pkg.IntraPredDSPContext()
            getattr(L, f"dav1d_mc_dsp_init_{sfx}")(C.byref(mc))
            getattr(L, f"dav1d_itx_dsp_init_{sfx}")(C.byref(itx), C.c_int(12 if hbd else 8))
            getattr(L, f"dav1d_intra_pred_dsp_init_{sfx}")(C.byref(ip))
            self.bpc[hbd] = DSPTables(mc, itx, ip, hbd)
        pe_common = [I, I, I, I, I, I, I, P, SZ, P, I, C.POINTER(I), I, I, I, P]
        self.prepare_intra_edges = {
            False: L.dav1d_prepare_intra_edges_8bpc, True: L.dav1d_prepare_intra_edges_16bpc}
        self.prepare_intra_edges[False].argtypes = pe_common
        self.prepare_intra_edges[True].argtypes = pe_common + [I]
        for f in self.prepare_intra_edges.values():
            f.restype = I

    def table(self, name, dtype):
        sz = C.c_size_t()
        p = self.lib.oracle_ref_table(name.encode(), C.byref(sz))
        assert p, name
        return np.frombuffer((C.c_char * sz.value).from_address(p), dtype=dtype).copy()

    def scan(self, tx, n):
        p = self.lib.oracle_ref_scan(tx)
        return np.frombuffer((C.c_uint16 * n).from_address(p), dtype=np.uint16).copy()
