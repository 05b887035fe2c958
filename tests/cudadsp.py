"""DSP tables as filled by the product (libdav1d_cuda.so), wrapped with the
same typed callables as the reference tables so the parity tests read
`ref.fn(args)` vs `cuda.fn(args)` like checkasm's call_ref / call_new."""
import ctypes as C

import _d1pkg
from refdsp import DSPTables

pkg = _d1pkg.load_pkg()


class CudaDSP:
    def __init__(self):
        L = pkg.lib()
        self.lib = L
        if not L.dav1d_cuda_available():
            raise RuntimeError("no CUDA device: the CUDA DSP path has no CPU fallback")
        self.bpc = {}
        for hbd, sfx in ((False, "8bpc"), (True, "16bpc")):
            mc, itx, ip = pkg.MCDSPContext(), pkg.InvTxfmDSPContext(), pkg.IntraPredDSPContext()
            for name, obj, extra in (("mc", mc, ()), ("itx", itx, (C.c_int(12 if hbd else 8),)),
                                     ("intra_pred", ip, ())):
                fn = getattr(L, f"dav1d_cuda_{name}_dsp_init_{sfx}", None)
                if fn is not None:
                    fn(C.byref(obj), *extra)
            self.bpc[hbd] = DSPTables(mc, itx, ip, hbd)
        pkg.check_error()
