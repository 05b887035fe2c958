"""Host-side mirror (ctypes) of the C ABI in include/dav1d_cuda.h.

The product is libdav1d_cuda.so (hand-written sm_100a kernels + C ABI); this
package only binds it for the test-suite and bench.py.  There is no Python or
CPU fallback: if the shared library is missing, importing `lib()` raises.
"""
from .binding import (  # noqa: F401
    lib, LIB_PATH, MCDSPContext, InvTxfmDSPContext, IntraPredDSPContext,
    ItxDesc, McSrc, McDesc, IntraDesc, WarpDesc, ReconBatch, Plane, Picture, TX_DIMS, check_error,
)
