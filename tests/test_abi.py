"""CPU-only: the C-ABI library loads and exports every symbol include/dav1d_cuda.h
declares; struct layouts of the ctypes mirror match the header's documented sizes."""
import ctypes as C
import os
import re

import _d1pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = _d1pkg.load_pkg()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dav1d_cuda.h")).read()
    return sorted(set(re.findall(r"DAV1D_CUDA_API[^;{]*?\b(dav1d_cuda_\w+)\s*\(", src, flags=re.S)))


def test_library_exports_every_declared_symbol():
    L = pkg.lib()
    syms = declared_symbols()
    assert len(syms) >= 35, syms
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_struct_layouts():
    from dav1d_mirror_b200 import binding as B
    assert C.sizeof(B.ItxDesc) == 16
    assert C.sizeof(B.McDesc) == 40
    assert C.sizeof(B.IntraDesc) == 40
    assert C.sizeof(B.WarpDesc) == 32
    assert C.sizeof(B.MCDSPContext) == 53 * 8        # == sizeof(Dav1dMCDSPContext), SURVEY 8a1
    assert C.sizeof(B.InvTxfmDSPContext) == 19 * 17 * 8
    assert C.sizeof(B.IntraPredDSPContext) == 24 * 8


def test_no_device_means_loud_failure_not_fallback():
    """Without a CUDA device the init functions must leave the tables untouched and set the error."""
    L = pkg.lib()
    if L.dav1d_cuda_available():
        return
    from dav1d_mirror_b200 import binding as B
    mc = B.MCDSPContext()
    L.dav1d_cuda_mc_dsp_init_8bpc(C.byref(mc))
    assert all(p is None for p in mc.mc) and mc.avg is None
    assert L.dav1d_cuda_last_error() != 0
    L.dav1d_cuda_clear_error()


def test_ctypes_mirror_matches_the_header(tmp_path):
    """sizeof / offsetof of the batch structs as the C compiler lays them out == the ctypes mirror."""
    import subprocess
    from dav1d_mirror_b200 import binding as B
    src = tmp_path / "sz.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "dav1d_cuda.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu ", sizeof(Dav1dCudaItxDesc), sizeof(Dav1dCudaMcDesc), sizeof(Dav1dCudaIntraDesc),
           sizeof(Dav1dCudaWarpDesc), sizeof(Dav1dCudaPicture), sizeof(Dav1dCudaReconBatch));
    printf("%zu %zu %zu %zu %zu\n", offsetof(Dav1dCudaReconBatch, mc_obmc), offsetof(Dav1dCudaReconBatch, itx_tasks),
           offsetof(Dav1dCudaReconBatch, intra_itx), offsetof(Dav1dCudaReconBatch, intra_cellmap),
           offsetof(Dav1dCudaIntraDesc, cw4));
    return 0;
}
''')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(B.ItxDesc), C.sizeof(B.McDesc), C.sizeof(B.IntraDesc), C.sizeof(B.WarpDesc),
            C.sizeof(B.Picture), C.sizeof(B.ReconBatch), B.ReconBatch.mc_obmc.offset, B.ReconBatch.itx_tasks.offset,
            B.ReconBatch.intra_itx.offset, B.ReconBatch.intra_cellmap.offset, B.IntraDesc.cw4.offset]
    assert got == want, (got, want)
