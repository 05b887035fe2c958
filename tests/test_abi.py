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


def test_picture_mirror_matches_the_reference_header(tmp_path):
    """Dav1dCudaDav1dPicture / Dav1dCudaPicAllocator are layout-identical to the reference's Dav1dPicture /
    Dav1dPicAllocator (include/dav1d/picture.h:53-146): compiled against both headers where the reference
    tree exists (this container); the GPU box checks the sizes recorded here."""
    import subprocess
    from dav1d_mirror_b200 import binding as B
    assert C.sizeof(B.Dav1dPictureMirror) == 272 and B.Dav1dPictureMirror.allocator_data.offset == 264
    assert B.Dav1dPictureMirror.data.offset == 16 and B.Dav1dPictureMirror.stride.offset == 40
    assert B.Dav1dPictureMirror.w.offset == 56 and C.sizeof(B.PicAllocator) == 24
    ref = "/root/reference/include"
    if not os.path.isdir(ref):
        return
    src = tmp_path / "pic.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "dav1d/picture.h"
#include "dav1d_cuda.h"
#define SAME(a, b) if ((a) != (b)) { printf("%s != %s\n", #a, #b); bad = 1; }
int main(void) {
    int bad = 0;
    SAME(sizeof(Dav1dPicture), sizeof(Dav1dCudaDav1dPicture))
    SAME(offsetof(Dav1dPicture, data), offsetof(Dav1dCudaDav1dPicture, data))
    SAME(offsetof(Dav1dPicture, stride), offsetof(Dav1dCudaDav1dPicture, stride))
    SAME(offsetof(Dav1dPicture, p.w), offsetof(Dav1dCudaDav1dPicture, p.w))
    SAME(offsetof(Dav1dPicture, p.layout), offsetof(Dav1dCudaDav1dPicture, p.layout))
    SAME(offsetof(Dav1dPicture, p.bpc), offsetof(Dav1dCudaDav1dPicture, p.bpc))
    SAME(offsetof(Dav1dPicture, allocator_data), offsetof(Dav1dCudaDav1dPicture, allocator_data))
    SAME(sizeof(Dav1dPicAllocator), sizeof(Dav1dCudaPicAllocator))
    SAME(offsetof(Dav1dPicAllocator, alloc_picture_callback), offsetof(Dav1dCudaPicAllocator, alloc_picture_callback))
    SAME(offsetof(Dav1dPicAllocator, release_picture_callback), offsetof(Dav1dCudaPicAllocator, release_picture_callback))
    SAME(DAV1D_PIXEL_LAYOUT_I400, 0) SAME(DAV1D_PIXEL_LAYOUT_I420, 1) SAME(DAV1D_PIXEL_LAYOUT_I422, 2) SAME(DAV1D_PIXEL_LAYOUT_I444, 3)
    puts(bad ? "MISMATCH" : "OK");
    return bad;
}
''')
    exe = tmp_path / "pic"
    subprocess.check_call(["gcc", "-I", ref, "-I", os.path.join(ROOT, "oracle", "ref_cfg"),
                           "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    assert subprocess.check_output([str(exe)]).decode().strip() == "OK"
