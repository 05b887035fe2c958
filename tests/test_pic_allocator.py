"""The Dav1dPicAllocator seam (include/dav1d/picture.h:107-146): dav1d_cuda_pic_allocator_init() hands dav1d
callbacks whose pictures live in HBM and in pinned host memory with the geometry of the reference's default
allocator (src/picture.c:46-84)."""
import ctypes as C

import numpy as np
import pytest

import _d1pkg

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import binding as B  # noqa: E402
from dav1d_mirror_b200 import frame as F  # noqa: E402


def default_alloc_geometry(w, h, layout, bpc):
    """src/picture.c:46-66"""
    hbd = bpc > 8
    aw, ah = (w + 127) & ~127, (h + 127) & ~127
    ss_ver, ss_hor = layout == 1, layout != 3
    ys = aw << hbd
    uvs = (ys >> ss_hor) if layout else 0
    if not ys & 1023:
        ys += 64
    if layout and not uvs & 1023:
        uvs += 64
    return ys, uvs, ys * ah, uvs * (ah >> ss_ver)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,layout,bpc", [(256, 192, 1, 10), (328, 200, 3, 8), (1920, 1080, 2, 12), (64, 64, 0, 8)])
def test_allocator_callbacks(w, h, layout, bpc):
    L = pkg.lib()
    ctx = F.open_context(0)
    a = B.PicAllocator()
    assert L.dav1d_cuda_pic_allocator_init(ctx, C.byref(a)) == 0
    pic = B.Dav1dPictureMirror()
    pic.w, pic.h, pic.layout, pic.bpc = w, h, layout, bpc
    assert a.alloc_picture_callback(C.byref(pic), a.cookie) == 0
    try:
        ys, uvs, ysz, uvsz = default_alloc_geometry(w, h, layout, bpc)
        assert (pic.stride[0], pic.stride[1]) == (ys, uvs)
        assert pic.data[0] % 64 == 0                                        # DAV1D_PICTURE_ALIGNMENT
        if layout:
            assert pic.data[1] == pic.data[0] + ysz and pic.data[2] == pic.data[1] + uvsz
        else:
            assert not pic.data[1] and not pic.data[2]
        dev = L.dav1d_cuda_picture_of(C.byref(pic)).contents
        assert dev.p[0].stride == ys and dev.p[0].w == w and dev.p[0].h == h and dev.bitdepth_max == (1 << bpc) - 1
        # host planes -> HBM -> read back through the plane download: the twins share one layout
        bpp = 2 if bpc > 8 else 1
        rng = np.random.default_rng(w + h)
        host_y = np.frombuffer((C.c_char * (ys * h)).from_address(pic.data[0]), dtype=np.uint8)
        host_y[:] = rng.integers(0, 256, size=host_y.size, dtype=np.uint8)
        assert L.dav1d_cuda_picture_to_device(ctx, C.byref(pic)) == 0
        out = np.zeros((h, w * bpp), dtype=np.uint8)
        assert L.dav1d_cuda_picture_download(ctx, C.byref(dev), 0, out.ctypes.data, out.strides[0]) == 0
        L.dav1d_cuda_synchronize(ctx)
        assert np.array_equal(out, host_y.reshape(h, ys)[:, :w * bpp])
        # HBM -> host: change the device copy, bring it back
        L.dav1d_cuda_memset(ctx, dev.p[0].data, 0x5a, ys * h)
        assert L.dav1d_cuda_picture_to_host(ctx, C.byref(pic)) == 0
        L.dav1d_cuda_synchronize(ctx)
        assert (host_y == 0x5a).all()
        # any other picture is not ours
        other = B.Dav1dPictureMirror()
        assert not L.dav1d_cuda_picture_of(C.byref(other))
    finally:
        a.release_picture_callback(C.byref(pic), a.cookie)
        assert not pic.allocator_data
        L.dav1d_cuda_close(ctx)
    pkg.check_error()
