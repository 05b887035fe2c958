"""ctypes declarations matching include/dav1d_cuda.h one to one."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdav1d_cuda.so")

N_2D_FILTERS = 10
N_RECT_TX_SIZES = 19
N_TX_TYPES_PLUS_LL = 17

# enum RectTxfmSize -> (w, h)   (reference src/levels.h:44-78)
TX_DIMS = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (4, 8), (8, 4), (8, 16), (16, 8),
           (16, 32), (32, 16), (32, 64), (64, 32), (4, 16), (16, 4), (8, 32), (32, 8),
           (16, 64), (64, 16)]


class MCDSPContext(C.Structure):
    _fields_ = [("mc", C.c_void_p * N_2D_FILTERS), ("mc_scaled", C.c_void_p * N_2D_FILTERS),
                ("mct", C.c_void_p * N_2D_FILTERS), ("mct_scaled", C.c_void_p * N_2D_FILTERS),
                ("avg", C.c_void_p), ("w_avg", C.c_void_p), ("mask", C.c_void_p),
                ("w_mask", C.c_void_p * 3), ("blend", C.c_void_p), ("blend_v", C.c_void_p),
                ("blend_h", C.c_void_p), ("warp8x8", C.c_void_p), ("warp8x8t", C.c_void_p),
                ("emu_edge", C.c_void_p), ("resize", C.c_void_p)]


class InvTxfmDSPContext(C.Structure):
    _fields_ = [("itxfm_add", (C.c_void_p * N_TX_TYPES_PLUS_LL) * N_RECT_TX_SIZES)]


class IntraPredDSPContext(C.Structure):
    _fields_ = [("intra_pred", C.c_void_p * 14), ("cfl_ac", C.c_void_p * 3),
                ("cfl_pred", C.c_void_p * 6), ("pal_pred", C.c_void_p)]


class ItxDesc(C.Structure):
    _fields_ = [("coef_off", C.c_uint32), ("x", C.c_uint16), ("y", C.c_uint16),
                ("eob", C.c_int16), ("plane", C.c_uint8), ("tx", C.c_uint8),
                ("txtp", C.c_uint8), ("cw4", C.c_uint8), ("ch4", C.c_uint8), ("pad", C.c_uint8)]


class McSrc(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("ref", C.c_uint8),
                ("filter_2d", C.c_uint8), ("mx", C.c_uint8), ("my", C.c_uint8)]


class McDesc(C.Structure):
    _fields_ = [("x", C.c_uint16), ("y", C.c_uint16), ("w", C.c_uint8), ("h", C.c_uint8),
                ("plane", C.c_uint8), ("kind", C.c_uint8), ("src", McSrc * 2),
                ("weight", C.c_uint8), ("mask_ss", C.c_uint8), ("aux16", C.c_uint16),
                ("aux_off", C.c_uint32)]


class Dav1dPictureMirror(C.Structure):
    """== Dav1dPicture (include/dav1d/picture.h:53-105): what the allocator callbacks fill."""
    _fields_ = [("seq_hdr", C.c_void_p), ("frame_hdr", C.c_void_p), ("data", C.c_void_p * 3),
                ("stride", C.c_ssize_t * 2), ("w", C.c_int32), ("h", C.c_int32), ("layout", C.c_int32),
                ("bpc", C.c_int32), ("opaque", C.c_uint64 * 24), ("allocator_data", C.c_void_p)]


class PicAllocator(C.Structure):
    _fields_ = [("cookie", C.c_void_p),
                ("alloc_picture_callback", C.CFUNCTYPE(C.c_int, C.POINTER(Dav1dPictureMirror), C.c_void_p)),
                ("release_picture_callback", C.CFUNCTYPE(None, C.POINTER(Dav1dPictureMirror), C.c_void_p))]


class BlockIntra(C.Structure):
    _fields_ = [("bx4", C.c_uint16), ("by4", C.c_uint16), ("bw4", C.c_uint8), ("bh4", C.c_uint8),
                ("y_mode", C.c_uint8), ("uv_mode", C.c_uint8), ("y_angle", C.c_int8), ("uv_angle", C.c_int8),
                ("tx", C.c_uint8), ("uvtx", C.c_uint8), ("pal_sz", C.c_uint8 * 2), ("cfl_alpha", C.c_int8 * 2),
                ("skip", C.c_uint8), ("edge_flags", C.c_uint8), ("sm_flags", C.c_uint8), ("pad", C.c_uint8),
                ("pal_off", C.c_uint32 * 3), ("pal_idx_off", C.c_uint32 * 2)]


class TxCoef(C.Structure):
    _fields_ = [("coef_off", C.c_uint32), ("eob", C.c_int16), ("txtp", C.c_uint8), ("cw4", C.c_uint8),
                ("ch4", C.c_uint8), ("pad", C.c_uint8 * 3)]


class Recorder(C.Structure):
    _fields_ = [("bw4", C.c_int32), ("bh4", C.c_int32), ("layout", C.c_int32), ("intra_edge_filter", C.c_int32),
                ("tile_col_start", C.c_int32), ("tile_col_end", C.c_int32), ("tile_row_start", C.c_int32),
                ("tile_row_end", C.c_int32), ("intra", C.c_void_p), ("n_intra", C.c_int32), ("cap_intra", C.c_int32)]


class NbMv(C.Structure):
    _fields_ = [("mvx", C.c_int16), ("mvy", C.c_int16), ("ref", C.c_int8), ("bw4", C.c_uint8), ("bh4", C.c_uint8),
                ("filter2d", C.c_uint8)]


class BlockInter(C.Structure):
    _fields_ = [("bx4", C.c_uint16), ("by4", C.c_uint16), ("bw4", C.c_uint8), ("bh4", C.c_uint8),
                ("comp_type", C.c_uint8), ("motion_mode", C.c_uint8), ("mvx", C.c_int16 * 2), ("mvy", C.c_int16 * 2),
                ("ref", C.c_int8 * 2), ("filter2d", C.c_uint8), ("mask_sign", C.c_uint8), ("skip", C.c_uint8),
                ("max_ytx", C.c_uint8), ("uvtx", C.c_uint8), ("interintra_type", C.c_uint8), ("tx_split", C.c_uint16 * 2),
                ("interintra_mode", C.c_uint8), ("warp", C.c_uint8), ("pad", C.c_uint8 * 2), ("wedge_mask", C.c_void_p * 3),
                ("ii_mask_off", C.c_uint32 * 3), ("pad2", C.c_uint32),
                ("warp_matrix", C.c_int32 * 6), ("warp_abcd", C.c_int16 * 4)]


class InterRecorder(C.Structure):
    _fields_ = [("bw4", C.c_int32), ("bh4", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("layout", C.c_int32),
                ("tile_col_start", C.c_int32), ("tile_row_start", C.c_int32),
                ("ref_w", C.c_int32 * 7), ("ref_h", C.c_int32 * 7), ("jnt_weights", (C.c_uint8 * 7) * 7),
                ("pad", C.c_uint8 * 7), ("above", C.c_void_p), ("left", C.c_void_p),
                ("put", C.c_void_p), ("n_put", C.c_int32), ("cap_put", C.c_int32),
                ("comp", C.c_void_p * 2), ("n_comp", C.c_int32 * 2), ("cap_comp", C.c_int32 * 2),
                ("obmc", C.c_void_p * 2), ("n_obmc", C.c_int32 * 2), ("cap_obmc", C.c_int32 * 2),
                ("scaled", C.c_void_p * 4), ("n_scaled", C.c_int32 * 4), ("cap_scaled", C.c_int32 * 4),
                ("itx", C.c_void_p), ("n_itx", C.c_int32), ("cap_itx", C.c_int32), ("masks_bytes", C.c_uint32),
                ("cap_masks", C.c_uint32), ("masks", C.c_void_p), ("intra", C.POINTER(Recorder)),
                ("warp", C.c_void_p), ("n_warp", C.c_int32), ("cap_warp", C.c_int32),
                ("intrabc", C.c_int32), ("pad3", C.c_int32), ("sub8", NbMv * 4), ("sub8_x", C.c_int32), ("sub8_y", C.c_int32)]


class Plane(C.Structure):
    _fields_ = [("data", C.c_void_p), ("stride", C.c_ssize_t), ("w", C.c_int32), ("h", C.c_int32)]


class Picture(C.Structure):
    _fields_ = [("p", Plane * 3), ("bitdepth_max", C.c_int32), ("ss_hor", C.c_int32),
                ("ss_ver", C.c_int32), ("tma", C.c_void_p)]


assert C.sizeof(ItxDesc) == 16 and C.sizeof(McDesc) == 40

_lib = None


def lib():
    """Load libdav1d_cuda.so. Raises if the CUDA library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C dav1d-mirror_b200` "
            "(there is no CPU fallback for the ported operators)")
    L = C.CDLL(LIB_PATH)
    L.dav1d_cuda_last_error.restype = C.c_int
    L.dav1d_cuda_last_error_string.restype = C.c_char_p
    L.dav1d_cuda_launch_count.restype = C.c_uint64
    L.dav1d_cuda_available.restype = C.c_int
    L.dav1d_cuda_open.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    L.dav1d_cuda_close.argtypes = [C.c_void_p]
    L.dav1d_cuda_synchronize.argtypes = [C.c_void_p]
    L.dav1d_cuda_picture_alloc.argtypes = [C.c_void_p, C.POINTER(Picture)] + [C.c_int] * 5
    L.dav1d_cuda_picture_free.argtypes = [C.c_void_p, C.POINTER(Picture)]
    L.dav1d_cuda_resize_frame.argtypes = [C.c_void_p, C.POINTER(Picture), C.POINTER(Picture), C.POINTER(C.c_int32),
                                          C.POINTER(C.c_int32)]
    L.dav1d_cuda_set_mc_tma.argtypes = [C.c_int]
    L.dav1d_cuda_set_mc_tma.restype = None
    L.dav1d_cuda_get_mc_tma.restype = C.c_int
    L.dav1d_cuda_picture_upload.argtypes = [C.c_void_p, C.POINTER(Picture), C.c_int, C.c_void_p, C.c_ssize_t]
    L.dav1d_cuda_picture_download.argtypes = [C.c_void_p, C.POINTER(Picture), C.c_int, C.c_void_p, C.c_ssize_t]
    L.dav1d_cuda_itx_batch.argtypes = [C.c_void_p, C.POINTER(Picture), C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_int32), C.c_int]
    bind_frame_api(L)
    _lib = L
    return L


def check_error():
    L = lib()
    e = L.dav1d_cuda_last_error()
    if e:
        msg = L.dav1d_cuda_last_error_string().decode()
        raise RuntimeError(f"dav1d_cuda error {e}: {msg}")


# ---------------------------------------------------------------- frame-level batch
class WarpDesc(C.Structure):
    _fields_ = [("x", C.c_uint16), ("y", C.c_uint16), ("sx", C.c_int32), ("sy", C.c_int32),
                ("mx", C.c_int32), ("my", C.c_int32), ("abcd", C.c_int16 * 4),
                ("plane", C.c_uint8), ("ref", C.c_uint8), ("pad", C.c_uint16)]


class IntraDesc40(C.Structure):
    _fields_ = [("x4", C.c_uint16), ("y4", C.c_uint16), ("tile_x4_start", C.c_uint16),
                ("tile_y4_start", C.c_uint16), ("tile_x4_end", C.c_uint16),
                ("tile_y4_end", C.c_uint16), ("plane", C.c_uint8), ("tw4", C.c_uint8),
                ("th4", C.c_uint8), ("mode", C.c_uint8), ("angle_delta", C.c_int8),
                ("edge_flags", C.c_uint8), ("flags", C.c_uint16), ("eob", C.c_int16),
                ("tx", C.c_uint8), ("txtp", C.c_uint8), ("coef_off", C.c_uint32),
                ("aux", C.c_uint32), ("reserved", C.c_uint32), ("cw4", C.c_uint8), ("ch4", C.c_uint8),
                ("pad", C.c_uint16)]


IntraDesc = IntraDesc40
assert C.sizeof(IntraDesc) == 40 and C.sizeof(WarpDesc) == 32


class McScaledSrc(C.Structure):
    _fields_ = [("pos_x", C.c_int32), ("pos_y", C.c_int32), ("step_x", C.c_int32), ("step_y", C.c_int32),
                ("ref", C.c_uint8), ("filter_2d", C.c_uint8), ("pad", C.c_uint16)]


class McScaledDesc(C.Structure):
    _fields_ = [("x", C.c_uint16), ("y", C.c_uint16), ("w", C.c_uint8), ("h", C.c_uint8), ("plane", C.c_uint8),
                ("kind", C.c_uint8), ("src", McScaledSrc * 2), ("weight", C.c_uint8), ("mask_ss", C.c_uint8),
                ("aux16", C.c_uint16), ("aux_off", C.c_uint32)]


class LfFrame(C.Structure):
    _fields_ = [("w4", C.c_int32), ("h4", C.c_int32), ("b4_stride", C.c_int32), ("sb128w", C.c_int32),
                ("filter_uv", C.c_int32), ("masks", C.c_void_p), ("level", C.c_void_p),
                ("lut_e", C.c_uint8 * 64), ("lut_i", C.c_uint8 * 64)]


class CdefFrame(C.Structure):
    _fields_ = [("bw", C.c_int32), ("bh", C.c_int32), ("sb128w", C.c_int32), ("damping", C.c_int32),
                ("y_strength", C.c_uint8 * 8), ("uv_strength", C.c_uint8 * 8), ("masks", C.c_void_p)]


class LrFrame(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("sb128w", C.c_int32), ("sb128", C.c_int32),
                ("unit_size_log2", C.c_int32 * 2), ("restore_planes", C.c_int32), ("lr_mask", C.c_void_p)]


class ReconBatch(C.Structure):
    _fields_ = [("dst", C.POINTER(Picture)), ("refs", C.POINTER(Picture) * 7),
                ("bw4", C.c_int32), ("bh4", C.c_int32),
                ("cf", C.c_void_p), ("masks", C.c_void_p), ("pal", C.c_void_p), ("pal_idx", C.c_void_p),
                ("mc_put", C.c_void_p), ("mc_put_tiles", C.c_void_p), ("n_mc_put_tiles", C.c_int32),
                ("mc_comp", C.c_void_p), ("mc_comp_tiles", C.c_void_p), ("n_mc_comp_tiles", C.c_int32 * 2),
                ("n_mc_put_small", C.c_int32), ("n_mc_comp_small", C.c_int32 * 2),
                ("mc_obmc", C.c_void_p), ("mc_obmc_tiles", C.c_void_p), ("n_mc_obmc_tiles", C.c_int32 * 2),
                ("warp", C.c_void_p), ("n_warp", C.c_int32),
                ("itx", C.c_void_p), ("itx_class_count", C.c_int32 * N_RECT_TX_SIZES),
                ("itx_tasks", C.c_void_p), ("n_itx_tasks", C.c_int32 * 2),
                ("intra", C.c_void_p), ("n_intra", C.c_int32),
                ("intra_cellmap", C.c_void_p),
                ("intra_itx", C.c_void_p), ("intra_itx_class_count", C.c_int32 * N_RECT_TX_SIZES),
                ("intra_itx_tasks", C.c_void_p), ("n_intra_itx_tasks", C.c_int32 * 2),
                ("intra_res", C.POINTER(Picture)), ("intra_levels_recorded", C.c_int32),
                ("mc_scaled", C.c_void_p), ("n_mc_scaled", C.c_int32 * 4),
                ("cf_int16", C.c_int32), ("cf_esc", C.c_void_p), ("n_cf_esc", C.c_int32)]


MAX_GROUP = 64


def bind_frame_api(L):
    L.dav1d_cuda_mc_put_batch.argtypes = [C.c_void_p, C.POINTER(Picture), C.POINTER(C.POINTER(Picture)),
                                          C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.dav1d_cuda_mc_compound_batch.argtypes = [C.c_void_p, C.POINTER(Picture), C.POINTER(C.POINTER(Picture)),
                                               C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.dav1d_cuda_warp_batch.argtypes = [C.c_void_p, C.POINTER(Picture), C.POINTER(C.POINTER(Picture)),
                                        C.c_void_p, C.c_int]
    L.dav1d_cuda_pic_allocator_init.argtypes = [C.c_void_p, C.POINTER(PicAllocator)]
    L.dav1d_cuda_picture_of.argtypes = [C.POINTER(Dav1dPictureMirror)]
    L.dav1d_cuda_picture_of.restype = C.POINTER(Picture)
    L.dav1d_cuda_picture_to_host.argtypes = [C.c_void_p, C.POINTER(Dav1dPictureMirror)]
    L.dav1d_cuda_picture_to_device.argtypes = [C.c_void_p, C.POINTER(Dav1dPictureMirror)]
    L.dav1d_cuda_record_b_intra.argtypes = [C.POINTER(Recorder), C.POINTER(BlockIntra), C.c_void_p, C.c_int]
    L.dav1d_cuda_loopfilter_frame.argtypes = [C.c_void_p, C.POINTER(Picture), C.POINTER(LfFrame)]
    L.dav1d_cuda_cdef_frame.argtypes = [C.c_void_p, C.POINTER(Picture), C.POINTER(Picture), C.POINTER(CdefFrame)]
    L.dav1d_cuda_lr_frame.argtypes = [C.c_void_p, C.POINTER(Picture), C.POINTER(Picture), C.POINTER(Picture),
                                      C.POINTER(LrFrame)]
    L.dav1d_cuda_pack_coefs.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
    L.dav1d_cuda_record_b_inter.argtypes = [C.POINTER(InterRecorder), C.POINTER(BlockInter), C.c_void_p, C.c_int]
    L.dav1d_cuda_record_nb_intra.argtypes = [C.POINTER(InterRecorder)] + [C.c_int] * 4
    L.dav1d_cuda_intra_cellmap_bytes.restype = C.c_size_t
    L.dav1d_cuda_intra_cellmap_bytes.argtypes = [C.c_int] * 4
    L.dav1d_cuda_intra_levels.argtypes = [C.c_void_p] + [C.c_int] * 5
    L.dav1d_cuda_intra_levels.restype = C.c_int
    L.dav1d_cuda_itx_tasks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32)]
    L.dav1d_cuda_itx_task_batch.argtypes = [C.c_void_p, C.POINTER(Picture), C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.c_int]
    L.dav1d_cuda_recon_submit.argtypes = [C.c_void_p, C.POINTER(ReconBatch)]
    L.dav1d_cuda_recon_submit_phases.argtypes = [C.c_void_p, C.POINTER(ReconBatch), C.c_int]
    L.dav1d_cuda_recon_group_submit.argtypes = [C.c_void_p, C.POINTER(C.POINTER(ReconBatch)), C.c_int]
    L.dav1d_cuda_recon_group_submit_phases.argtypes = [C.c_void_p, C.POINTER(C.POINTER(ReconBatch)), C.c_int, C.c_int]
    L.dav1d_cuda_recon_graph_build.argtypes = [C.c_void_p, C.POINTER(ReconBatch), C.POINTER(C.c_void_p)]
    L.dav1d_cuda_recon_graph_build_multi_phases.argtypes = [C.c_void_p, C.POINTER(C.POINTER(ReconBatch)), C.c_int,
                                                            C.c_int, C.POINTER(C.c_void_p)]
    L.dav1d_cuda_recon_graph_build_multi.argtypes = [C.c_void_p, C.POINTER(C.POINTER(ReconBatch)), C.c_int,
                                                     C.POINTER(C.c_void_p)]
    L.dav1d_cuda_recon_graph_launch.argtypes = [C.c_void_p, C.c_void_p]
    L.dav1d_cuda_recon_graph_free.argtypes = [C.c_void_p]
    L.dav1d_cuda_malloc.restype = C.c_void_p
    L.dav1d_cuda_malloc.argtypes = [C.c_size_t]
    L.dav1d_cuda_free.argtypes = [C.c_void_p]
    L.dav1d_cuda_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.dav1d_cuda_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.dav1d_cuda_memset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]
    L.dav1d_cuda_host_alloc.restype = C.c_void_p
    L.dav1d_cuda_host_alloc.argtypes = [C.c_size_t]
    L.dav1d_cuda_host_free.argtypes = [C.c_void_p]
    L.dav1d_cuda_event_create.restype = C.c_void_p
    L.dav1d_cuda_event_record.argtypes = [C.c_void_p, C.c_void_p]
    L.dav1d_cuda_stream_wait_event.argtypes = [C.c_void_p, C.c_void_p]
    L.dav1d_cuda_event_elapsed_ms.restype = C.c_float
    L.dav1d_cuda_event_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p]
    L.dav1d_cuda_event_destroy.argtypes = [C.c_void_p]
