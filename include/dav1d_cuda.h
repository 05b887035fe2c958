/*
 * dav1d_cuda.h - C ABI of the B200 (sm_100a) backend for dav1d's
 * pixel-reconstruction DSP: motion compensation, inverse transforms and
 * intra prediction, 8/10/12 bit.
 *
 * Two surfaces, both plain C (pointers + sizes, no C++/torch types):
 *
 *  (1) DSP-table overrides.  dav1d_cuda_{mc,itx,intra_pred}_dsp_init_{8,16}bpc
 *      fill the reference's own struct-of-function-pointers
 *      (Dav1dMCDSPContext  src/mc.h:116-132,
 *       Dav1dInvTxfmDSPContext src/itx.h:42-44,
 *       Dav1dIntraPredDSPContext src/ipred.h:81-90) exactly like a per-arch
 *      `*_dsp_init_x86(c)` does (src/mc_tmpl.c:948-956, src/itx_tmpl.c:270-283,
 *      src/ipred_tmpl.c:767-773, src/x86/mc.h:113-118).  Every entry keeps the
 *      reference signature (decl_*_fn in src/mc.h:38-114, src/itx.h:37-40,
 *      src/ipred.h:44-79) and the reference semantics: caller-owned HOST
 *      buffers, byte strides, synchronous, void return.  Each call stages its
 *      operands to the GPU, runs the same kernel the batched path uses with a
 *      single descriptor, and copies the result back.  This surface exists
 *      for checkasm-style bit-exact comparison, not for speed.
 *
 *  (2) Batched entry points.  The caller (recon_tmpl.c turned into a
 *      descriptor recorder: pass 2 of the frame-threading split,
 *      src/decode.c:741-830, src/recon_tmpl.c:1195-2036) hands over
 *      device-resident arrays of block descriptors; one kernel launch per
 *      operator class consumes them against pictures that live in HBM.
 *
 * Errors: the reference DSP functions return void and cannot fail
 * (src/recon_tmpl.c:1068,1131,1192).  Failures of this backend are therefore
 * reported out of band through dav1d_cuda_last_error(); batched entry points
 * additionally return 0 / a negative errno-style value like the reference's
 * public API (include/dav1d/common.h DAV1D_ERR).  There is no CPU fallback
 * for any ported operator: without a usable device the init functions leave
 * the table untouched and set the sticky error.
 */
#ifndef DAV1D_CUDA_H
#define DAV1D_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DAV1D_CUDA_API __attribute__((visibility("default")))
#else
#define DAV1D_CUDA_API
#endif

/* ------------------------------------------------------------------ enums
 * Values are the reference's (src/levels.h:44-133,184-196); restated so the
 * header stands alone. */
enum { DAV1D_CUDA_N_2D_FILTERS = 10 };        /* enum Filter2d, levels.h:184-196 */
enum { DAV1D_CUDA_N_RECT_TX_SIZES = 19 };     /* enum RectTxfmSize, levels.h:44-78 */
enum { DAV1D_CUDA_N_TX_TYPES_PLUS_LL = 17 };  /* enum TxfmType, levels.h:80-100 */
enum { DAV1D_CUDA_N_IMPL_INTRA_PRED_MODES = 14 }; /* levels.h:108-133 */

enum Dav1dCudaIntraMode {                     /* DSP-table index, levels.h:108-133 */
    DAV1D_CUDA_DC_PRED = 0, DAV1D_CUDA_VERT_PRED, DAV1D_CUDA_HOR_PRED,
    DAV1D_CUDA_LEFT_DC_PRED, DAV1D_CUDA_TOP_DC_PRED, DAV1D_CUDA_DC_128_PRED,
    DAV1D_CUDA_Z1_PRED, DAV1D_CUDA_Z2_PRED, DAV1D_CUDA_Z3_PRED,
    DAV1D_CUDA_SMOOTH_PRED, DAV1D_CUDA_SMOOTH_V_PRED, DAV1D_CUDA_SMOOTH_H_PRED,
    DAV1D_CUDA_PAETH_PRED, DAV1D_CUDA_FILTER_PRED,
    DAV1D_CUDA_CFL_PRED = 13                  /* bitstream uv mode, same value */
};

/* ------------------------------------------------- (1) DSP-table overrides
 * Layout-compatible mirrors of the reference structs.  Slots are typed
 * void* here because the 8 bpc and 16 bpc variants differ by a trailing
 * `int bitdepth_max` argument (include/common/bitdepth.h:72); the exact
 * signatures are the reference typedefs named beside each slot. */
typedef struct Dav1dCudaMCDSPContext {        /* == Dav1dMCDSPContext */
    void *mc[DAV1D_CUDA_N_2D_FILTERS];        /* mc_fn          */
    void *mc_scaled[DAV1D_CUDA_N_2D_FILTERS]; /* mc_scaled_fn   */
    void *mct[DAV1D_CUDA_N_2D_FILTERS];       /* mct_fn         */
    void *mct_scaled[DAV1D_CUDA_N_2D_FILTERS];/* mct_scaled_fn  */
    void *avg;                                /* avg_fn         */
    void *w_avg;                              /* w_avg_fn       */
    void *mask;                               /* mask_fn        */
    void *w_mask[3];                          /* w_mask_fn: 444, 422, 420 */
    void *blend;                              /* blend_fn       */
    void *blend_v;                            /* blend_dir_fn   */
    void *blend_h;                            /* blend_dir_fn   */
    void *warp8x8;                            /* warp8x8_fn     */
    void *warp8x8t;                           /* warp8x8t_fn    */
    void *emu_edge;                           /* emu_edge_fn    */
    void *resize;                             /* resize_fn      */
} Dav1dCudaMCDSPContext;

typedef struct Dav1dCudaInvTxfmDSPContext {   /* == Dav1dInvTxfmDSPContext */
    void *itxfm_add[DAV1D_CUDA_N_RECT_TX_SIZES][DAV1D_CUDA_N_TX_TYPES_PLUS_LL]; /* itxfm_fn */
} Dav1dCudaInvTxfmDSPContext;

typedef struct Dav1dCudaIntraPredDSPContext { /* == Dav1dIntraPredDSPContext */
    void *intra_pred[DAV1D_CUDA_N_IMPL_INTRA_PRED_MODES]; /* angular_ipred_fn */
    void *cfl_ac[3];                          /* cfl_ac_fn: 420, 422, 444 */
    void *cfl_pred[6];                        /* cfl_pred_fn: DC, -, -, LEFT_DC, TOP_DC, DC_128 */
    void *pal_pred;                           /* pal_pred_fn */
} Dav1dCudaIntraPredDSPContext;

/* Replace dav1d_mc_dsp_init_{8,16}bpc's arch tail (src/mc_tmpl.c:948-956). */
DAV1D_CUDA_API void dav1d_cuda_mc_dsp_init_8bpc(Dav1dCudaMCDSPContext *c);
DAV1D_CUDA_API void dav1d_cuda_mc_dsp_init_16bpc(Dav1dCudaMCDSPContext *c);
/* Replace dav1d_itx_dsp_init_{8,16}bpc's arch tail (src/itx_tmpl.c:270-283). */
DAV1D_CUDA_API void dav1d_cuda_itx_dsp_init_8bpc(Dav1dCudaInvTxfmDSPContext *c, int bpc);
DAV1D_CUDA_API void dav1d_cuda_itx_dsp_init_16bpc(Dav1dCudaInvTxfmDSPContext *c, int bpc);
/* Replace dav1d_intra_pred_dsp_init_{8,16}bpc's arch tail (src/ipred_tmpl.c:767-773). */
DAV1D_CUDA_API void dav1d_cuda_intra_pred_dsp_init_8bpc(Dav1dCudaIntraPredDSPContext *c);
DAV1D_CUDA_API void dav1d_cuda_intra_pred_dsp_init_16bpc(Dav1dCudaIntraPredDSPContext *c);

/* On-device counterpart of dav1d_prepare_intra_edges_{8,16}bpc
 * (src/ipred_prepare_tmpl.c:77-204, declared src/ipred_prepare.h:77-84):
 * same arguments, same return value (the DSP-table mode index), HOST buffers.
 * `prefilter_toplevel_sb_edge` may be NULL. */
DAV1D_CUDA_API int dav1d_cuda_prepare_intra_edges_8bpc(
    int x, int have_left, int y, int have_top, int w, int h, int edge_flags,
    const uint8_t *dst, ptrdiff_t stride, const uint8_t *prefilter_toplevel_sb_edge,
    int mode, int *angle, int tw, int th, int filter_edge, uint8_t *topleft_out);
DAV1D_CUDA_API int dav1d_cuda_prepare_intra_edges_16bpc(
    int x, int have_left, int y, int have_top, int w, int h, int edge_flags,
    const uint16_t *dst, ptrdiff_t stride, const uint16_t *prefilter_toplevel_sb_edge,
    int mode, int *angle, int tw, int th, int filter_edge, uint16_t *topleft_out,
    int bitdepth_max);

/* Sticky out-of-band error (0 = none).  Cleared by dav1d_cuda_clear_error(). */
DAV1D_CUDA_API int dav1d_cuda_last_error(void);
DAV1D_CUDA_API const char *dav1d_cuda_last_error_string(void);
DAV1D_CUDA_API void dav1d_cuda_clear_error(void);
/* Number of kernels this library has launched so far (all surfaces). */
DAV1D_CUDA_API uint64_t dav1d_cuda_launch_count(void);
/* 1 if a CUDA device is usable by this process, else 0 (sets the error). */
DAV1D_CUDA_API int dav1d_cuda_available(void);

/* ------------------------------------------------- (2) batched entry points */

/* One plane / picture resident in HBM.  Same geometry rules as the
 * reference's default allocator (src/picture.c:46-84): planar Y,U,V, byte
 * strides, 64-byte aligned rows. `data` is a DEVICE pointer. */
typedef struct Dav1dCudaPlane {
    void     *data;
    ptrdiff_t stride;      /* bytes */
    int32_t   w, h;        /* visible size in pixels (edge clamp uses these) */
} Dav1dCudaPlane;

typedef struct Dav1dCudaPicture {
    Dav1dCudaPlane p[3];
    int32_t bitdepth_max;  /* 0xff, 0x3ff or 0xfff */
    int32_t ss_hor, ss_ver;/* chroma subsampling (4:2:0 = 1,1) */
    const void *tma;       /* set by dav1d_cuda_picture_alloc: the planes' TMA tensor maps (device memory,
                            * part of the picture's allocation) - a picture that carries them is staged with
                            * cp.async.bulk.tensor when it is a reference of the batched MC kernels.
                            * NULL (pictures described by hand over other memory): cp.async staging */
} Dav1dCudaPicture;

/* -- inverse transform + add (itxfm_add call sites: recon_tmpl.c:816,1347,1567,2017)
 * 16 bytes.  `coef_off` indexes the device cf stream in coef units (int16 for
 * 8 bpc, int32 for 10/12 bpc); the block's coefficients are stored exactly as
 * the reference stores them: column-major, min(w,32) x min(h,32)
 * (itx_tmpl.c:79-87).  Blocks with eob < 0 are not recorded. */
typedef struct Dav1dCudaItxDesc {
    uint32_t coef_off;
    uint16_t x, y;         /* top-left of the tx block in `plane`, pixels */
    int16_t  eob;
    uint8_t  plane;
    uint8_t  tx;           /* enum RectTxfmSize */
    uint8_t  txtp;         /* enum TxfmType (incl. WHT_WHT = 16) */
    uint8_t  cw4, ch4;     /* packed coefficients: only the first 4*cw4 columns x 4*ch4 rows (the
                            * bounding box of the non-zero coefficients, rounded up to 4) are stored,
                            * column-major with stride 4*ch4: coef(row y, col x) = cf[coef_off + y +
                            * x * 4*ch4]; everything outside is zero and is neither stored, shipped
                            * nor read.  0, 0 = dense min(w,32) x min(h,32) as the reference stores it */
    uint8_t  pad;
} Dav1dCudaItxDesc;

/* -- motion compensation (mc()/obmc() call sites: recon_tmpl.c:957-1069).
 * One descriptor = one prediction block of one plane.  Sources are
 * (ref picture, integer position, sub-pel phase, filter); positions may lie
 * outside the reference plane - the kernel clamps coordinates, which is what
 * emu_edge (mc_tmpl.c:827-875) followed by a normal read amounts to. */
enum Dav1dCudaMcKind {
    DAV1D_CUDA_MC_PUT    = 0,  /* dsp->mc.mc[filter]                       */
    DAV1D_CUDA_MC_AVG    = 1,  /* 2x mct + avg     (recon_tmpl.c:1843-1846) */
    DAV1D_CUDA_MC_W_AVG  = 2,  /* 2x mct + w_avg   (:1847-1851)             */
    DAV1D_CUDA_MC_MASK   = 3,  /* 2x mct + mask    (:1859-1868, wedge/chroma seg) */
    DAV1D_CUDA_MC_W_MASK = 4,  /* 2x mct + w_mask  (:1852-1858) writes seg mask */
    DAV1D_CUDA_MC_PREP   = 5,  /* dsp->mc.mct[filter] into the int16 tmp pool */
    /* overlapped block motion compensation, obmc() (recon_tmpl.c:1071-1131): the prediction of a
     * neighbour's motion vector over the block's edge region into a scratch tile + blend onto the
     * block's own prediction.  w, h = the size handed to mc() (it selects the 4-tap sets). */
    DAV1D_CUDA_MC_OBMC_H = 6,  /* top neighbour:  mc + blend_h(dst, lap, w, aux16)  (:1093-1103) */
    DAV1D_CUDA_MC_OBMC_V = 7   /* left neighbour: mc + blend_v(dst, lap, w, h)      (:1117-1127) */
};

typedef struct Dav1dCudaMcSrc {
    int32_t  x, y;         /* integer sample position of the block's top-left in the ref plane */
    uint8_t  ref;          /* index into Dav1dCudaReconBatch.refs[] */
    uint8_t  filter_2d;    /* enum Filter2d */
    uint8_t  mx, my;       /* 1/16 sample phase, 0..15 */
} Dav1dCudaMcSrc;

typedef struct Dav1dCudaMcDesc {   /* 40 bytes */
    uint16_t x, y;         /* destination position in `plane`, pixels */
    uint8_t  w, h;         /* 2..128 */
    uint8_t  plane;
    uint8_t  kind;         /* enum Dav1dCudaMcKind */
    Dav1dCudaMcSrc src[2]; /* src[1] unused for PUT/PREP */
    uint8_t  weight;       /* W_AVG: jnt weight 1..15; W_MASK: sign */
    uint8_t  mask_ss;      /* W_MASK: 0=444 1=422 2=420 layout of the emitted mask */
    uint16_t aux16;        /* OBMC_H: the height blend_h gets (v_mul * oh4 >= h) */
    uint32_t aux_off;      /* MASK: byte offset of the w*h mask in `masks`;
                              W_MASK: byte offset the emitted mask is written to in `masks`;
                              PREP: int16 offset into the tmp pool */
} Dav1dCudaMcDesc;

/* -- motion compensation from a reference of ANOTHER size (the scaled branch of mc(),
 * recon_tmpl.c:1010-1065; dsp->mc.mc_scaled / mct_scaled, src/mc.h:45-69).  The recorder runs the
 * scale_mv arithmetic (:1015-1021) and hands over the 1/1024-sample position of the block's
 * top-left in the reference plane: left = pos_x >> 10 (may be negative or beyond the plane - the
 * kernel clamps, which is what emu_edge :1036-1044 amounts to), mx = pos_x & 0x3ff, and the
 * steps f->svc[ref][0/1].step.  A compound block whose other reference has the frame's own size
 * describes that source with step 1024 and pos = (integer position << 10) + (phase << 6): the
 * scaled filters then compute exactly what mc / mct compute. */
typedef struct Dav1dCudaMcScaledSrc {   /* 20 bytes */
    int32_t  pos_x, pos_y;
    int32_t  step_x, step_y;
    uint8_t  ref;          /* index into Dav1dCudaReconBatch.refs[] */
    uint8_t  filter_2d;    /* enum Filter2d */
    uint16_t pad;
} Dav1dCudaMcScaledSrc;

typedef struct Dav1dCudaMcScaledDesc {  /* 56 bytes; fields as in Dav1dCudaMcDesc */
    uint16_t x, y;
    uint8_t  w, h;
    uint8_t  plane;
    uint8_t  kind;         /* PUT, AVG, W_AVG, MASK, W_MASK, OBMC_H, OBMC_V */
    Dav1dCudaMcScaledSrc src[2];
    uint8_t  weight;
    uint8_t  mask_ss;
    uint16_t aux16;
    uint32_t aux_off;
} Dav1dCudaMcScaledDesc;

/* -- intra-class operations, one descriptor per transform block in DECODE
 * order (recon_tmpl.c:1259-1300 luma, :1372-1417 CfL, :1226-1243 palette,
 * :1503-1576 chroma), optionally fused with the block's residual.  The fields
 * are the arguments of dav1d_prepare_intra_edges + intra_pred[m] /
 * cfl_ac + cfl_pred[m] / pal_pred; mode -> DSP index and absolute angle are
 * resolved on the device exactly like ipred_prepare_tmpl.c:94-117. */
enum Dav1dCudaIntraKind {
    /* 0..12: bitstream enum IntraPredMode (DC, VERT, HOR, D45.., SMOOTH*, PAETH) */
    DAV1D_CUDA_INTRA_FILTER = 13,   /* filter-intra, angle_delta = filter index 0..4 */
    DAV1D_CUDA_INTRA_CFL    = 14,   /* cfl_ac + DC edges + cfl_pred, angle_delta = alpha */
    DAV1D_CUDA_INTRA_PAL    = 15,   /* pal_pred: coef_off = byte offset of the packed indices,
                                       aux = palette offset (in pixels) in the palette pool */
    DAV1D_CUDA_INTRA_II     = 16,   /* inter-intra (recon_tmpl.c:1658-1681, 1779-1817): intra prediction of
                                       the whole block (angle_delta = IntraPredMode DC/VERT/HOR/SMOOTH,
                                       edge_flags 0, no edge filter) into scratch + mc.blend onto the inter
                                       prediction; coef_off = byte offset of the w*h blend mask in the
                                       `pal_idx` byte pool.  The block's residuals follow as
                                       DAV1D_CUDA_INTRA_NONE operations */
    DAV1D_CUDA_INTRA_IBC    = 17,   /* intrabc (recon_tmpl.c:1624-1637): bilinear mc() from the CURRENT picture,
                                       aux = source x | y << 16 (int16 each, pixels of this plane),
                                       angle_delta = mx, flags = my (0 or 8); reads are clamped to the
                                       4*bw4 x 4*bh4 area like emu_edge does (:974-995).  Scheduled after the
                                       operations that produced the source area; residuals follow as
                                       DAV1D_CUDA_INTRA_NONE operations */
    DAV1D_CUDA_INTRA_NONE   = 255   /* residual only (e.g. tx blocks of a palette or inter-intra block) */
};

typedef struct Dav1dCudaIntraDesc {  /* 40 bytes */
    uint16_t x4, y4;       /* position in 4-px units of THIS plane */
    uint16_t tile_x4_start;/* have_left = x4 > tile_x4_start */
    uint16_t tile_y4_start;/* have_top  = y4 > tile_y4_start */
    uint16_t tile_x4_end, tile_y4_end;   /* w, h arguments (tile end, frame end) */
    uint8_t  plane;
    uint8_t  tw4, th4;     /* transform (prediction) size in 4-px units */
    uint8_t  mode;         /* enum IntraPredMode 0..12 or enum Dav1dCudaIntraKind */
    int8_t   angle_delta;  /* -3..3 | filter index | cfl alpha */
    uint8_t  edge_flags;   /* bit0 top-has-right, bit3 left-has-bottom (EDGE_I444_*) */
    uint16_t flags;        /* bit9 smooth neighbour, bit10 edge filter (ipred_prepare.h:87-93) */
    int16_t  eob;          /* < 0: no residual */
    uint8_t  tx, txtp;     /* residual transform, if any */
    uint32_t coef_off;     /* into the cf stream (PAL: into the index pool) */
    uint32_t aux;          /* CFL: w_pad | h_pad << 8 (4-px units); PAL: palette offset */
    uint32_t reserved;     /* 0, or level + 1 in the low 16 bits: dav1d_cuda_intra_levels() */
    uint8_t  cw4, ch4;     /* packed residual coefficients, see Dav1dCudaItxDesc (0, 0 = dense) */
    uint16_t pad;
} Dav1dCudaIntraDesc;

/* -- warped motion, one descriptor per 8x8 (warp_affine(), recon_tmpl.c:1134-1193) */
typedef struct Dav1dCudaWarpDesc {   /* 32 bytes */
    uint16_t x, y;         /* destination of the 8x8 in `plane` */
    int32_t  sx, sy;       /* integer source position (dx, dy of recon_tmpl.c:1163-1164) */
    int32_t  mx, my;       /* fractional parts handed to warp8x8 (recon_tmpl.c:1165-1167) */
    int16_t  abcd[4];
    uint8_t  plane, ref;
    uint16_t pad;
} Dav1dCudaWarpDesc;

/* ---- the recorder: recon_tmpl.c's drivers as descriptor emitters (host, no device work).
 * dav1d_cuda_record_b_intra() is dav1d_recon_b_intra() (src/recon_tmpl.c:1195-1596) with every DSP call
 * replaced by one appended Dav1dCudaIntraDesc: same loops over the 64x64 units and transform blocks of
 * the block, same refinement of the block-level edge flags per transform block (:1270-1274,:1466-1476),
 * same order of palette / CfL / prediction / residual, the tile from `ts->tiling`.  The block is given
 * as the Av1Block fields that function reads (src/levels.h:262-287) plus what decode_b() hands over. */
typedef struct Dav1dCudaBlockIntra {
    uint16_t bx4, by4;            /* t->bx, t->by (4-px units) */
    uint8_t  bw4, bh4;            /* dav1d_block_dimensions[bs] */
    uint8_t  y_mode, uv_mode;     /* enum IntraPredMode; FILTER_PRED (luma) / CFL_PRED (chroma) = 13 */
    int8_t   y_angle, uv_angle;   /* b->y_angle (filter index for FILTER_PRED), b->uv_angle */
    uint8_t  tx, uvtx;            /* b->tx, b->uvtx: enum RectTxfmSize */
    uint8_t  pal_sz[2];           /* b->pal_sz */
    int8_t   cfl_alpha[2];        /* b->cfl_alpha */
    uint8_t  skip;                /* b->skip */
    uint8_t  edge_flags;          /* enum EdgeFlags as passed by decode_b() (src/intra_edge.h:27-32) */
    uint8_t  sm_flags;            /* bit 0: sm_flag(t->a, bx4) | sm_flag(&t->l, by4) != 0; bit 1: the same for
                                     sm_uv_flag (src/ipred_prepare.h:95-107) */
    uint8_t  pad;
    uint32_t pal_off[3];          /* the block's palettes in the palette pool (pixels): Y, U, V */
    uint32_t pal_idx_off[2];      /* its packed indices in the index pool (bytes): luma, chroma */
} Dav1dCudaBlockIntra;
/* One entry per transform block in the order dav1d_recon_b_intra consumes `cbi` / `cf` (luma raster,
 * then U raster, then V raster; none for a skip block): eob / txtp = the cbi entry, coef_off (+ cw4, ch4
 * for a packed box, else 0) = where the block's coefficients are in the cf stream. */
typedef struct Dav1dCudaTxCoef {
    uint32_t coef_off;
    int16_t  eob;
    uint8_t  txtp, cw4, ch4, pad[3];
} Dav1dCudaTxCoef;
typedef struct Dav1dCudaRecorder {
    int32_t bw4, bh4;             /* f->bw, f->bh */
    int32_t layout;               /* enum Dav1dPixelLayout: I400 0, I420 1, I422 2, I444 3 */
    int32_t intra_edge_filter;    /* f->seq_hdr->intra_edge_filter */
    int32_t tile_col_start, tile_col_end, tile_row_start, tile_row_end;   /* ts->tiling (luma 4-px units) */
    Dav1dCudaIntraDesc *intra;    /* output array (decode order) */
    int32_t n_intra, cap_intra;
} Dav1dCudaRecorder;
/* Appends the block's operations; returns how many, or a negative errno (-ENOSPC: `intra` is full,
 * -EINVAL: fields that no stream produces, n_tx != the block's transform blocks). */
DAV1D_CUDA_API int dav1d_cuda_record_b_intra(Dav1dCudaRecorder *r, const Dav1dCudaBlockIntra *b,
                                             const Dav1dCudaTxCoef *tx, int n_tx);

/* ---- inter half of the recorder: dav1d_recon_b_inter (recon_tmpl.c:1598-2036) with mc() (:957-1069, both
 * branches), obmc() (:1071-1132) and read_coef_tree() (:726-823) as descriptor emission.  Covers
 * translational single-reference blocks (optionally with OBMC) and the AVG / WEIGHTED_AVG / SEG compounds,
 * the WEDGE compound, inter-intra blocks (the intra prediction + blend and the block's residuals become
 * intra-class operations appended through `intra`), from references of any size, with their residual
 * transform trees, warped blocks (local warp and global motion: warp_affine(), :1134-1193, as one
 * Dav1dCudaWarpDesc per 8x8), the intrabc blocks of key / intra-only frames and the chroma of 4xN / Nx4 blocks
 * (:1685-1751: up to four predictions with the partners' vectors; an intrabc block that narrow predicts the
 * chroma of its 8x8 with its own vector, :1631-1635).  Every branch of both drivers is transcribed; -ENOSYS is
 * only returned for residual trees split deeper than one level inside inter-intra / intrabc blocks. */
typedef struct Dav1dCudaNbMv {          /* what obmc() reads of a neighbour (refmvs rows + filter contexts) */
    int16_t mvx, mvy;                   /* r->mv.mv[0] */
    int8_t  ref;                        /* r->ref.ref[0] - 1: reference index, -1 = intra */
    uint8_t bw4, bh4;                   /* dav1d_block_dimensions[r->bs] */
    uint8_t filter2d;                   /* dav1d_filter_2d[ctx->filter[1]][ctx->filter[0]] */
} Dav1dCudaNbMv;
typedef struct Dav1dCudaBlockInter {    /* the Av1Block fields dav1d_recon_b_inter reads (src/levels.h:262-287) */
    uint16_t bx4, by4;                  /* t->bx, t->by */
    uint8_t  bw4, bh4;                  /* dav1d_block_dimensions[bs] */
    uint8_t  comp_type;                 /* enum CompInterType: NONE 0, WEIGHTED_AVG 1, AVG 2, SEG 3, WEDGE 4 */
    uint8_t  motion_mode;               /* enum MotionMode: TRANSLATION 0, OBMC 1, WARP 2 */
    int16_t  mvx[2], mvy[2];            /* b->mv[i].x / .y */
    int8_t   ref[2];
    uint8_t  filter2d, mask_sign, skip;
    uint8_t  max_ytx, uvtx;             /* enum RectTxfmSize */
    uint8_t  interintra_type;           /* enum InterIntraType: NONE 0, BLEND 1, WEDGE 2 */
    uint16_t tx_split[2];               /* b->tx_split0, b->tx_split1 */
    uint8_t  interintra_mode;           /* enum InterIntraPredMode: II_DC 0, VERT 1, HOR 2, SMOOTH 3 */
    uint8_t  warp;                      /* 1: the block takes recon_b_inter's warp_affine branch (recon_tmpl.c:1641-1649:
                                           inter_mode == GLOBALMV && f->gmv_warp_allowed[ref], or motion_mode == MM_WARP &&
                                           t->warpmv.type > TRANSLATION) with the model in warp_matrix / warp_abcd */
    uint8_t  pad[2];
    /* COMP_INTER_WEDGE: the block's masks per plane as the driver picks them - luma WEDGE_MASK(0, bs, 0,
     * wedge_idx), chroma WEDGE_MASK(chr_layout_idx, bs, mask_sign, wedge_idx) (recon_tmpl.c:1861-1866; host
     * pointers, w * h bytes of the plane block each): copied into the recorder's mask pool. */
    const uint8_t *wedge_mask[3];
    /* inter-intra: where the caller put II_MASK(layout of the plane, bs, b) of each plane (wedge.h:88-93,
     * w * h bytes) in the byte pool that also holds the palette indices (Dav1dCudaReconBatch.pal_idx) */
    uint32_t ii_mask_off[3];
    uint32_t pad2;
    /* warp: the Dav1dWarpedMotionParams handed to warp_affine - t->warpmv or f->frame_hdr->gmv[ref]:
     * matrix[6] and the shear parameters u.abcd (alpha, beta, gamma, delta) */
    int32_t  warp_matrix[6];
    int16_t  warp_abcd[4];
} Dav1dCudaBlockInter;
typedef struct Dav1dCudaInterRecorder {
    int32_t bw4, bh4;                   /* f->bw, f->bh */
    int32_t w, h;                       /* f->cur.p.w, f->cur.p.h */
    int32_t layout;                     /* enum Dav1dPixelLayout */
    int32_t tile_col_start, tile_row_start;   /* ts->tiling (luma 4-px units) */
    int32_t ref_w[7], ref_h[7];         /* f->refp[i].p.p.w / .h (0 = the frame's size) */
    uint8_t jnt_weights[7][7];          /* f->jnt_weights */
    uint8_t pad[7];
    Dav1dCudaNbMv *above, *left;        /* bw4 + 1 / bh4 + 1 entries, zero-initialised by the caller once per
                                           frame; updated by the recorder after every block */
    /* outputs, each appended in decode order: the arrays the Dav1dCudaReconBatch takes (tiles from
     * dav1d_cuda_mc_tiles(), transform tasks from dav1d_cuda_itx_tasks() after sorting `itx` by size) */
    Dav1dCudaMcDesc *put;      int32_t n_put, cap_put;
    Dav1dCudaMcDesc *comp[2];  int32_t n_comp[2], cap_comp[2];     /* wave 0, wave 1 (chroma of SEG blocks) */
    Dav1dCudaMcDesc *obmc[2];  int32_t n_obmc[2], cap_obmc[2];     /* OBMC_H, OBMC_V */
    Dav1dCudaMcScaledDesc *scaled[4]; int32_t n_scaled[4], cap_scaled[4];
    Dav1dCudaItxDesc *itx;     int32_t n_itx, cap_itx;
    uint32_t masks_bytes;               /* running size of the mask pool (segmentation masks are allotted here) */
    uint32_t cap_masks;                 /* capacity of `masks` */
    uint8_t *masks;                     /* host mirror of the mask pool: wedge masks are copied into it (the
                                           segmentation masks' space is only reserved: the device writes them) */
    Dav1dCudaRecorder *intra;           /* the frame's intra recorder: inter-intra blocks append their intra-class
                                           operations to its array, in decode order with the intra blocks'
                                           (its tile_* fields must describe the current tile) */
    Dav1dCudaWarpDesc *warp;   int32_t n_warp, cap_warp;   /* warped blocks: one descriptor per 8x8 and plane */
    int32_t intrabc;                    /* IS_KEY_OR_INTRA(f->frame_hdr): every block handed to record_b_inter is an
                                           intrabc block (recon_tmpl.c:1624-1637) - its prediction (bilinear mc() from the
                                           picture being decoded) and residuals become intra-class operations through `intra` */
    int32_t pad3;
    /* the blocks of the 8x8 being decoded, per 4x4 (index (y & 1) * 2 + (x & 1)): what the refmvs rows and
     * f->frame_thread.b still hold of the left / top / top-left partner when the chroma of a 4xN / Nx4 block is
     * predicted (recon_tmpl.c:1685-1751) - kept by the recorder itself, zero-initialised by the caller */
    Dav1dCudaNbMv sub8[4];
    int32_t sub8_x, sub8_y;             /* origin of that 8x8 (4-px units) */
} Dav1dCudaInterRecorder;
/* Appends the block's descriptors; returns how many, or a negative errno (-ENOSPC: an array is full, -EINVAL,
 * -ENOSYS: see above).  On error nothing of the block is kept.  `tx`: the block's cbi / cf entries in
 * consumption order (luma tree, then U, then V per 64x64 unit), none when b->skip. */
DAV1D_CUDA_API int dav1d_cuda_record_b_inter(Dav1dCudaInterRecorder *r, const Dav1dCudaBlockInter *b,
                                             const Dav1dCudaTxCoef *tx, int n_tx);
/* An intra block in an inter frame: what obmc() of later blocks sees of it (decode.c:756-767). */
DAV1D_CUDA_API int dav1d_cuda_record_nb_intra(Dav1dCudaInterRecorder *r, int bx4, int by4, int bw4, int bh4);

typedef struct Dav1dCudaContext Dav1dCudaContext;

/* `stream` is a cudaStream_t the caller already orders its work on, or NULL:
 * the context then creates its own non-blocking stream (one context per
 * decoder instance / stream of frames; contexts run concurrently). */
DAV1D_CUDA_API int  dav1d_cuda_open(Dav1dCudaContext **out, int device, void *stream);
DAV1D_CUDA_API void dav1d_cuda_close(Dav1dCudaContext *c);
/* Waits for the context's stream; also reads the context's device status word: intra-class
 * operations that wait for each other (inconsistent descriptors; the frame is then incomplete) are
 * reported as -EIO (-5) through the return value and dav1d_cuda_last_error(). */
DAV1D_CUDA_API int  dav1d_cuda_synchronize(Dav1dCudaContext *c);

/* HBM picture allocation with the reference's geometry (src/picture.c:46-84:
 * width/height padded to 128, stride = aligned width << hbd, +64 B when the
 * stride is a multiple of 1024). */
DAV1D_CUDA_API int  dav1d_cuda_picture_alloc(Dav1dCudaContext *c, Dav1dCudaPicture *pic,
                                             int w, int h, int ss_hor, int ss_ver, int bitdepth_max);
DAV1D_CUDA_API void dav1d_cuda_picture_free(Dav1dCudaContext *c, Dav1dCudaPicture *pic);
/* Staging of the reference windows in the batched 32x32-tile MC kernels (process-wide switch, for
 * measurements and tests).  With TMA a window is ONE cp.async.bulk.tensor.2d request against the reference's
 * tensor maps, completing on an mbarrier, into one of two window buffers per warp: the next window is in
 * flight while the current one is filtered.  mode 0: per-lane cp.async copies everywhere; 1 (default): TMA
 * for single-reference predictions (measured 5 % faster than cp.async); 2: TMA for compound predictions too
 * (measured 3 % slower: DESIGN.md section 8).  Same pixels in every mode. */
DAV1D_CUDA_API void dav1d_cuda_set_mc_tma(int mode);
DAV1D_CUDA_API int  dav1d_cuda_get_mc_tma(void);
DAV1D_CUDA_API int  dav1d_cuda_picture_upload(Dav1dCudaContext *c, const Dav1dCudaPicture *pic, int plane,
                                              const void *host, ptrdiff_t host_stride);
DAV1D_CUDA_API int  dav1d_cuda_picture_download(Dav1dCudaContext *c, const Dav1dCudaPicture *pic, int plane,
                                                void *host, ptrdiff_t host_stride);

/* ---- Dav1dPicAllocator seam (include/dav1d/picture.h:107-146, src/picture.c:46-88).
 * dav1d asks its allocator for every frame (dav1d_thread_picture_alloc -> picture_alloc_with_edges ->
 * allocator.alloc_picture_callback).  dav1d_cuda_pic_allocator_init() fills a Dav1dPicAllocator whose
 * pictures live TWICE with the geometry of the default allocator: in HBM (the Dav1dCudaPicture the
 * batches reconstruct into and predict from) and in pinned host memory (what `data[]` points to, so
 * that dav1d_get_picture() consumers and the CPU post-filters keep working).  The two are moved with
 * ONE copy (same strides, one allocation each):
 *   dav1d_cuda_picture_to_host()    after reconstruction, before the CPU reads the frame;
 *   dav1d_cuda_picture_to_device()  after the CPU changed it (post-filters), before it serves as a reference.
 * Dav1dSettings.allocator = *(Dav1dPicAllocator *) &a;  the structs below are layout-identical mirrors. */
typedef struct Dav1dCudaDav1dPicture {        /* == Dav1dPicture, include/dav1d/picture.h:53-105 (272 bytes) */
    void *seq_hdr, *frame_hdr;
    void *data[3];
    ptrdiff_t stride[2];
    struct { int32_t w, h, layout /* enum Dav1dPixelLayout: I400 0, I420 1, I422 2, I444 3 */, bpc; } p;
    uint64_t opaque[24];                      /* m, metadata, reference bookkeeping: not touched */
    void *allocator_data;
} Dav1dCudaDav1dPicture;
typedef struct Dav1dCudaPicAllocator {        /* == Dav1dPicAllocator */
    void *cookie;
    int (*alloc_picture_callback)(Dav1dCudaDav1dPicture *pic, void *cookie);
    void (*release_picture_callback)(Dav1dCudaDav1dPicture *pic, void *cookie);
} Dav1dCudaPicAllocator;
DAV1D_CUDA_API int dav1d_cuda_pic_allocator_init(Dav1dCudaContext *c, Dav1dCudaPicAllocator *a);
/* The HBM twin of a picture this allocator handed out (NULL for any other picture). */
DAV1D_CUDA_API const Dav1dCudaPicture *dav1d_cuda_picture_of(const Dav1dCudaDav1dPicture *pic);
DAV1D_CUDA_API int dav1d_cuda_picture_to_host(Dav1dCudaContext *c, const Dav1dCudaDav1dPicture *pic);
DAV1D_CUDA_API int dav1d_cuda_picture_to_device(Dav1dCudaContext *c, const Dav1dCudaDav1dPicture *pic);

/* ---- In-loop post-filters, first stage: the deblocking loop filter of a whole frame, in place in HBM.
 * Replaces dav1d_loopfilter_sbrow_cols / _rows over all superblock rows (src/lf_apply_tmpl.c:306-466,
 * called from dav1d_filter_sbrow_deblock_cols / _rows, recon_tmpl.c:2038-2071) and the loop_filter_sb
 * functions they call (src/loopfilter_tmpl.c).  The inputs are dav1d's own structures, copied to the device
 * as they are when the filter starts:
 *   masks = f->lf.mask: Av1Filter[f->sb128w * f->sb128h] (src/lf_mask.h:52-58, 1348 bytes each), after the
 *           tile-edge fix-ups of lf_apply_tmpl.c:323-387 where the frame has several tiles;
 *   level = f->lf.level: uint8_t[4] per 4x4 block, row stride f->b4_stride (lf_mask.c:307-312);
 *   lut_e / lut_i = f->lf.lim_lut.e / .i (dav1d_calc_eih).
 * Two passes (all column edges, then all row edges) of independent edges; asynchronous on the context's
 * stream.  Super-resolution and loop restoration are not built: after CDEF (below) the frame goes to the host
 * for them (dav1d_cuda_picture_to_host). */
typedef struct Dav1dCudaLfFrame {
    int32_t w4, h4;                     /* f->w4, f->h4 */
    int32_t b4_stride, sb128w;          /* f->b4_stride, f->sb128w */
    int32_t filter_uv;                  /* frame_hdr->loopfilter.level_u || level_v */
    const void *masks;                  /* device */
    const uint8_t *level;               /* device */
    uint8_t lut_e[64], lut_i[64];
} Dav1dCudaLfFrame;
DAV1D_CUDA_API int dav1d_cuda_loopfilter_frame(Dav1dCudaContext *c, const Dav1dCudaPicture *pic,
                                               const Dav1dCudaLfFrame *lf);

/* Second stage: CDEF of a whole frame, OUT OF PLACE (dst != src; src = the deblocked picture).  Replaces
 * dav1d_filter_sbrow_cdef over all superblock rows (recon_tmpl.c:2073-2100 -> dav1d_cdef_brow,
 * src/cdef_apply_tmpl.c:98-309 -> dsp->cdef.dir / .fb[], src/cdef_tmpl.c).  The reference filters in place and
 * keeps backups so that every tap reads pre-CDEF pixels; a second picture gives the same without backups.
 *   masks = f->lf.mask (device): Av1Filter.cdef_idx (strength index per 64x64, -1 = unset) and
 *           Av1Filter.noskip_mask (blocks with coefficients, decode.c:1990-1999);
 *   y_strength / uv_strength / damping = frame_hdr->cdef.
 * Blocks the reference leaves alone are copied.  Asynchronous on the context's stream. */
typedef struct Dav1dCudaCdefFrame {
    int32_t bw, bh;                     /* f->bw, f->bh: frame size in 4-px units, rounded up to 8 pixels */
    int32_t sb128w;                     /* f->sb128w */
    int32_t damping;                    /* frame_hdr->cdef.damping, 3..6 */
    uint8_t y_strength[8], uv_strength[8];
    const void *masks;                  /* device */
} Dav1dCudaCdefFrame;
DAV1D_CUDA_API int dav1d_cuda_cdef_frame(Dav1dCudaContext *c, const Dav1dCudaPicture *dst,
                                         const Dav1dCudaPicture *src, const Dav1dCudaCdefFrame *p);

/* Third stage: loop restoration (Wiener / self-guided) of a whole frame, OUT OF PLACE.  Replaces dav1d_lr_sbrow
 * over all superblock rows (src/lr_apply_tmpl.c:165-202 -> lr_sbrow / lr_stripe -> dsp->lr.wiener[] / .sgr[],
 * src/looprestoration_tmpl.c).  src = the CDEF output; deblocked = the picture before CDEF (the two lines above
 * and below every 64-row stripe are read from it - what dav1d_copy_lpf saves into f->lf.lr_lpf_line,
 * lf_apply_tmpl.c:108-175); pass src twice when the frame has no CDEF.
 *   lr_mask = f->lf.lr_mask (device): Av1Restoration[f->sr_sb128w * f->sb128h] (src/lf_mask.h:40-46,61-63,
 *             108 bytes each); unit_size_log2 = frame_hdr->restoration.unit_size; restore_planes = f->lf.restore_planes.
 * 128x128 superblocks (sb128 != 0) need units of at least 128 luma pixels, which is what the frame header gives such
 * streams (obu.c:944-954): every 64-row stripe then finds the unit lr_sbrow looks up once per 128-row superblock
 * row (lr_apply_tmpl.c:137-143); smaller units with sb128: -EINVAL.
 * Super-resolution (frame_hdr->width[0] != width[1]): src = dav1d_cuda_resize_frame of the CDEF output, deblocked =
 * dav1d_cuda_resize_frame of the deblocked picture (the reference resizes the lines it backs up row by row,
 * lf_apply_tmpl.c:76-91 - the same pixels), w / sb128w = the upscaled width / f->sr_sb128w; bit-exact with
 * dav1d_filter_sbrow on eight super-resolved frames (tests/test_postfilter_chain.py, sr_* cases). */
typedef struct Dav1dCudaLrFrame {
    int32_t w, h;                       /* f->sr_cur.p.p.w / .h */
    int32_t sb128w;                     /* f->sr_sb128w */
    int32_t sb128;                      /* seq_hdr->sb128 */
    int32_t unit_size_log2[2];          /* luma, chroma */
    int32_t restore_planes;             /* bit 0 Y, 1 U, 2 V */
    const void *lr_mask;                /* device */
} Dav1dCudaLrFrame;
DAV1D_CUDA_API int dav1d_cuda_lr_frame(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, const Dav1dCudaPicture *src,
                                       const Dav1dCudaPicture *deblocked, const Dav1dCudaLrFrame *p);

/* Operator-class launches.  All pointers inside the argument list that are
 * documented as "device" must be device pointers; the calls are asynchronous
 * on the context's stream. */

/* itxfm_add over `n` descriptors (device array).  `cf` = device coefficient
 * stream.  Descriptors must be grouped by `tx` (class_count[t] of them for
 * each tx size t, in increasing t).  If `zero_coefs` the consumed coefficient
 * blocks are cleared like the per-call contract requires (itx_tmpl.c:89). */
DAV1D_CUDA_API int dav1d_cuda_itx_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst,
                                        void *cf, const Dav1dCudaItxDesc *descs,
                                        const int32_t class_count[DAV1D_CUDA_N_RECT_TX_SIZES],
                                        int zero_coefs);

/* All transform sizes in (at most) two launches.  A task = up to 32/G
 * consecutive descriptors of one size for one warp (G = lanes per block);
 * dav1d_cuda_itx_tasks() (host) builds the task codes for `n` descriptors that
 * are grouped by `tx` (any order of the groups): tasks[0 .. *n_small) cover the
 * sizes up to 16x16, the next *n_big the larger ones; `index_base` is added to
 * the descriptor indices (position of descs_host[0] inside the device array).
 * Returns the number of tasks (<= n). */
DAV1D_CUDA_API int dav1d_cuda_itx_tasks(const Dav1dCudaItxDesc *descs_host, int n, int index_base,
                                        uint32_t *tasks, int32_t *n_small, int32_t *n_big);
DAV1D_CUDA_API int dav1d_cuda_itx_task_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, void *cf,
                                             const Dav1dCudaItxDesc *descs, const uint32_t *tasks,
                                             int n_small, int n_big, int zero_coefs);

/* Motion compensation.  `tiles` (device) lists the 32x32 work tiles:
 * tiles[i] = desc_index * 16 + (tile_row * 4 + tile_col); the recorder emits
 * one entry per 32x32 (or smaller, at the block edge) tile of each block
 * (dav1d_cuda_mc_tiles() does it for one block).  The first `n_small` entries of a tile list
 * must be the tiles of blocks of at most 8x8 samples: they are processed four per warp.
 *  - put_batch: kinds PUT (-> dst picture) and PREP (-> int16 pool `tmp`).
 *  - compound_batch: kinds AVG / W_AVG / MASK / W_MASK, the two predictions
 *    and the combine fused in one kernel (the int16 intermediates never
 *    reach HBM).  `masks` (device) holds wedge masks (read by MASK) and
 *    receives the segmentation masks W_MASK emits; a MASK descriptor that
 *    consumes a mask emitted by a W_MASK descriptor must be submitted in a
 *    later launch (luma before chroma, recon_tmpl.c:1852-1905). */
DAV1D_CUDA_API int dav1d_cuda_mc_put_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst,
                                           const Dav1dCudaPicture *const refs[7],
                                           const Dav1dCudaMcDesc *descs, const uint32_t *tiles,
                                           int n_tiles, int n_small, int16_t *tmp);
DAV1D_CUDA_API int dav1d_cuda_mc_compound_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst,
                                                const Dav1dCudaPicture *const refs[7],
                                                const Dav1dCudaMcDesc *descs, const uint32_t *tiles,
                                                int n_tiles, int n_small, uint8_t *masks);
/* Host helper: append the tile codes of block `desc_index` (w x h) to `out`
 * (room for 16 entries); returns the number written. */
DAV1D_CUDA_API int dav1d_cuda_mc_tiles(uint32_t desc_index, int w, int h, uint32_t *out);

DAV1D_CUDA_API int dav1d_cuda_warp_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst,
                                         const Dav1dCudaPicture *const refs[7],
                                         const Dav1dCudaWarpDesc *descs, int n);

/* Super-resolution of a frame on the device: dav1d_filter_sbrow_resize (recon_tmpl.c:2104-2137) over every
 * superblock row - mc.resize from `src` (the CDEF output, f->cur geometry) into `dst` (f->sr_cur geometry: the
 * upscaled width, same height), with f->resize_step[] / f->resize_start[] (luma, chroma; decode.c:3576-3583).
 * Out of place; rows are independent, so the frame is one launch per plane on the context's stream. */
DAV1D_CUDA_API int dav1d_cuda_resize_frame(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, const Dav1dCudaPicture *src,
                                           const int32_t resize_step[2], const int32_t resize_start[2]);

/* Intra-class operations are executed by ONE persistent launch per group of frames
 * (csrc/recon2.cu): the recorder hands the descriptors over in decode order and that is all - the
 * kernel finds the dependency levels itself, round by round, from a per-4x4-cell map in device
 * memory that counts the operations which still have to write the cell.  Nothing is scheduled on
 * the host.
 *
 * Cell map: dav1d_cuda_intra_cellmap_bytes() bytes of device memory per stream, zeroed ONCE by
 * the caller (every frame leaves it at zero again). */
DAV1D_CUDA_API size_t dav1d_cuda_intra_cellmap_bytes(int bw4, int bh4, int ss_hor, int ss_ver);
/* Optional recorder-side pass (host, linear in n, no device work): the dependency level of every
 * intra-class operation - one more than the highest level among the operations that wrote a pixel
 * it reads (0: nothing it reads is written in this phase) - from ONE walk over the descriptors in
 * decode order, exactly the order the recorder emits them in.  level + 1 is stored in the low 16
 * bits of `reserved`; set Dav1dCudaReconBatch.intra_levels_recorded.  The device only uses the
 * levels to ORDER its work (operations still wait for the pixels they read), so a wrong level costs
 * time or raises the status word, never a wrong pixel.  Returns the number of levels or -errno. */
DAV1D_CUDA_API int dav1d_cuda_intra_levels(Dav1dCudaIntraDesc *descs, int n, int bw4, int bh4,
                                           int ss_hor, int ss_ver);

/* Compact coefficient stream of a high-bit-depth frame (see Dav1dCudaReconBatch.cf_int16). */
typedef struct Dav1dCudaCoefEsc { uint32_t off; int32_t value; } Dav1dCudaCoefEsc;
/* Host: out16[i] = cf32[i] where it fits int16 (and is not -32768), else -32768 plus an entry of `esc`
 * (capacity cap_esc, offsets ascending).  Returns the number of escapes, or -ENOSPC. */
DAV1D_CUDA_API int dav1d_cuda_pack_coefs(const int32_t *cf32, size_t n, int16_t *out16,
                                         Dav1dCudaCoefEsc *esc, int cap_esc);

/* A whole frame's reconstruction as device-resident batches:
 *   phase A  motion compensation (put, fused compound in two waves, warp)
 *   phase B  inter residuals (itxfm_add per size class)
 *   phase C  intra-class operations: residual pre-pass, cell-map set-up, one persistent executor launch.
 * Replaces pass 2 (DAV1D_TASK_TYPE_TILE_RECONSTRUCTION, thread_task.c:757-761)
 * for one frame.  Asynchronous on the context's stream. */
typedef struct Dav1dCudaReconBatch {
    const Dav1dCudaPicture *dst;
    const Dav1dCudaPicture *refs[7];
    int32_t bw4, bh4;                 /* f->bw, f->bh: frame size in 4-px units */
    void *cf;                         /* device coefficient stream */
    uint8_t *masks;                   /* device: wedge masks / emitted seg masks */
    const void *pal;                  /* device: palette pixels */
    const uint8_t *pal_idx;           /* device: packed palette indices */
    const Dav1dCudaMcDesc *mc_put;    const uint32_t *mc_put_tiles;  int32_t n_mc_put_tiles;
    const Dav1dCudaMcDesc *mc_comp;   const uint32_t *mc_comp_tiles; int32_t n_mc_comp_tiles[2];
    /* leading small (<= 8x8) tiles of mc_put_tiles and of each compound wave */
    int32_t n_mc_put_small;           int32_t n_mc_comp_small[2];
    /* optional: OBMC blends, run after the prediction launches and before the residuals in two
     * waves - [0] all OBMC_H tiles, [1] all OBMC_V tiles (the two regions of a block overlap in
     * its top-left corner and the reference blends top neighbours first) */
    const Dav1dCudaMcDesc *mc_obmc;   const uint32_t *mc_obmc_tiles; int32_t n_mc_obmc_tiles[2];
    const Dav1dCudaWarpDesc *warp;    int32_t n_warp;
    const Dav1dCudaItxDesc *itx;      int32_t itx_class_count[DAV1D_CUDA_N_RECT_TX_SIZES];
    /* optional (device): task codes from dav1d_cuda_itx_tasks() over `itx`; when set phase B is two
     * launches (small / large sizes) instead of one per size */
    const uint32_t *itx_tasks;        int32_t n_itx_tasks[2];
    /* intra-class operations in decode order (device); intra_cellmap: see
     * dav1d_cuda_intra_cellmap_bytes() */
    const Dav1dCudaIntraDesc *intra;  int32_t n_intra;
    uint8_t *intra_cellmap;
    /* The residuals of the intra-class operations (eob >= 0) do not depend on any neighbour: the
     * recorder lists them a second time as plain transform descriptors, grouped by size exactly
     * like `itx` (same coef_off / eob / cw4 / ch4 as in the Dav1dCudaIntraDesc), and a pre-pass that
     * runs next to the motion compensation writes their inverse transforms into the int16
     * residual planes `intra_res` (a picture of the frame's geometry allocated with
     * bitdepth_max = 0xffff); the executor adds them to its predictions.  intra_itx_tasks optional
     * like itx_tasks. */
    const Dav1dCudaItxDesc *intra_itx; int32_t intra_itx_class_count[DAV1D_CUDA_N_RECT_TX_SIZES];
    const uint32_t *intra_itx_tasks;   int32_t n_intra_itx_tasks[2];
    const Dav1dCudaPicture *intra_res;
    /* non-zero: the recorder ran dav1d_cuda_intra_levels() over `intra` (the descriptors carry their
     * dependency level); zero: the device works the levels out itself (a cluster of blocks per frame) */
    int32_t intra_levels_recorded;
    /* optional: predictions from references of another size, four consecutive sections of
     * `mc_scaled`: [0] PUT / AVG / W_AVG / wedge MASK / W_MASK, [1] MASK descriptors that read a
     * mask section [0] emitted (chroma of a segmentation-mask block), [2] OBMC_H, [3] OBMC_V.
     * [0] and [1] run after the same-size predictions, [2] right after the same-size OBMC_H wave
     * and [3] after the same-size OBMC_V wave (a block's top blends come before its left blends
     * whatever the size of the neighbours' references). */
    const Dav1dCudaMcScaledDesc *mc_scaled; int32_t n_mc_scaled[4];
    /* high bit depth only: non-zero = `cf` is the compact stream - int16 storage (coef_off counts int16
     * elements), half the upload and half the HBM read of the int32 layout the reference keeps
     * (src/internal.h:288).  The value -32768 marks a coefficient that does not fit (10/12-bit streams
     * reach 2^(bitdepth+7) - 1 in magnitude, cf_max, recon_tmpl.c:594): its value is in `cf_esc` (device),
     * sorted by stream offset; dav1d_cuda_pack_coefs() produces both from an int32 stream. */
    int32_t cf_int16;
    const Dav1dCudaCoefEsc *cf_esc;   int32_t n_cf_esc;
} Dav1dCudaReconBatch;

enum { DAV1D_CUDA_MAX_GROUP = 64 };   /* frames per group submission */

DAV1D_CUDA_API int dav1d_cuda_recon_submit(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b);
/* Only the launch classes selected by phase_mask (bit0 put + OBMC, bit1 compound, bit2 warp, bit3
 * inter residual, bit4 intra), exactly as dav1d_cuda_recon_submit() would launch them - used to
 * time one class with CUDA events. */
DAV1D_CUDA_API int dav1d_cuda_recon_submit_phases(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b,
                                                  int phase_mask);
/* Frames of `n` (<= DAV1D_CUDA_MAX_GROUP) independent streams, same pixel type, in one submission:
 * MC / residual launches per frame on parallel branches and ONE intra executor launch for the
 * whole group (server-side batching of independent decoder instances).  No frame of the group may
 * use another member's destination as a reference (-EINVAL).  Nothing is prepared on the host:
 * the call only launches; a fresh set of descriptors costs the same as a repeated one. */
DAV1D_CUDA_API int dav1d_cuda_recon_group_submit(Dav1dCudaContext *c,
                                                 const Dav1dCudaReconBatch *const *batches, int n);
DAV1D_CUDA_API int dav1d_cuda_recon_group_submit_phases(Dav1dCudaContext *c,
                                                        const Dav1dCudaReconBatch *const *batches, int n,
                                                        int phase_mask);
/* The same launches captured into a CUDA graph once and replayed: _graph_build*() record,
 * dav1d_cuda_recon_graph_launch() replays on the context's stream (the device buffers the batches
 * point to may be rewritten between launches, their addresses and counts may not). */
typedef struct Dav1dCudaReconGraph Dav1dCudaReconGraph;
DAV1D_CUDA_API int dav1d_cuda_recon_graph_build(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b,
                                                Dav1dCudaReconGraph **out);
DAV1D_CUDA_API int dav1d_cuda_recon_graph_build_multi(Dav1dCudaContext *c,
                                                      const Dav1dCudaReconBatch *const *batches, int n,
                                                      Dav1dCudaReconGraph **out);
DAV1D_CUDA_API int dav1d_cuda_recon_graph_build_multi_phases(Dav1dCudaContext *c,
                                                             const Dav1dCudaReconBatch *const *batches, int n,
                                                             int phase_mask, Dav1dCudaReconGraph **out);
DAV1D_CUDA_API int dav1d_cuda_recon_graph_launch(Dav1dCudaContext *c, Dav1dCudaReconGraph *g);
DAV1D_CUDA_API void dav1d_cuda_recon_graph_free(Dav1dCudaReconGraph *g);

/* Plain device memory for descriptor arrays / coefficient streams. */
DAV1D_CUDA_API void *dav1d_cuda_malloc(size_t bytes);
DAV1D_CUDA_API void  dav1d_cuda_free(void *p);
DAV1D_CUDA_API int   dav1d_cuda_upload(Dav1dCudaContext *c, void *dev, const void *host, size_t bytes);
DAV1D_CUDA_API int   dav1d_cuda_download(Dav1dCudaContext *c, void *host, const void *dev, size_t bytes);
DAV1D_CUDA_API int   dav1d_cuda_memset(Dav1dCudaContext *c, void *dev, int value, size_t bytes);
DAV1D_CUDA_API void *dav1d_cuda_host_alloc(size_t bytes);   /* pinned */
DAV1D_CUDA_API void  dav1d_cuda_host_free(void *p);
/* CUDA-event timing on the context's stream (bench.py times the launching stream). */
DAV1D_CUDA_API void *dav1d_cuda_event_create(void);
DAV1D_CUDA_API int   dav1d_cuda_event_record(Dav1dCudaContext *c, void *ev);
DAV1D_CUDA_API int   dav1d_cuda_stream_wait_event(Dav1dCudaContext *c, void *ev);   /* orders c's stream after ev */
DAV1D_CUDA_API float dav1d_cuda_event_elapsed_ms(void *start, void *stop);   /* syncs on stop */
DAV1D_CUDA_API void  dav1d_cuda_event_destroy(void *ev);

#ifdef __cplusplus
}
#endif
#endif /* DAV1D_CUDA_H */
