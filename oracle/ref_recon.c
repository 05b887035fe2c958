/*
 * oracle/ref_recon.c - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Frame-level checker that runs the reference's OWN reconstruction driver:
 * dav1d_recon_b_intra_{8,16}bpc (src/recon_tmpl.c:1195-1596, compiled where it
 * lies under /root/reference by oracle/Makefile) is called block by block, in
 * decode order, on a frame context laid out the way pass 2 of frame threading
 * finds it (src/internal.h:276-293: cbi, cf, pal, pal_idx handed over by pass 1;
 * src/decode.c:741-772 for the block-context updates after each block;
 * src/decode.c:2677 for the per-superblock-row edge backup).
 *
 * The blocks come from the synthetic generator's block records (D1SynthBlock,
 * dav1d-mirror_b200/csrc/synth.cpp with real_blocks = 1): the Av1Block fields
 * the driver reads.  Nothing here decides a DSP call: which predictor, which
 * edge flags per transform block, CfL / palette order, tile edges, the
 * superblock-row edge backup - all of that is the reference's code.  Only the
 * coefficient / index streams are re-packed into the layout pass 1 leaves them
 * in.  Compiled twice (BITDEPTH 8 / 16); this file is this repo's own code.
 */
#include "config.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "common/attributes.h"
#include "common/bitdepth.h"
#include "common/intops.h"
#include "src/internal.h"
#include "src/levels.h"
#include "src/intra_edge.h"
#include "src/tables.h"
#include "src/recon.h"
#include "src/mc.h"
#include "src/itx.h"
#include "src/ipred.h"
#include "src/wedge.h"

#include "dav1d_cuda.h"

#define EXPORT __attribute__((visibility("default")))

/* == D1SynthBlock of dav1d-mirror_b200/csrc/synth.cpp */
typedef struct D1SynthBlock {
    uint16_t bx4, by4;
    uint8_t  w4, h4;
    uint8_t  intra, has_chroma, skip, tile;
    uint8_t  edge_tr, edge_bl;
    uint8_t  y_mode, uv_mode;
    int8_t   y_angle, uv_angle;
    uint8_t  tx, uvtx;
    uint8_t  pal_sz[2];
    int8_t   cfl_alpha[2];
    uint16_t tile_x0, tile_y0, tile_x1, tile_y1;
    uint32_t pal_off[3];
    uint32_t pal_idx_off[2];
    uint32_t first_op, n_ops;
    uint8_t  sm_flags, pad[3];
    int16_t  mvx[2], mvy[2];
    uint8_t  ref[2];
    uint8_t  comp_kind;
    uint8_t  filter2d, mask_sign, max_ytx, tx_split, jnt_weight;
    uint32_t first_tx, n_tx;
    struct { int32_t matrix[6]; int16_t abcd[4]; } warp;    /* comp_kind == 255: t->warpmv of an MM_WARP block */
} D1SynthBlock;
typedef struct D1SynthTx {
    uint32_t coef_off;
    int16_t  eob;
    uint8_t  txtp, cw4, ch4, tx, plane, pad;
} D1SynthTx;

typedef struct OracleReconFrame {
    void *dst[3];
    ptrdiff_t dst_stride[3];
    int32_t w, h, ss_hor, ss_ver, bitdepth_max, no_chroma, intra_edge_filter;
    const D1SynthBlock *blocks;
    int32_t n_blocks;
    const Dav1dCudaIntraDesc *ops;     /* the blocks' operations: eob / txtp / coefficient offsets per transform block */
    const void *cf;                    /* generator coefficient stream (packed or dense blocks) */
    const void *pal;                   /* palette pool */
    const uint8_t *pal_idx;            /* packed index pool */
    const void *ref[7][3];             /* reference pictures (inter blocks), same geometry as dst */
    ptrdiff_t ref_stride[7][2];
    int32_t n_refs;
    const D1SynthTx *tx_recs;          /* cbi / cf entries of the inter blocks */
    int32_t ref_w[7], ref_h[7];        /* luma size of the references (0 = the frame's): scaled prediction */
} OracleReconFrame;

#if BITDEPTH == 8
#define SUFFIX(name) name##_8bpc
#else
#define SUFFIX(name) name##_16bpc
#endif

static int bs_from_dims(const int w4, const int h4) {
    for (int bs = 0; bs < N_BS_SIZES; bs++)
        if (dav1d_block_dimensions[bs][0] == w4 && dav1d_block_dimensions[bs][1] == h4) return bs;
    return -1;
}

/* one transform block of the generator's stream -> the dense min(w,32) x min(h,32) block pass 1 stores */
static void dense_coefs2(coef *out, const OracleReconFrame *fr, const int tx, const uint32_t coef_off,
                         const int cw4, const int ch4)
{
    const TxfmInfo *const t = &dav1d_txfm_dimensions[tx];
    const int sw = imin(t->w * 4, 32), sh = imin(t->h * 4, 32);
    const coef *src = (const coef *)fr->cf + coef_off;
    memset(out, 0, sizeof(coef) * sw * sh);
    if (!cw4 || !ch4) {
        memcpy(out, src, sizeof(coef) * sw * sh);
    } else {
        const int cw = cw4 * 4, ch = ch4 * 4;
        for (int x = 0; x < cw; x++)
            for (int y = 0; y < ch; y++) out[y + x * sh] = src[y + x * ch];
    }
}
static void dense_coefs(coef *out, const OracleReconFrame *fr, const Dav1dCudaIntraDesc *d) {
    dense_coefs2(out, fr, d->tx, d->coef_off, d->cw4, d->ch4);
}

/* Reconstructs every block of the frame through dav1d_recon_b_intra / dav1d_recon_b_inter.  Returns 0, or
 * a negative value when the records hold something this harness does not drive (global motion). */
EXPORT int SUFFIX(oracle_recon_frame)(const OracleReconFrame *const fr) {
    int ret = 0;
    dav1d_init_ii_wedge_masks();                   /* src/wedge.c: what dav1d_init_once does (lib.c) */
    Dav1dDSPContext dsp;
    memset(&dsp, 0, sizeof(dsp));
    SUFFIX(dav1d_mc_dsp_init)(&dsp.mc);
    SUFFIX(dav1d_itx_dsp_init)(&dsp.itx, fr->bitdepth_max > 1023 ? 12 : fr->bitdepth_max > 255 ? 10 : 8);
    SUFFIX(dav1d_intra_pred_dsp_init)(&dsp.ipred);

    Dav1dSequenceHeader seq;
    Dav1dFrameHeader hdr;
    memset(&seq, 0, sizeof(seq));
    memset(&hdr, 0, sizeof(hdr));
    seq.intra_edge_filter = fr->intra_edge_filter;
    seq.hbd = fr->bitdepth_max > 1023 ? 2 : fr->bitdepth_max > 255 ? 1 : 0;
    hdr.frame_type = fr->n_refs > 0 ? DAV1D_FRAME_TYPE_INTER : DAV1D_FRAME_TYPE_KEY;

    Dav1dFrameContext *const f = calloc(1, sizeof(*f));
    Dav1dTileState *const ts = calloc(1, sizeof(*ts));
    Dav1dTaskContext *t = NULL;
    if (posix_memalign((void **)&t, 64, sizeof(*t))) { free(f); free(ts); return -12; }
    memset(t, 0, sizeof(*t));
    f->seq_hdr = &seq;
    f->frame_hdr = &hdr;
    f->dsp = &dsp;
    f->bitdepth_max = fr->bitdepth_max;
    f->cur.data[0] = fr->dst[0]; f->cur.data[1] = fr->dst[1]; f->cur.data[2] = fr->dst[2];
    f->cur.stride[0] = fr->dst_stride[0]; f->cur.stride[1] = fr->dst_stride[1];
    f->cur.p.w = fr->w; f->cur.p.h = fr->h;
    f->cur.p.bpc = fr->bitdepth_max > 1023 ? 12 : fr->bitdepth_max > 255 ? 10 : 8;
    f->cur.p.layout = fr->no_chroma ? DAV1D_PIXEL_LAYOUT_I400 :
                      fr->ss_ver ? DAV1D_PIXEL_LAYOUT_I420 : fr->ss_hor ? DAV1D_PIXEL_LAYOUT_I422 : DAV1D_PIXEL_LAYOUT_I444;
    const int ss_hor = !fr->no_chroma && fr->ss_hor, ss_ver = !fr->no_chroma && fr->ss_ver;
    /* decode.c:3557-3570 */
    f->bw = ((fr->w + 7) >> 3) << 1;
    f->bh = ((fr->h + 7) >> 3) << 1;
    f->sb128w = (f->bw + 31) >> 5;
    f->sb128h = (f->bh + 31) >> 5;
    f->sb_shift = 4;
    f->sb_step = 16;
    f->sbh = (f->bh + f->sb_step - 1) >> f->sb_shift;
    f->b4_stride = (f->bw + 31) & ~31;
    /* the superblock-row edge backups (decode.c:3045-3060) */
    const size_t edge_px = (size_t)f->sb128w * 128 * (f->sbh + 1);
    pixel *const edge_buf = calloc(3 * edge_px, sizeof(pixel));
    f->ipred_edge[0] = edge_buf; f->ipred_edge[1] = edge_buf + edge_px; f->ipred_edge[2] = edge_buf + 2 * edge_px;
    /* palettes per 8x8 (internal.h:285) */
    const size_t n_pal = (size_t)(f->b4_stride >> 1) * (size_t)((f->bh + 33) >> 1);
    f->frame_thread.pal = calloc(n_pal, sizeof(*f->frame_thread.pal));
    BlockContext *const a = calloc((size_t)f->sb128w + 1, sizeof(*a));
    f->a = a;
    /* reference pictures (f->svc stays zero for those of the frame's size) and the weights of
     * COMP_INTER_WEIGHTED_AVG, a table of the frame (the generator draws the same table) */
    for (int i = 0; i < fr->n_refs && i < 7; i++) {
        Dav1dPicture *const rp = &f->refp[i].p;
        rp->data[0] = (void *)fr->ref[i][0]; rp->data[1] = (void *)fr->ref[i][1]; rp->data[2] = (void *)fr->ref[i][2];
        rp->stride[0] = fr->ref_stride[i][0]; rp->stride[1] = fr->ref_stride[i][1];
        rp->p = f->cur.p;
        /* a reference of another size: decode.c:3511-3527 */
        if (fr->ref_w[i]) rp->p.w = fr->ref_w[i];
        if (fr->ref_h[i]) rp->p.h = fr->ref_h[i];
        if (rp->p.w != f->cur.p.w || rp->p.h != f->cur.p.h) {
#define scale_fac(ref_sz, this_sz) ((((ref_sz) << 14) + ((this_sz) >> 1)) / (this_sz))
            f->svc[i][0].scale = scale_fac(rp->p.w, f->cur.p.w);
            f->svc[i][1].scale = scale_fac(rp->p.h, f->cur.p.h);
            f->svc[i][0].step = (f->svc[i][0].scale + 8) >> 4;
            f->svc[i][1].step = (f->svc[i][1].scale + 8) >> 4;
#undef scale_fac
        }
    }
    for (int i = 0; i < 7; i++)
        for (int j = 0; j < 7; j++) f->jnt_weights[i][j] = 1 + (i * 7 + j * 3 + 4) % 15;

    /* intrabc predicts from the picture being decoded: mc() is handed &f->sr_cur (recon_tmpl.c:1624-1637) */
    f->sr_cur.p = f->cur;
    t->f = f;
    t->ts = ts;
    t->frame_thread.pass = 2;
    /* the refmvs rows obmc() reads its neighbours from (recon_tmpl.c:1078): one block array for the whole
     * frame, 5 rows of margin above; every block writes its area after reconstruction the way
     * decode.c:756-767 / 815-826 do through splat_mv (bottom row and right column are what is read) */
    const int rstride = f->b4_stride + 32;
    refmvs_block *const rmv = calloc((size_t)(f->bh + 64 + 5) * rstride, sizeof(*rmv));

    /* pass 2 reads the filter of a sub8x8 block's neighbours from the frame's Av1Block array (recon_tmpl.c:1710,1726,1741) */
    Av1Block *const fb = calloc((size_t)f->b4_stride * (f->bh + 32), sizeof(*fb));
    f->frame_thread.b = fb;

    /* per-block streams: pass 1 leaves cbi / cf / pal_idx as consecutive runs; a block's share is
     * built right before the call */
    int16_t cbi[3 * 256 + 8];
    coef *const cfbuf = aligned_alloc(64, sizeof(coef) * 64 * 1024);
    uint8_t *const idxbuf = aligned_alloc(64, 8192);
    if (!edge_buf || !f->frame_thread.pal || !a || !cfbuf || !idxbuf || !rmv || !fb) { ret = -12; goto done; }

    int cur_tile = -1, cur_sbrow = -1;
    for (int i = 0; i < fr->n_blocks; i++) {
        const D1SynthBlock *const s = &fr->blocks[i];
        if (!s->intra && s->comp_kind > DAV1D_CUDA_MC_W_MASK && s->comp_kind < 254) { ret = -38; goto done; }
        if (!s->intra && s->comp_kind == 254 && (hdr.frame_type & 1)) { ret = -22; goto done; }   /* intrabc: key / intra-only frames (common/frame.h:42) */
        const int sbrow = s->by4 >> f->sb_shift;
        if (s->tile != cur_tile || sbrow != cur_sbrow) {
            if (cur_tile >= 0) {            /* decode.c:2677: end of a tile's superblock row */
                t->by = cur_sbrow << f->sb_shift;
                SUFFIX(dav1d_backup_ipred_edge)(t);
            }
            if (s->tile != cur_tile) {      /* decode.c:2593-2600, dav1d_reset_context: above context of the tile */
                for (int x = s->tile_x0 >> 5; x <= (s->tile_x1 - 1) >> 5; x++) memset(&a[x], 0, sizeof(a[x]));
            }
            memset(&t->l, 0, sizeof(t->l)); /* decode.c:2560-2565: left context at the start of a superblock row */
            ts->tiling.col_start = s->tile_x0; ts->tiling.col_end = s->tile_x1;
            ts->tiling.row_start = s->tile_y0; ts->tiling.row_end = s->tile_y1;
            cur_tile = s->tile; cur_sbrow = sbrow;
        }
        const int bs = bs_from_dims(s->w4, s->h4);
        if (bs < 0) { ret = -22; goto done; }
        if (!s->intra) {
            /* ---- inter block: dav1d_recon_b_inter (recon_tmpl.c:1598-2036) */
            Av1Block b;
            memset(&b, 0, sizeof(b));
            b.bs = bs; b.intra = 0; b.skip = s->skip; b.uvtx = s->uvtx;
            b.comp_type = s->comp_kind == DAV1D_CUDA_MC_PUT || s->comp_kind >= 254 ? COMP_INTER_NONE :
                          s->comp_kind == DAV1D_CUDA_MC_AVG ? COMP_INTER_AVG :
                          s->comp_kind == DAV1D_CUDA_MC_W_AVG ? COMP_INTER_WEIGHTED_AVG :
                          s->comp_kind == DAV1D_CUDA_MC_MASK ? COMP_INTER_WEDGE : COMP_INTER_SEG;
            b.wedge_idx = s->pad[1];
            b.inter_mode = 0; b.motion_mode = s->pad[0] ? MM_OBMC : MM_TRANSLATION;
            if (s->comp_kind == 255) {
                /* local warp (pad[1] == 0): what decode_b() leaves in t->warpmv (decode.c:1828-1860); global
                 * motion (pad[1] == 1): a GLOBALMV block whose reference carries the model in the frame header
                 * (recon_tmpl.c:1641-1649; the header's model is set right before the call - the records give
                 * every block its own) */
                Dav1dWarpedMotionParams wm;
                memset(&wm, 0, sizeof(wm));
                wm.type = DAV1D_WM_TYPE_AFFINE;
                for (int k = 0; k < 6; k++) wm.matrix[k] = s->warp.matrix[k];
                for (int k = 0; k < 4; k++) wm.u.abcd[k] = s->warp.abcd[k];
                memset(&t->warpmv, 0, sizeof(t->warpmv));
                memset(f->gmv_warp_allowed, 0, sizeof(f->gmv_warp_allowed));
                b.wedge_idx = 0;
                if (s->pad[1]) {
                    b.inter_mode = GLOBALMV;
                    hdr.gmv[s->ref[0]] = wm;
                    f->gmv_warp_allowed[s->ref[0]] = 1;
                } else {
                    b.motion_mode = MM_WARP;
                    t->warpmv = wm;
                }
            }
            b.interintra_type = s->pad[2] & 3; b.interintra_mode = s->pad[2] >> 2;     /* INTER_INTRA_BLEND / _WEDGE + II_*_PRED */
            for (int r = 0; r < 32 + 5; r++)
                t->rt.r[r] = rmv + (size_t)((s->by4 & ~31) + r) * rstride;      /* row (by & ~31) - 5 + r, margin 5 */
            for (int k = 0; k < 2; k++) { b.mv[k].x = s->mvx[k]; b.mv[k].y = s->mvy[k]; b.ref[k] = (int8_t)s->ref[k]; }
            b.filter2d = s->filter2d; b.mask_sign = s->mask_sign;
            b.max_ytx = s->max_ytx; b.tx_split0 = s->tx_split ? 1 : 0; b.tx_split1 = 0;
            if (f->jnt_weights[b.ref[0]][b.ref[1]] != s->jnt_weight && s->comp_kind == DAV1D_CUDA_MC_W_AVG) { ret = -22; goto done; }
            t->bx = s->bx4; t->by = s->by4;
            t->a = &a[s->bx4 >> 5];
            int n_cbi = 0;
            coef *cfp = cfbuf;
            for (unsigned k = 0; k < s->n_tx; k++) {
                const D1SynthTx *const x = &fr->tx_recs[s->first_tx + k];
                const TxfmInfo *const td = &dav1d_txfm_dimensions[x->tx];
                cbi[n_cbi++] = (int16_t)((x->eob << 5) + x->txtp);
                if (x->eob >= 0) dense_coefs2(cfp, fr, x->tx, x->coef_off, x->cw4, x->ch4);
                /* luma: recon_tmpl.c:776-778; chroma: :1951 */
                cfp += x->plane ? td->w * td->h * 16 : imin(td->w, 8) * imin(td->h, 8) * 16;
            }
            ts->frame_thread[0].cbi = cbi;
            ts->frame_thread[0].cf = cfbuf;
            if (SUFFIX(dav1d_recon_b_inter)(t, bs, &b)) { ret = -5; goto done; }
            /* decode.c:815-826 (splat_oneref_mv / splat_tworef_mv) */
            for (int y = 0; y < s->h4 && s->by4 + y < f->bh + 32; y++)
                for (int x = 0; x < s->w4 && s->bx4 + x < f->b4_stride + 32; x++) {
                    refmvs_block *const rb = &rmv[(size_t)(s->by4 + y + 5) * rstride + s->bx4 + x];
                    rb->mv.mv[0] = b.mv[0]; rb->mv.mv[1] = b.mv[1];
                    rb->ref.ref[0] = b.ref[0] + 1;
                    rb->ref.ref[1] = b.comp_type == COMP_INTER_NONE ? -1 : b.ref[1] + 1;
                    rb->bs = bs; rb->mf = 0;
                    if (fb && s->by4 + y < f->bh + 32 && s->bx4 + x < f->b4_stride)
                        fb[(size_t)(s->by4 + y) * f->b4_stride + s->bx4 + x].filter2d = b.filter2d;
                }
            /* decode.c:808-830: filters, intra = 0, uvmode = DC_PRED into the contexts */
            const int bx4 = s->bx4 & 31, by4 = s->by4 & 31;
            const uint8_t *const filter = dav1d_filter_dir[b.filter2d];
            for (int x = 0; x < s->w4; x++) { t->a->filter[0][bx4 + x] = filter[0]; t->a->filter[1][bx4 + x] = filter[1]; t->a->intra[bx4 + x] = 0; }
            for (int y = 0; y < s->h4; y++) { t->l.filter[0][by4 + y] = filter[0]; t->l.filter[1][by4 + y] = filter[1]; t->l.intra[by4 + y] = 0; }
            if (s->has_chroma) {
                const int cbx4 = bx4 >> ss_hor, cby4 = by4 >> ss_ver;
                for (int x = 0; x < (s->w4 + ss_hor) >> ss_hor; x++) t->a->uvmode[cbx4 + x] = DC_PRED;
                for (int y = 0; y < (s->h4 + ss_ver) >> ss_ver; y++) t->l.uvmode[cby4 + y] = DC_PRED;
            }
            continue;
        }
        Av1Block b;
        memset(&b, 0, sizeof(b));
        b.bs = bs; b.intra = 1; b.skip = s->skip; b.uvtx = s->uvtx;
        b.y_mode = s->y_mode; b.uv_mode = s->uv_mode; b.tx = s->tx;
        b.pal_sz[0] = s->pal_sz[0]; b.pal_sz[1] = s->pal_sz[1];
        b.y_angle = s->y_angle; b.uv_angle = s->uv_angle;
        b.cfl_alpha[0] = s->cfl_alpha[0]; b.cfl_alpha[1] = s->cfl_alpha[1];
        t->bx = s->bx4; t->by = s->by4;
        t->a = &a[s->bx4 >> 5];

        /* the block's cbi / cf run: one entry per transform block with a residual field, in the
         * order the block's operations were recorded (luma raster, then U raster, then V raster) */
        int n_cbi = 0;
        coef *cfp = cfbuf;
        for (unsigned k = 0; k < s->n_ops; k++) {
            const Dav1dCudaIntraDesc *const d = &fr->ops[s->first_op + k];
            if (d->mode == DAV1D_CUDA_INTRA_PAL) continue;         /* whole-block palette prediction: no coefficients */
            if (s->skip) continue;
            const TxfmInfo *const td = &dav1d_txfm_dimensions[d->tx];
            cbi[n_cbi++] = (int16_t)((d->eob << 5) + d->txtp);
            if (d->eob >= 0) dense_coefs(cfp, fr, d);
            cfp += imin(td->w, 8) * imin(td->h, 8) * 16;
        }
        ts->frame_thread[0].cbi = cbi;
        ts->frame_thread[0].cf = cfbuf;
        /* palette colours and indices where pass 1 puts them (recon_tmpl.c:1238-1241, 1430-1434) */
        if (s->pal_sz[0] || s->pal_sz[1]) {
            pixel (*const pal)[8] = f->frame_thread.pal[((s->by4 >> 1) + (s->bx4 & 1)) * (f->b4_stride >> 1) +
                                                         ((s->bx4 >> 1) + (s->by4 & 1))];
            for (int pl = 0; pl < 3; pl++)
                if (s->pal_sz[pl ? 1 : 0])
                    memcpy(pal[pl], (const pixel *)fr->pal + s->pal_off[pl], 8 * sizeof(pixel));
            uint8_t *ip = idxbuf;
            if (s->pal_sz[0]) {
                memcpy(ip, fr->pal_idx + s->pal_idx_off[0], s->w4 * s->h4 * 8);
                ip += s->w4 * s->h4 * 8;
            }
            if (s->pal_sz[1] && s->has_chroma) {
                const int cbw4 = (s->w4 + ss_hor) >> ss_hor, cbh4 = (s->h4 + ss_ver) >> ss_ver;
                memcpy(ip, fr->pal_idx + s->pal_idx_off[1], cbw4 * cbh4 * 8);
            }
            ts->frame_thread[0].pal_idx = idxbuf;
        }
        /* block-level edge flags as decode_b() gets them from the partition tree (intra_edge.h) */
        const int ef = ((s->edge_tr & 1) ? EDGE_I444_TOP_HAS_RIGHT : 0) | ((s->edge_bl & 1) ? EDGE_I444_LEFT_HAS_BOTTOM : 0) |
                       ((s->edge_tr & 2) ? (EDGE_I420_TOP_HAS_RIGHT | EDGE_I422_TOP_HAS_RIGHT) : 0) |
                       ((s->edge_bl & 2) ? (EDGE_I420_LEFT_HAS_BOTTOM | EDGE_I422_LEFT_HAS_BOTTOM) : 0);
        SUFFIX(dav1d_recon_b_intra)(t, bs, ef, &b);

        /* decode.c:756-767 (splat_intraref): ref 0 = intra */
        for (int y = 0; y < s->h4 && s->by4 + y < f->bh + 32; y++)
            for (int x = 0; x < s->w4 && s->bx4 + x < f->b4_stride + 32; x++) {
                refmvs_block *const rb = &rmv[(size_t)(s->by4 + y + 5) * rstride + s->bx4 + x];
                memset(rb, 0, sizeof(*rb));
                rb->ref.ref[1] = -1; rb->bs = bs;
            }
        /* decode.c:744-772: the block's modes into the above / left contexts */
        const int bx4 = s->bx4 & 31, by4 = s->by4 & 31;
        const int y_mode_nofilt = s->y_mode == FILTER_PRED ? DC_PRED : s->y_mode;
        for (int x = 0; x < s->w4; x++) { t->a->mode[bx4 + x] = y_mode_nofilt; t->a->intra[bx4 + x] = 1; }
        for (int y = 0; y < s->h4; y++) { t->l.mode[by4 + y] = y_mode_nofilt; t->l.intra[by4 + y] = 1; }
        if (s->has_chroma) {
            const int cbx4 = bx4 >> ss_hor, cby4 = by4 >> ss_ver;
            const int cbw4 = (s->w4 + ss_hor) >> ss_hor, cbh4 = (s->h4 + ss_ver) >> ss_ver;
            for (int x = 0; x < cbw4; x++) t->a->uvmode[cbx4 + x] = s->uv_mode;
            for (int y = 0; y < cbh4; y++) t->l.uvmode[cby4 + y] = s->uv_mode;
        }
    }
done:
    free(rmv); free(f->frame_thread.b); free(idxbuf); free(cfbuf); free(a); free(f->frame_thread.pal); free(edge_buf);
    free(t); free(ts); free(f);
    return ret;
}

#if BITDEPTH == 8
/* The decoder's wedge / inter-intra masks (src/wedge.c tables through the WEDGE_MASK / II_MASK macros of
 * src/wedge.h:88-97) for the generator's real-block frames.  lay: 0 = 4:4:4 / luma, 1 = 4:2:2, 2 = 4:2:0. */
static void masks_once(void) {
    static int done;
    if (!done) { dav1d_init_ii_wedge_masks(); done = 1; }
}
EXPORT const uint8_t *oracle_wedge_mask(const int lay, const int w4, const int h4, const int sign, const int idx) {
    const int bs = bs_from_dims(w4, h4);
    if (bs < BS_32x32 || bs > BS_8x8 || lay < 0 || lay > 2 || idx < 0 || idx > 15) return NULL;
    masks_once();
    return WEDGE_MASK(lay, bs, sign & 1, idx);
}
EXPORT const uint8_t *oracle_ii_mask(const int lay, const int w4, const int h4, const int mode) {
    const int bs = bs_from_dims(w4, h4);
    if (bs < BS_32x32 || bs > BS_8x8 || lay < 0 || lay > 2 || mode < 0 || mode > 3) return NULL;
    masks_once();
    return (const uint8_t *) &dav1d_masks + (size_t) dav1d_masks.offsets[lay][bs - BS_32x32].ii[mode] * 8;
}

/* recon_tmpl.c also holds the inter and post-filter drivers; what they call outside the files this
 * checker compiles is never reached from dav1d_recon_b_intra */
#define STUB(name) void name(void) { abort(); }
#include "ref_recon_stubs.h"
#endif
