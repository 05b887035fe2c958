/*
 * oracle/ref_frame.c - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Frame-level checker: walks the same block descriptors the CUDA backend
 * consumes (include/dav1d_cuda.h), but sequentially, in decode order, on host
 * memory, calling the REFERENCE's own DSP function tables
 * (dav1d_{mc,itx,intra_pred}_dsp_init_*bpc, dav1d_prepare_intra_edges_*bpc)
 * the way src/recon_tmpl.c does:
 *   mc()            recon_tmpl.c:957-1011   (incl. the emu_edge decision :986-999)
 *   compound        recon_tmpl.c:1823-1868
 *   warp_affine()   recon_tmpl.c:1134-1193  (per-8x8 part :1169-1186)
 *   itxfm_add       recon_tmpl.c:816,1347,1567,2017
 *   intra / cfl / pal  recon_tmpl.c:1226-1300, 1372-1417, 1503-1540
 * Compiled twice (BITDEPTH 8 / 16) against the reference headers where they
 * lie; this file is this repo's own code.
 */
#include "config.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "common/attributes.h"
#include "common/bitdepth.h"
#include "common/intops.h"
#include "src/levels.h"
#include "src/mc.h"
#include "src/itx.h"
#include "src/ipred.h"
#include "src/ipred_prepare.h"

#include "dav1d_cuda.h"

#define EXPORT __attribute__((visibility("default")))

typedef struct OracleFrame {
    void *dst[3];
    ptrdiff_t dst_stride[3];
    const void *ref[7][3];
    ptrdiff_t ref_stride[7][3];
    int32_t w, h, ss_hor, ss_ver, bitdepth_max;
    int32_t bw4, bh4;
    const void *cf;
    uint8_t *masks;
    const void *pal;
    const uint8_t *pal_idx;
    const Dav1dCudaMcDesc *mc_put;
    const Dav1dCudaMcDesc *mc_comp;
    const Dav1dCudaWarpDesc *warp;
    const Dav1dCudaItxDesc *itx;
    const Dav1dCudaIntraDesc *intra;
    const uint32_t *order;        /* (class << 28) | index; class: 0 put 1 comp 2 warp 3 itx 4 intra 6 obmc 8 scaled */
    int32_t n_order;
    const Dav1dCudaMcDesc *mc_obmc;
    const Dav1dCudaMcScaledDesc *mc_scaled;   /* class 8: index into the four sections stored back to back */
    int32_t ref_w[7], ref_h[7];               /* luma size of the references (0 = the frame's) */
} OracleFrame;

#if BITDEPTH == 8
#define BD_ARG
#define BD_DECL
#else
#define BD_ARG , bdmax
#define BD_DECL const int bdmax = f->bitdepth_max;
#endif

typedef struct Scratch {
    ALIGN(pixel emu[320 * (256 + 7)], 64);
    ALIGN(int16_t tmp[2][128 * 128], 64);
    ALIGN(coef cf[32 * 32], 64);
    ALIGN(int16_t ac[32 * 32], 64);
    ALIGN(pixel edge_buf[257 + 63], 64);
} Scratch;

static const uint8_t tx_w4[19] = { 1, 2, 4, 8, 16, 1, 2, 2, 4, 4, 8, 8, 16, 1, 4, 2, 8, 4, 16 };
static const uint8_t tx_h4[19] = { 1, 2, 4, 8, 16, 2, 1, 4, 2, 8, 4, 16, 8, 4, 1, 8, 2, 16, 4 };

static void plane_dims(const OracleFrame *f, int pl, int *w, int *h) {
    const int sh = pl ? f->ss_hor : 0, sv = pl ? f->ss_ver : 0;
    *w = (f->w + sh) >> sh;
    *h = (f->h + sv) >> sv;
}

/* source pointer for one prediction, with the reference's emu_edge rule */
static const pixel *mc_src(const OracleFrame *f, Scratch *s, const Dav1dCudaMcSrc *src, int pl,
                           int bw, int bh, ptrdiff_t *stride, const Dav1dMCDSPContext *mc)
{
    int w, h;
    plane_dims(f, pl, &w, &h);
    const int mx = src->mx, my = src->my, dx = src->x, dy = src->y;
    const pixel *ref = f->ref[src->ref][pl];
    ptrdiff_t ref_stride = f->ref_stride[src->ref][pl];
    if (dx < !!mx * 3 || dy < !!my * 3 || dx + bw + !!mx * 4 > w || dy + bh + !!my * 4 > h) {
        mc->emu_edge(bw + !!mx * 7, bh + !!my * 7, w, h, dx - !!mx * 3, dy - !!my * 3,
                     s->emu, 192 * sizeof(pixel), ref, ref_stride);
        *stride = 192 * sizeof(pixel);
        return &s->emu[192 * !!my * 3 + !!mx * 3];
    }
    *stride = ref_stride;
    return ref + PXSTRIDE(ref_stride) * dy + dx;
}

/* the scaled branch of mc() (recon_tmpl.c:1022-1064) from the position the recorder worked out:
 * emu_edge decision, then mc_scaled (dst8) or mct_scaled (dst16) */
static void mc_scaled_src(const OracleFrame *f, Scratch *s, const Dav1dCudaMcScaledSrc *src, int pl, int bw, int bh,
                          pixel *dst8, ptrdiff_t dst_stride, int16_t *dst16, const Dav1dMCDSPContext *mc)
{
    BD_DECL
    const int sh = pl ? f->ss_hor : 0, sv = pl ? f->ss_ver : 0;
    const int w = ((f->ref_w[src->ref] ? f->ref_w[src->ref] : f->w) + sh) >> sh;
    const int h = ((f->ref_h[src->ref] ? f->ref_h[src->ref] : f->h) + sv) >> sv;
    const int pos_x = src->pos_x, pos_y = src->pos_y;
    const int left = pos_x >> 10, top = pos_y >> 10;
    const int right = ((pos_x + (bw - 1) * src->step_x) >> 10) + 1;
    const int bottom = ((pos_y + (bh - 1) * src->step_y) >> 10) + 1;
    const pixel *ref = f->ref[src->ref][pl];
    ptrdiff_t ref_stride = f->ref_stride[src->ref][pl];
    if (left < 3 || top < 3 || right + 4 > w || bottom + 4 > h) {
        mc->emu_edge(right - left + 7, bottom - top + 7, w, h, left - 3, top - 3,
                     s->emu, 320 * sizeof(pixel), ref, ref_stride);
        ref = &s->emu[320 * 3 + 3];
        ref_stride = 320 * sizeof(pixel);
    } else {
        ref += PXSTRIDE(ref_stride) * top + left;
    }
    if (dst8) mc->mc_scaled[src->filter_2d](dst8, dst_stride, ref, ref_stride, bw, bh, pos_x & 0x3ff, pos_y & 0x3ff,
                                            src->step_x, src->step_y BD_ARG);
    else mc->mct_scaled[src->filter_2d](dst16, ref, ref_stride, bw, bh, pos_x & 0x3ff, pos_y & 0x3ff,
                                        src->step_x, src->step_y BD_ARG);
}

static void run_itx(const OracleFrame *f, Scratch *s, const Dav1dInvTxfmDSPContext *itx, pixel *dst,
                    ptrdiff_t stride, int tx, int txtp, int eob, uint32_t coef_off, int cw4, int ch4)
{
    BD_DECL
    const int w = tx_w4[tx] * 4, h = tx_h4[tx] * 4;
    const int sw = imin(w, 32), sh = imin(h, 32);
    const coef *src = (const coef *) f->cf + coef_off;
    if (!cw4 || !ch4) {
        memcpy(s->cf, src, sw * sh * sizeof(coef));   /* the call zeroes its input */
    } else {
        /* packed descriptor format (include/dav1d_cuda.h): 4*cw4 columns x 4*ch4 rows,
         * column-major with stride 4*ch4 -> the dense sw x sh block the reference takes */
        const int cw = cw4 * 4, ch = ch4 * 4;
        memset(s->cf, 0, sw * sh * sizeof(coef));
        for (int x = 0; x < cw; x++)
            memcpy(s->cf + x * sh, src + x * ch, ch * sizeof(coef));
    }
    itx->itxfm_add[tx][txtp](dst, stride, s->cf, eob BD_ARG);
}

void bitfn(run_frame)(const OracleFrame *const f);
void bitfn(run_frame)(const OracleFrame *const f) {
    Dav1dMCDSPContext mc;
    Dav1dInvTxfmDSPContext itx;
    Dav1dIntraPredDSPContext ip;
    memset(&itx, 0, sizeof(itx));
    bitfn(dav1d_mc_dsp_init)(&mc);
    bitfn(dav1d_itx_dsp_init)(&itx, bitdepth_from_max(f->bitdepth_max));
    bitfn(dav1d_intra_pred_dsp_init)(&ip);
    BD_DECL
    Scratch *const s = aligned_alloc(64, (sizeof(Scratch) + 63) & ~(size_t) 63);
    pixel *const edge = s->edge_buf + 160;

    for (int i = 0; i < f->n_order; i++) {
        const uint32_t code = f->order[i];
        const int cls = code >> 28, idx = code & 0x0fffffff;
        if (cls == 0 || cls == 1) {
            const Dav1dCudaMcDesc *const d = cls ? &f->mc_comp[idx] : &f->mc_put[idx];
            const int pl = d->plane;
            const ptrdiff_t dstride = f->dst_stride[pl];
            pixel *const dst = (pixel *) f->dst[pl] + PXSTRIDE(dstride) * d->y + d->x;
            ptrdiff_t rs;
            if (d->kind == DAV1D_CUDA_MC_PUT) {
                const pixel *src = mc_src(f, s, &d->src[0], pl, d->w, d->h, &rs, &mc);
                mc.mc[d->src[0].filter_2d](dst, dstride, src, rs, d->w, d->h, d->src[0].mx, d->src[0].my BD_ARG);
                continue;
            }
            for (int k = 0; k < 2; k++) {
                const pixel *src = mc_src(f, s, &d->src[k], pl, d->w, d->h, &rs, &mc);
                mc.mct[d->src[k].filter_2d](s->tmp[k], src, rs, d->w, d->h, d->src[k].mx, d->src[k].my BD_ARG);
            }
            switch (d->kind) {
            case DAV1D_CUDA_MC_AVG:
                mc.avg(dst, dstride, s->tmp[0], s->tmp[1], d->w, d->h BD_ARG);
                break;
            case DAV1D_CUDA_MC_W_AVG:
                mc.w_avg(dst, dstride, s->tmp[0], s->tmp[1], d->w, d->h, d->weight BD_ARG);
                break;
            case DAV1D_CUDA_MC_MASK:
                mc.mask(dst, dstride, s->tmp[0], s->tmp[1], d->w, d->h, f->masks + d->aux_off BD_ARG);
                break;
            default:
                mc.w_mask[d->mask_ss](dst, dstride, s->tmp[0], s->tmp[1], d->w, d->h,
                                      f->masks + d->aux_off, d->weight BD_ARG);
                break;
            }
        } else if (cls == 8) {
            const Dav1dCudaMcScaledDesc *const d = &f->mc_scaled[idx];
            const int pl = d->plane;
            const ptrdiff_t dstride = f->dst_stride[pl];
            pixel *const dst = (pixel *) f->dst[pl] + PXSTRIDE(dstride) * d->y + d->x;
            if (d->kind == DAV1D_CUDA_MC_PUT) {
                mc_scaled_src(f, s, &d->src[0], pl, d->w, d->h, dst, dstride, NULL, &mc);
            } else if (d->kind == DAV1D_CUDA_MC_OBMC_H || d->kind == DAV1D_CUDA_MC_OBMC_V) {
                pixel *const lap = (pixel *) s->tmp[0];
                mc_scaled_src(f, s, &d->src[0], pl, d->w, d->h, lap, d->w * sizeof(pixel), NULL, &mc);
                if (d->kind == DAV1D_CUDA_MC_OBMC_H) mc.blend_h(dst, dstride, lap, d->w, d->aux16);
                else mc.blend_v(dst, dstride, lap, d->w, d->h);
            } else {
                for (int k = 0; k < 2; k++) mc_scaled_src(f, s, &d->src[k], pl, d->w, d->h, NULL, 0, s->tmp[k], &mc);
                switch (d->kind) {
                case DAV1D_CUDA_MC_AVG:
                    mc.avg(dst, dstride, s->tmp[0], s->tmp[1], d->w, d->h BD_ARG);
                    break;
                case DAV1D_CUDA_MC_W_AVG:
                    mc.w_avg(dst, dstride, s->tmp[0], s->tmp[1], d->w, d->h, d->weight BD_ARG);
                    break;
                case DAV1D_CUDA_MC_MASK:
                    mc.mask(dst, dstride, s->tmp[0], s->tmp[1], d->w, d->h, f->masks + d->aux_off BD_ARG);
                    break;
                default:
                    mc.w_mask[d->mask_ss](dst, dstride, s->tmp[0], s->tmp[1], d->w, d->h,
                                          f->masks + d->aux_off, d->weight BD_ARG);
                    break;
                }
            }
        } else if (cls == 6) {
            /* obmc(), recon_tmpl.c:1071-1131: neighbour's prediction into the lap buffer + blend */
            const Dav1dCudaMcDesc *const d = &f->mc_obmc[idx];
            const int pl = d->plane;
            const ptrdiff_t dstride = f->dst_stride[pl];
            pixel *const dst = (pixel *) f->dst[pl] + PXSTRIDE(dstride) * d->y + d->x;
            pixel *const lap = (pixel *) s->tmp[0];
            ptrdiff_t rs;
            const pixel *src = mc_src(f, s, &d->src[0], pl, d->w, d->h, &rs, &mc);
            mc.mc[d->src[0].filter_2d](lap, d->w * sizeof(pixel), src, rs, d->w, d->h,
                                       d->src[0].mx, d->src[0].my BD_ARG);
            if (d->kind == DAV1D_CUDA_MC_OBMC_H) mc.blend_h(dst, dstride, lap, d->w, d->aux16);
            else mc.blend_v(dst, dstride, lap, d->w, d->h);
        } else if (cls == 2) {
            const Dav1dCudaWarpDesc *const d = &f->warp[idx];
            const int pl = d->plane;
            int w, h;
            plane_dims(f, pl, &w, &h);
            const ptrdiff_t dstride = f->dst_stride[pl];
            pixel *const dst = (pixel *) f->dst[pl] + PXSTRIDE(dstride) * d->y + d->x;
            const pixel *ref = f->ref[d->ref][pl];
            ptrdiff_t rs = f->ref_stride[d->ref][pl];
            const int dx = d->sx, dy = d->sy;
            if (dx < 3 || dx + 8 + 4 > w || dy < 3 || dy + 8 + 4 > h) {   /* recon_tmpl.c:1171-1182 */
                mc.emu_edge(15, 15, w, h, dx - 3, dy - 3, s->emu, 32 * sizeof(pixel), ref, rs);
                ref = &s->emu[32 * 3 + 3];
                rs = 32 * sizeof(pixel);
            } else {
                ref += PXSTRIDE(rs) * dy + dx;
            }
            mc.warp8x8(dst, dstride, ref, rs, d->abcd, d->mx, d->my BD_ARG);
        } else if (cls == 3) {
            const Dav1dCudaItxDesc *const d = &f->itx[idx];
            const ptrdiff_t dstride = f->dst_stride[d->plane];
            pixel *const dst = (pixel *) f->dst[d->plane] + PXSTRIDE(dstride) * d->y + d->x;
            run_itx(f, s, &itx, dst, dstride, d->tx, d->txtp, d->eob, d->coef_off, d->cw4, d->ch4);
        } else {
            const Dav1dCudaIntraDesc *const d = &f->intra[idx];
            const int pl = d->plane;
            const int ss_hor = pl ? f->ss_hor : 0, ss_ver = pl ? f->ss_ver : 0;
            const ptrdiff_t dstride = f->dst_stride[pl];
            pixel *const dst = (pixel *) f->dst[pl] + PXSTRIDE(dstride) * d->y4 * 4 + d->x4 * 4;
            const int w = d->tw4 * 4, h = d->th4 * 4;
            const int have_left = d->x4 > d->tile_x4_start, have_top = d->y4 > d->tile_y4_start;
            if (d->mode == DAV1D_CUDA_INTRA_PAL) {
                ip.pal_pred(dst, dstride, (const pixel *) f->pal + d->aux, f->pal_idx + d->coef_off, w, h);
            } else if (d->mode == DAV1D_CUDA_INTRA_CFL) {
                const ptrdiff_t ls = f->dst_stride[0];
                const pixel *const y_src = (const pixel *) f->dst[0] +
                    PXSTRIDE(ls) * ((d->y4 * 4) << ss_ver) + ((d->x4 * 4) << ss_hor);
                const int layout_idx = ss_hor ? (ss_ver ? 0 : 1) : 2;
                ip.cfl_ac[layout_idx](s->ac, y_src, ls, d->aux & 0xff, (d->aux >> 8) & 0xff, w, h);
                int angle = 0;
                const enum IntraPredMode m =
                    bytefn(dav1d_prepare_intra_edges)(d->x4, have_left, d->y4, have_top, d->tile_x4_end,
                                                      d->tile_y4_end, 0, dst, dstride, NULL, DC_PRED, &angle,
                                                      d->tw4, d->th4, 0, edge BD_ARG);
                ip.cfl_pred[m](dst, dstride, edge, w, h, s->ac, d->angle_delta BD_ARG);
            } else if (d->mode == DAV1D_CUDA_INTRA_IBC) {
                /* intrabc: mc() with FILTER_2D_BILINEAR from the current picture, recon_tmpl.c:1624-1637 */
                const int dx = (int16_t) (d->aux & 0xffff), dy = (int16_t) (d->aux >> 16);
                const int mx = d->angle_delta, my = d->flags;
                const int pw = (4 * f->bw4) >> ss_hor, ph = (4 * f->bh4) >> ss_ver;
                const pixel *ref = (const pixel *) f->dst[pl];
                ptrdiff_t rs = dstride;
                if (dx < !!mx * 3 || dy < !!my * 3 || dx + w + !!mx * 4 > pw || dy + h + !!my * 4 > ph) {
                    mc.emu_edge(w + !!mx * 7, h + !!my * 7, pw, ph, dx - !!mx * 3, dy - !!my * 3,
                                s->emu, 192 * sizeof(pixel), ref, rs);
                    ref = &s->emu[192 * !!my * 3 + !!mx * 3];
                    rs = 192 * sizeof(pixel);
                } else {
                    ref += PXSTRIDE(rs) * dy + dx;
                }
                mc.mc[FILTER_2D_BILINEAR](dst, dstride, ref, rs, w, h, mx, my BD_ARG);
            } else if (d->mode == DAV1D_CUDA_INTRA_II) {
                /* inter-intra, recon_tmpl.c:1658-1681 */
                int angle = 0;
                pixel *const tmp = (pixel *) s->tmp[0];
                const enum IntraPredMode m =
                    bytefn(dav1d_prepare_intra_edges)(d->x4, have_left, d->y4, have_top, d->tile_x4_end,
                                                      d->tile_y4_end, 0, dst, dstride, NULL, d->angle_delta,
                                                      &angle, d->tw4, d->th4, 0, edge BD_ARG);
                ip.intra_pred[m](tmp, w * sizeof(pixel), edge, w, h, 0, 0, 0 BD_ARG);
                mc.blend(dst, dstride, tmp, w, h, f->pal_idx + d->coef_off);
            } else if (d->mode != DAV1D_CUDA_INTRA_NONE) {
                int angle = d->angle_delta;
                const enum IntraPredMode m =
                    bytefn(dav1d_prepare_intra_edges)(d->x4, have_left, d->y4, have_top, d->tile_x4_end,
                                                      d->tile_y4_end, d->edge_flags, dst, dstride, NULL,
                                                      d->mode, &angle, d->tw4, d->th4, (d->flags >> 10) & 1,
                                                      edge BD_ARG);
                const int max_w = ((4 * f->bw4 + ss_hor) >> ss_hor) - 4 * d->x4;
                const int max_h = ((4 * f->bh4 + ss_ver) >> ss_ver) - 4 * d->y4;
                ip.intra_pred[m](dst, dstride, edge, w, h, angle | d->flags, max_w, max_h BD_ARG);
            }
            if (d->eob >= 0 && d->mode != DAV1D_CUDA_INTRA_PAL)
                run_itx(f, s, &itx, dst, dstride, d->tx, d->txtp, d->eob, d->coef_off, d->cw4, d->ch4);
        }
    }
    free(s);
}

#if BITDEPTH == 16
/* dispatcher + multi-threaded runner live in the 16 bpc object only */
void run_frame_8bpc(const OracleFrame *f);

EXPORT void oracle_ref_frame_run(const OracleFrame *f) {
    if (f->bitdepth_max > 0xff) run_frame_16bpc(f);
    else run_frame_8bpc(f);
}

typedef struct Job { const OracleFrame *frames; int n, next; pthread_mutex_t mu; } Job;

static void *worker(void *arg) {
    Job *const j = arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const int i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->n) return NULL;
        oracle_ref_frame_run(&j->frames[i]);
    }
}

/* Run n independent frames (streams) on up to `threads` host threads. */
EXPORT void oracle_ref_frames_run_mt(const OracleFrame *frames, int n, int threads) {
    Job j = { frames, n, 0, PTHREAD_MUTEX_INITIALIZER };
    if (threads > n) threads = n;
    if (threads <= 1) { worker(&j); return; }
    pthread_t *t = malloc(sizeof(*t) * threads);
    for (int i = 0; i < threads; i++) pthread_create(&t[i], NULL, worker, &j);
    for (int i = 0; i < threads; i++) pthread_join(t[i], NULL);
    free(t);
}
#endif
