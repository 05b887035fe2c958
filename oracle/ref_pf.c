/*
 * oracle/ref_pf.c - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference's OWN in-loop post-filter chain on a frame: dav1d_filter_sbrow() (src/recon_tmpl.c:2149-2160)
 * per superblock row = deblock columns / rows (+ dav1d_copy_lpf: the deblocked lines CDEF and loop restoration
 * read across stripe edges) -> CDEF -> loop restoration (dav1d_lr_sbrow, src/lr_apply_tmpl.c:165-202 ->
 * lr_sbrow / lr_stripe -> dsp->lr.wiener[] / .sgr[], src/looprestoration_tmpl.c), every stage switchable.
 * Masks and levels as oracle/ref_lf.c, CDEF fields as oracle/ref_cdef.c, restoration units drawn at random into
 * f->lf.lr_mask.  Hands out what dav1d holds when the filters start.  Compiled twice (BITDEPTH 8 / 16); this
 * repo's own code.
 * Super-resolution (sr_w > w): frame_hdr->width[0] != width[1], f->sr_cur a second picture of the upscaled width -
 * dav1d_filter_sbrow then runs dav1d_filter_sbrow_resize between CDEF and loop restoration (recon_tmpl.c:2104-2137,
 * 2155-2158), dav1d_copy_lpf keeps RESIZED deblocked lines (lf_apply_tmpl.c:76-91) and the restoration units are
 * indexed with f->sr_sb128w (lr_apply_tmpl.c:142).
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
#include "src/tables.h"
#include "src/lf_mask.h"
#include "src/lf_apply.h"
#include "src/lr_apply.h"
#include "src/loopfilter.h"
#include "src/looprestoration.h"
#include "src/cdef.h"
#include "src/recon.h"

#define EXPORT __attribute__((visibility("default")))

typedef struct D1SynthBlock {          /* == dav1d-mirror_b200/csrc/synth.cpp */
    uint16_t bx4, by4;
    uint8_t  w4, h4;
    uint8_t  intra, has_chroma, skip, tile;
    uint8_t  edge_tr, edge_bl;
    uint8_t  y_mode, uv_mode;
    int8_t   y_angle, uv_angle;
    uint8_t  tx, uvtx;
    uint8_t  other1[57];
    uint8_t  filter2d, mask_sign, max_ytx, tx_split, jnt_weight;
    uint8_t  other2[8 + 32];
} D1SynthBlock;

typedef struct OraclePfFrame {
    void *dst[3];
    ptrdiff_t dst_stride[3];
    int32_t w, h, ss_hor, ss_ver, bitdepth_max, no_chroma;
    const D1SynthBlock *blocks;
    int32_t n_blocks;
    uint64_t seed;
    int32_t do_deblock, do_cdef, do_lr, run;
    int32_t sharpness, p_zero_level;            /* deblock */
    int32_t damping, p_unset;                   /* cdef */
    uint8_t y_strength[8], uv_strength[8];
    int32_t unit_size_log2[2];                  /* loop restoration: frame_hdr->restoration.unit_size */
    int32_t restore_planes;                     /* LR_RESTORE_Y | _U | _V */
    int32_t p_lr_none;                          /* per mille of the units with DAV1D_RESTORATION_NONE */
    /* out */
    void *masks;                                /* Av1Filter[sb128w * sb128h] */
    uint8_t *level;                             /* uint8_t[b4_stride * 32 * sb128h][4] */
    uint8_t *lut;                               /* Av1FilterLUT */
    void *lr_mask;                              /* Av1Restoration[sb128w * sb128h] */
    int32_t b4_stride, sb128w, sb128h, w4, h4, bw, bh, sizeof_av1filter, sizeof_av1restoration;
    /* super-resolution: sr_w > w switches it on (in); sr_sb128w = f->sr_sb128w (out: lr_mask holds sr_sb128w * sb128h) */
    int32_t sr_w, sr_sb128w;
    int32_t resize_step[2], resize_start[2];    /* f->resize_step / _start (decode.c:3576-3583), in */
    void *sr_dst[3];                            /* f->sr_cur planes (in: allocated by the caller; out: the frame) */
    ptrdiff_t sr_stride[2];
    int32_t sb128;                              /* seq_hdr->sb128 (in): superblock rows of 128 luma rows */
} OraclePfFrame;

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
static uint64_t next_u64(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static int rnd_range(uint64_t *s, const int lo, const int hi) { return lo + (int) (next_u64(s) % (uint64_t) (hi - lo + 1)); }

EXPORT void SUFFIX(oracle_pf_geometry)(OraclePfFrame *const fr) {
    fr->bw = ((fr->w + 7) >> 3) << 1; fr->bh = ((fr->h + 7) >> 3) << 1;
    fr->b4_stride = (fr->bw + 31) & ~31;
    fr->sb128w = (fr->bw + 31) >> 5; fr->sb128h = (fr->bh + 31) >> 5;
    fr->w4 = (fr->w + 3) >> 2; fr->h4 = (fr->h + 3) >> 2;
    fr->sizeof_av1filter = (int) sizeof(Av1Filter);
    fr->sizeof_av1restoration = (int) sizeof(Av1Restoration);
    fr->sr_sb128w = fr->sr_w > fr->w ? (fr->sr_w + 127) >> 7 : fr->sb128w;     /* decode.c:3063 */
}

EXPORT int SUFFIX(oracle_pf_frame)(OraclePfFrame *const fr) {
    _Static_assert(sizeof(D1SynthBlock) == 120, "block record");
    SUFFIX(oracle_pf_geometry)(fr);
    Dav1dDSPContext dsp;
    memset(&dsp, 0, sizeof(dsp));
    const int bpc = fr->bitdepth_max > 1023 ? 12 : fr->bitdepth_max > 255 ? 10 : 8;
    SUFFIX(dav1d_loop_filter_dsp_init)(&dsp.lf);
    SUFFIX(dav1d_cdef_dsp_init)(&dsp.cdef);
    SUFFIX(dav1d_loop_restoration_dsp_init)(&dsp.lr, bpc);
    SUFFIX(dav1d_mc_dsp_init)(&dsp.mc);          /* mc.resize: dav1d_filter_sbrow_resize, backup_lpf */
    Dav1dSequenceHeader seq;
    Dav1dFrameHeader hdr;
    memset(&seq, 0, sizeof(seq));
    memset(&hdr, 0, sizeof(hdr));
    seq.cdef = fr->do_cdef;
    seq.sb128 = !!fr->sb128;
    hdr.loopfilter.level_y[0] = hdr.loopfilter.level_y[1] = fr->do_deblock;
    hdr.loopfilter.level_u = hdr.loopfilter.level_v = fr->do_deblock && !fr->no_chroma;
    hdr.loopfilter.sharpness = fr->sharpness;
    hdr.tiling.cols = hdr.tiling.rows = 1;
    hdr.tiling.col_start_sb[1] = hdr.tiling.row_start_sb[1] = 0x7fff;
    hdr.cdef.damping = fr->damping; hdr.cdef.n_bits = 3;
    memcpy(hdr.cdef.y_strength, fr->y_strength, 8);
    memcpy(hdr.cdef.uv_strength, fr->uv_strength, 8);
    hdr.restoration.unit_size[0] = fr->unit_size_log2[0];
    hdr.restoration.unit_size[1] = fr->unit_size_log2[1];

    Dav1dContext *const c = calloc(1, sizeof(*c));
    Dav1dFrameContext *const f = calloc(1, sizeof(*f));
    Dav1dTaskContext *tc = NULL;
    if (!c || !f || posix_memalign((void **) &tc, 64, sizeof(*tc))) { free(c); free(f); return -12; }
    memset(tc, 0, sizeof(*tc));
    c->n_tc = 1; c->tc = tc;
    c->inloop_filters = DAV1D_INLOOPFILTER_ALL;
    tc->f = f;
    f->c = c; f->seq_hdr = &seq; f->frame_hdr = &hdr; f->dsp = &dsp;
    f->bitdepth_max = fr->bitdepth_max;
    f->cur.data[0] = fr->dst[0]; f->cur.data[1] = fr->dst[1]; f->cur.data[2] = fr->dst[2];
    f->cur.stride[0] = fr->dst_stride[0]; f->cur.stride[1] = fr->dst_stride[1];
    f->cur.p.w = fr->w; f->cur.p.h = fr->h; f->cur.p.bpc = bpc;
    f->cur.p.layout = fr->no_chroma ? DAV1D_PIXEL_LAYOUT_I400 :
                      fr->ss_ver ? DAV1D_PIXEL_LAYOUT_I420 : fr->ss_hor ? DAV1D_PIXEL_LAYOUT_I422 : DAV1D_PIXEL_LAYOUT_I444;
    f->sr_cur.p = f->cur;                         /* no super-resolution: the same picture */
    const int superres = fr->sr_w > fr->w;
    hdr.width[0] = fr->w; hdr.width[1] = superres ? fr->sr_w : fr->w;
    hdr.super_res.enabled = superres;
    if (superres) {                               /* decode.c:3566-3584 */
        if (!fr->sr_dst[0] || fr->sb128w * 128 < fr->w) { free(c); free(f); free(tc); return -22; }
        f->sr_cur.p.data[0] = fr->sr_dst[0]; f->sr_cur.p.data[1] = fr->sr_dst[1]; f->sr_cur.p.data[2] = fr->sr_dst[2];
        f->sr_cur.p.stride[0] = fr->sr_stride[0]; f->sr_cur.p.stride[1] = fr->sr_stride[1];
        f->sr_cur.p.p.w = fr->sr_w;
        f->resize_step[0] = fr->resize_step[0]; f->resize_step[1] = fr->resize_step[1];
        f->resize_start[0] = fr->resize_start[0]; f->resize_start[1] = fr->resize_start[1];
    }
    const int ss_hor = !fr->no_chroma && fr->ss_hor, ss_ver = !fr->no_chroma && fr->ss_ver;
    f->bw = fr->bw; f->bh = fr->bh; f->w4 = fr->w4; f->h4 = fr->h4;
    f->sb128w = fr->sb128w; f->sr_sb128w = fr->sr_sb128w; f->sb128h = fr->sb128h;
    f->sb_shift = 4 + seq.sb128; f->sb_step = 16 << seq.sb128;        /* decode.c:3413-3414 */
    f->sbh = (f->bh + f->sb_step - 1) >> f->sb_shift;
    f->b4_stride = fr->b4_stride;
    const int n128 = f->sb128w * f->sb128h, n128_sr = f->sr_sb128w * f->sb128h;
    Av1Filter *const masks = fr->masks;
    Av1Restoration *const lrm = fr->lr_mask;
    memset(masks, 0, sizeof(Av1Filter) * n128);
    memset(lrm, 0, sizeof(Av1Restoration) * n128_sr);
    memset(fr->level, 0, (size_t) f->b4_stride * 32 * f->sb128h * 4);
    f->lf.mask = masks; f->lf.lr_mask = lrm;
    f->lf.level = (uint8_t (*)[4]) fr->level;
    f->lf.restore_planes = fr->do_lr ? fr->restore_planes & (fr->no_chroma ? 1 : 7) : 0;
    for (int pl = 0; pl < 3; pl++) {
        f->lf.p[pl] = fr->dst[pl];
        f->lf.sr_p[pl] = superres ? fr->sr_dst[pl] : fr->dst[pl];
    }
    dav1d_calc_eih(&f->lf.lim_lut, fr->sharpness);
    memcpy(fr->lut, &f->lf.lim_lut, sizeof(Av1FilterLUT));
    int ret = 0;
    uint8_t *const right_edge = calloc((size_t) 2 * 32 * (f->sb128h + 1) * 2, 1);
    uint8_t *const tile_rows = calloc((size_t) f->sbh + 1, 1);
    BlockContext *const a = calloc((size_t) f->sb128w + 1, sizeof(*a));
    pixel *bufs[9] = { NULL };
    if (!right_edge || !tile_rows || !a) { ret = -12; goto done; }
    f->lf.tx_lpf_right_edge[0] = right_edge;
    f->lf.tx_lpf_right_edge[1] = right_edge + 32 * (f->sb128h + 1) * 2;
    f->lf.start_of_tile_row = tile_rows;
    f->a = a;
    for (int i = 0; i < 9; i++) {                 /* cdef_line[2][3]: 2 rows; lr_lpf_line[3]: 12 rows (n_tc == 1) */
        const int pl = i % 3;
        const ptrdiff_t st = i < 6 ? f->cur.stride[!!pl] : f->sr_cur.p.stride[!!pl];   /* lf_apply_tmpl.c:118 */
        bufs[i] = calloc((size_t) st * (i < 6 ? 2 : 12) + 256, 1);
        if (!bufs[i]) { ret = -12; goto done; }
        if (i < 6) f->lf.cdef_line[i / 3][pl] = bufs[i];
        else f->lf.lr_lpf_line[pl] = bufs[i];
    }
    for (int x = 0; x <= f->sb128w; x++) {
        memset(a[x].tx_lpf_y, 2, sizeof(a[x].tx_lpf_y));
        memset(a[x].tx_lpf_uv, 1, sizeof(a[x].tx_lpf_uv));
    }

    /* ---- what pass 1 leaves behind: deblock masks + levels, skip mask, cdef index, restoration units */
    uint64_t rng = fr->seed * 2 + 1;
    BlockContext l;
    int cur_sbrow = -1;
    for (int i = 0; i < fr->n_blocks; i++) {
        const D1SynthBlock *const s = &fr->blocks[i];
        if (s->tile) { ret = -38; goto done; }
        const int sbrow = s->by4 >> f->sb_shift;
        if (sbrow != cur_sbrow) {
            memset(l.tx_lpf_y, 2, sizeof(l.tx_lpf_y));
            memset(l.tx_lpf_uv, 1, sizeof(l.tx_lpf_uv));
            cur_sbrow = sbrow;
        }
        const int bs = bs_from_dims(s->w4, s->h4);
        if (bs < 0) { ret = -22; goto done; }
        uint8_t lv[4][8][2];
        memset(lv, 0, sizeof(lv));
        for (int k = 0; k < 4; k++) {
            const uint64_t r = next_u64(&rng);
            lv[k][0][0] = (int) (r % 1000) < fr->p_zero_level ? 0 : (uint8_t) ((r >> 20) & 63);
        }
        Av1Filter *const lflvl = &masks[(s->by4 >> 5) * f->sb128w + (s->bx4 >> 5)];
        BlockContext *const ac = &a[s->bx4 >> 5];
        const int bx4 = s->bx4 & 31, by4 = s->by4 & 31;
        uint8_t *const auv = s->has_chroma ? &ac->tx_lpf_uv[bx4 >> ss_hor] : NULL;
        uint8_t *const luv = s->has_chroma ? &l.tx_lpf_uv[by4 >> ss_ver] : NULL;
        if (s->intra) {
            dav1d_create_lf_mask_intra(lflvl, f->lf.level, f->b4_stride, (const uint8_t (*)[8][2]) lv,
                                       s->bx4, s->by4, f->w4, f->h4, bs, s->tx, s->uvtx, f->cur.p.layout,
                                       &ac->tx_lpf_y[bx4], &l.tx_lpf_y[by4], auv, luv);
        } else {
            const uint16_t tx_split[2] = { s->tx_split ? 1 : 0, 0 };
            dav1d_create_lf_mask_inter(lflvl, f->lf.level, f->b4_stride, (const uint8_t (*)[8][2]) lv,
                                       s->bx4, s->by4, f->w4, f->h4, s->skip, bs, s->max_ytx, tx_split, s->uvtx,
                                       f->cur.p.layout, &ac->tx_lpf_y[bx4], &l.tx_lpf_y[by4], auv, luv);
        }
        if (!s->skip) {                               /* decode.c:1990-1999 */
            uint16_t (*noskip_mask)[2] = &lflvl->noskip_mask[by4 >> 1];
            const unsigned mask = (~0U >> (32 - s->w4)) << (bx4 & 15);
            const int bx_idx = (bx4 & 16) >> 4;
            for (int y = 0; y < s->h4; y += 2, noskip_mask++) {
                (*noskip_mask)[bx_idx] |= mask;
                if (s->w4 == 32) (*noskip_mask)[1] |= mask;
            }
        }
    }
    for (int i = 0; i < n128 || i < n128_sr; i++) {
        for (int k = 0; k < 4 && i < n128; k++) {
            const uint64_t r = next_u64(&rng);
            masks[i].cdef_idx[k] = (int) (r % 1000) < fr->p_unset ? -1 : (int8_t) ((r >> 20) & 7);
        }
        for (int pl = 0; pl < 3 && i < n128_sr; pl++)
            for (int k = 0; k < 4; k++) {
                Av1RestorationUnit *const u = &lrm[i].lr[pl][k];
                const int kind = rnd_range(&rng, 0, 999);
                if (kind < fr->p_lr_none) { u->type = DAV1D_RESTORATION_NONE; continue; }
                if (kind & 1) {                       /* ranges of the bitstream (decode.c read_restoration_info) */
                    u->type = DAV1D_RESTORATION_WIENER;
                    u->filter_v[0] = pl ? 0 : rnd_range(&rng, -5, 10);
                    u->filter_v[1] = rnd_range(&rng, -23, 8);
                    u->filter_v[2] = rnd_range(&rng, -17, 46);
                    u->filter_h[0] = pl ? 0 : rnd_range(&rng, -5, 10);
                    u->filter_h[1] = rnd_range(&rng, -23, 8);
                    u->filter_h[2] = rnd_range(&rng, -17, 46);
                } else {
                    const int idx = rnd_range(&rng, 0, 15);
                    u->type = DAV1D_RESTORATION_SGRPROJ + idx;
                    u->sgr_weights[0] = dav1d_sgr_params[idx][0] ? rnd_range(&rng, -96, 31) : 0;
                    u->sgr_weights[1] = dav1d_sgr_params[idx][1] ? rnd_range(&rng, -32, 95) : 95;
                }
            }
    }
    if (!fr->run) goto done;
    {
        Av1Filter *const keep = malloc(sizeof(Av1Filter) * n128);   /* the caller's copy stays as the filters found it */
        if (!keep) { ret = -12; goto done; }
        memcpy(keep, masks, sizeof(Av1Filter) * n128);
        for (int sby = 0; sby < f->sbh; sby++) SUFFIX(dav1d_filter_sbrow)(f, sby);
        memcpy(masks, keep, sizeof(Av1Filter) * n128);
        free(keep);
    }
done:
    for (int i = 0; i < 9; i++) free(bufs[i]);
    free(a); free(tile_rows); free(right_edge); free(tc); free(f); free(c);
    return ret;
}
