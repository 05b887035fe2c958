/*
 * oracle/ref_lf.c - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Deblocking checker that runs the reference's OWN loop-filter path on a frame:
 *   dav1d_create_lf_mask_intra / _inter   (src/lf_mask.c:286-401)   per block, decode order
 *   dav1d_calc_eih                        (src/lf_mask.c:403-433)
 *   dav1d_loopfilter_sbrow_cols / _rows   (src/lf_apply_tmpl.c:306-466) per superblock row,
 *   which call dsp->lf.loop_filter_sb[plane][dir] (src/loopfilter_tmpl.c), all compiled where they lie
 *   under /root/reference by oracle/Makefile.
 * The blocks are the generator's Av1Block-style records (real_blocks = 1, one tile); filter levels
 * are drawn per block (what segments / deltas do in a stream).  Besides filtering the picture in place
 * it hands out what dav1d holds when the loop filter starts - f->lf.mask, f->lf.level, f->lf.lim_lut -
 * so that the CUDA path gets the reference's own masks.  Compiled twice (BITDEPTH 8 / 16); this file is
 * this repo's own code.
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
#include "src/loopfilter.h"

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
    struct { int32_t matrix[6]; int16_t abcd[4]; } warp;
} D1SynthBlock;

typedef struct OracleLfFrame {
    void *dst[3];
    ptrdiff_t dst_stride[3];            /* [1] == [2] */
    int32_t w, h, ss_hor, ss_ver, bitdepth_max, no_chroma;
    const D1SynthBlock *blocks;
    int32_t n_blocks;
    uint64_t seed;                      /* per-block filter levels */
    int32_t sharpness;                  /* frame_hdr->loopfilter.sharpness */
    int32_t p_zero_level;               /* per mille of the drawn levels that are 0 */
    int32_t run;                        /* 0: masks / levels only; 1: also filter dst */
    /* out: what the loop filter starts from */
    void *masks;                        /* Av1Filter[sb128w * sb128h] */
    uint8_t *level;                     /* uint8_t[b4_stride * 32 * sb128h][4] */
    uint8_t *lut;                       /* Av1FilterLUT */
    int32_t b4_stride, sb128w, sb128h, w4, h4, sizeof_av1filter;
} OracleLfFrame;

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

static uint64_t next_u64(uint64_t *s) {       /* splitmix64 */
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

/* Sizes the caller allocates the outputs with. */
EXPORT void SUFFIX(oracle_lf_geometry)(OracleLfFrame *const fr) {
    const int bw = ((fr->w + 7) >> 3) << 1, bh = ((fr->h + 7) >> 3) << 1;
    fr->b4_stride = (bw + 31) & ~31;
    fr->sb128w = (bw + 31) >> 5;
    fr->sb128h = (bh + 31) >> 5;
    fr->w4 = (fr->w + 3) >> 2;
    fr->h4 = (fr->h + 3) >> 2;
    fr->sizeof_av1filter = (int) sizeof(Av1Filter);
}

EXPORT int SUFFIX(oracle_lf_frame)(OracleLfFrame *const fr) {
    SUFFIX(oracle_lf_geometry)(fr);
    Dav1dDSPContext dsp;
    memset(&dsp, 0, sizeof(dsp));
    SUFFIX(dav1d_loop_filter_dsp_init)(&dsp.lf);
    Dav1dSequenceHeader seq;
    Dav1dFrameHeader hdr;
    memset(&seq, 0, sizeof(seq));
    memset(&hdr, 0, sizeof(hdr));
    seq.sb128 = 0;
    hdr.loopfilter.level_y[0] = hdr.loopfilter.level_y[1] = 1;
    hdr.loopfilter.level_u = hdr.loopfilter.level_v = !fr->no_chroma;
    hdr.loopfilter.sharpness = fr->sharpness;
    hdr.tiling.cols = hdr.tiling.rows = 1;
    hdr.tiling.col_start_sb[1] = hdr.tiling.row_start_sb[1] = 0x7fff;   /* one tile: no edge fix-ups */

    Dav1dFrameContext *const f = calloc(1, sizeof(*f));
    if (!f) return -12;
    f->seq_hdr = &seq; f->frame_hdr = &hdr; f->dsp = &dsp;
    f->bitdepth_max = fr->bitdepth_max;
    f->cur.data[0] = fr->dst[0]; f->cur.data[1] = fr->dst[1]; f->cur.data[2] = fr->dst[2];
    f->cur.stride[0] = fr->dst_stride[0]; f->cur.stride[1] = fr->dst_stride[1];
    f->cur.p.w = fr->w; f->cur.p.h = fr->h;
    f->cur.p.layout = fr->no_chroma ? DAV1D_PIXEL_LAYOUT_I400 :
                      fr->ss_ver ? DAV1D_PIXEL_LAYOUT_I420 : fr->ss_hor ? DAV1D_PIXEL_LAYOUT_I422 : DAV1D_PIXEL_LAYOUT_I444;
    const int ss_hor = !fr->no_chroma && fr->ss_hor, ss_ver = !fr->no_chroma && fr->ss_ver;
    f->bw = ((fr->w + 7) >> 3) << 1; f->bh = ((fr->h + 7) >> 3) << 1;
    f->w4 = fr->w4; f->h4 = fr->h4;
    f->sb128w = fr->sb128w; f->sb128h = fr->sb128h;
    f->sb_shift = 4; f->sb_step = 16;
    f->sbh = (f->bh + f->sb_step - 1) >> f->sb_shift;
    f->b4_stride = fr->b4_stride;
    Av1Filter *const masks = fr->masks;
    memset(masks, 0, sizeof(Av1Filter) * f->sb128w * f->sb128h);
    memset(fr->level, 0, (size_t) f->b4_stride * 32 * f->sb128h * 4);
    f->lf.mask = masks;
    f->lf.level = (uint8_t (*)[4]) fr->level;
    uint8_t *const right_edge = calloc((size_t) 2 * 32 * (f->sb128h + 1) * 2, 1);
    f->lf.tx_lpf_right_edge[0] = right_edge;
    f->lf.tx_lpf_right_edge[1] = right_edge + 32 * (f->sb128h + 1) * 2;
    dav1d_calc_eih(&f->lf.lim_lut, fr->sharpness);
    memcpy(fr->lut, &f->lf.lim_lut, sizeof(Av1FilterLUT));
    BlockContext *const a = calloc((size_t) f->sb128w + 1, sizeof(*a));
    BlockContext l;
    int ret = 0;
    if (!right_edge || !a) { ret = -12; goto done; }
    for (int x = 0; x <= f->sb128w; x++) {          /* reset_context(), decode.c:2445-2446 */
        memset(a[x].tx_lpf_y, 2, sizeof(a[x].tx_lpf_y));
        memset(a[x].tx_lpf_uv, 1, sizeof(a[x].tx_lpf_uv));
    }

    uint64_t rng = fr->seed * 2 + 1;
    int cur_sbrow = -1;
    for (int i = 0; i < fr->n_blocks; i++) {
        const D1SynthBlock *const s = &fr->blocks[i];
        if (s->tile) { ret = -38; goto done; }      /* one tile only */
        const int sbrow = s->by4 >> f->sb_shift;
        if (sbrow != cur_sbrow) {                    /* decode.c: the left context restarts with every superblock row */
            memset(l.tx_lpf_y, 2, sizeof(l.tx_lpf_y));
            memset(l.tx_lpf_uv, 1, sizeof(l.tx_lpf_uv));
            cur_sbrow = sbrow;
        }
        const int bs = bs_from_dims(s->w4, s->h4);
        if (bs < 0) { ret = -22; goto done; }
        /* this block's filter levels: [dir / plane][ref][mode], only [k][0][0] is read (lf_mask.c:310-311) */
        uint8_t lv[4][8][2];
        memset(lv, 0, sizeof(lv));
        for (int k = 0; k < 4; k++) {
            const uint64_t r = next_u64(&rng);
            lv[k][0][0] = (int) (r % 1000) < fr->p_zero_level ? 0 : (uint8_t) ((r >> 20) & 63);
        }
        Av1Filter *const lflvl = &masks[(s->by4 >> 5) * f->sb128w + (s->bx4 >> 5)];
        BlockContext *const ac = &a[s->bx4 >> 5];
        const int bx4 = s->bx4 & 31, by4 = s->by4 & 31;
        const int cbx4 = bx4 >> ss_hor, cby4 = by4 >> ss_ver;
        uint8_t *const auv = s->has_chroma ? &ac->tx_lpf_uv[cbx4] : NULL;
        uint8_t *const luv = s->has_chroma ? &l.tx_lpf_uv[cby4] : NULL;
        if (s->intra) {                              /* decode.c:1257-1264 */
            dav1d_create_lf_mask_intra(lflvl, f->lf.level, f->b4_stride, (const uint8_t (*)[8][2]) lv,
                                       s->bx4, s->by4, f->w4, f->h4, bs, s->tx, s->uvtx, f->cur.p.layout,
                                       &ac->tx_lpf_y[bx4], &l.tx_lpf_y[by4], auv, luv);
        } else {                                     /* decode.c:1935-1940 */
            const uint16_t tx_split[2] = { s->tx_split ? 1 : 0, 0 };
            dav1d_create_lf_mask_inter(lflvl, f->lf.level, f->b4_stride, (const uint8_t (*)[8][2]) lv,
                                       s->bx4, s->by4, f->w4, f->h4, s->skip, bs, s->max_ytx, tx_split, s->uvtx,
                                       f->cur.p.layout, &ac->tx_lpf_y[bx4], &l.tx_lpf_y[by4], auv, luv);
        }
    }
    if (!fr->run) goto done;
    /* hand-over state is final (one tile: dav1d_loopfilter_sbrow_cols changes no mask); a private copy of the
     * masks is filtered with, so that the caller's copy is exactly what the filter started from */
    {
        Av1Filter *const work = malloc(sizeof(Av1Filter) * f->sb128w * f->sb128h);
        if (!work) { ret = -12; goto done; }
        memcpy(work, masks, sizeof(Av1Filter) * f->sb128w * f->sb128h);
        for (int sby = 0; sby < f->sbh; sby++) {     /* dav1d_filter_sbrow_deblock_cols / _rows, recon_tmpl.c:2038-2071 */
            const int y = sby * f->sb_step * 4;
            pixel *const p[3] = {
                (pixel *) fr->dst[0] + y * PXSTRIDE(f->cur.stride[0]),
                fr->no_chroma ? NULL : (pixel *) fr->dst[1] + (y * PXSTRIDE(f->cur.stride[1]) >> ss_ver),
                fr->no_chroma ? NULL : (pixel *) fr->dst[2] + (y * PXSTRIDE(f->cur.stride[1]) >> ss_ver),
            };
            Av1Filter *const mask = work + (sby >> !seq.sb128) * f->sb128w;
            SUFFIX(dav1d_loopfilter_sbrow_cols)(f, p, mask, sby, 0);
            SUFFIX(dav1d_loopfilter_sbrow_rows)(f, p, mask, sby);
        }
        free(work);
    }
done:
    free(a); free(right_edge); free(f);
    return ret;
}
