/*
 * oracle/ref_cdef.c - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CDEF checker that runs the reference's OWN path on a frame: dav1d_filter_sbrow_cdef
 * (src/recon_tmpl.c:2073-2100) per superblock row -> dav1d_cdef_brow (src/cdef_apply_tmpl.c:98-309, with its
 * line / column backups of pre-filter pixels) -> dsp->cdef.dir / .fb[] (src/cdef_tmpl.c), compiled where
 * they lie under /root/reference by oracle/Makefile.  The per-64x64 strength index (Av1Filter.cdef_idx) is
 * drawn at random (-1 = "unset" included), the skip mask (Av1Filter.noskip_mask) is built from the generator's
 * block records the way decode.c:1990-1999 does.  Compiled twice (BITDEPTH 8 / 16); this repo's own code.
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
#include "src/cdef.h"
#include "src/recon.h"

#define EXPORT __attribute__((visibility("default")))

typedef struct D1SynthBlock {          /* == dav1d-mirror_b200/csrc/synth.cpp */
    uint16_t bx4, by4;
    uint8_t  w4, h4;
    uint8_t  intra, has_chroma, skip, tile;
    uint8_t  other[78 + 32];
} D1SynthBlock;

typedef struct OracleCdefFrame {
    void *dst[3];
    ptrdiff_t dst_stride[3];            /* [1] == [2] */
    int32_t w, h, ss_hor, ss_ver, bitdepth_max, no_chroma;
    const D1SynthBlock *blocks;
    int32_t n_blocks;
    uint64_t seed;                      /* cdef_idx per 64x64 */
    int32_t damping;                    /* frame_hdr->cdef.damping: 3..6 */
    uint8_t y_strength[8], uv_strength[8];
    int32_t p_unset;                    /* per mille of the 64x64 areas with cdef_idx = -1 */
    int32_t run;
    void *masks;                        /* out: Av1Filter[sb128w * sb128h] (cdef_idx, noskip_mask filled) */
    int32_t sb128w, sb128h, bw, bh;
} OracleCdefFrame;

#if BITDEPTH == 8
#define SUFFIX(name) name##_8bpc
#else
#define SUFFIX(name) name##_16bpc
#endif

static uint64_t next_u64(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

EXPORT int SUFFIX(oracle_cdef_frame)(OracleCdefFrame *const fr) {
    _Static_assert(sizeof(D1SynthBlock) == 120, "block record");
    Dav1dDSPContext dsp;
    memset(&dsp, 0, sizeof(dsp));
    SUFFIX(dav1d_cdef_dsp_init)(&dsp.cdef);
    Dav1dSequenceHeader seq;
    Dav1dFrameHeader hdr;
    memset(&seq, 0, sizeof(seq));
    memset(&hdr, 0, sizeof(hdr));
    seq.cdef = 1;
    hdr.cdef.damping = fr->damping;
    hdr.cdef.n_bits = 3;
    memcpy(hdr.cdef.y_strength, fr->y_strength, 8);
    memcpy(hdr.cdef.uv_strength, fr->uv_strength, 8);

    Dav1dContext *const c = calloc(1, sizeof(*c));
    Dav1dFrameContext *const f = calloc(1, sizeof(*f));
    Dav1dTaskContext *tc = NULL;
    if (!c || !f || posix_memalign((void **) &tc, 64, sizeof(*tc))) { free(c); free(f); return -12; }
    memset(tc, 0, sizeof(*tc));
    c->n_tc = 1;
    c->inloop_filters = DAV1D_INLOOPFILTER_ALL;
    f->c = c; f->seq_hdr = &seq; f->frame_hdr = &hdr; f->dsp = &dsp;
    f->bitdepth_max = fr->bitdepth_max;
    f->cur.data[0] = fr->dst[0]; f->cur.data[1] = fr->dst[1]; f->cur.data[2] = fr->dst[2];
    f->cur.stride[0] = fr->dst_stride[0]; f->cur.stride[1] = fr->dst_stride[1];
    f->cur.p.w = fr->w; f->cur.p.h = fr->h;
    f->cur.p.bpc = fr->bitdepth_max > 1023 ? 12 : fr->bitdepth_max > 255 ? 10 : 8;
    f->cur.p.layout = fr->no_chroma ? DAV1D_PIXEL_LAYOUT_I400 :
                      fr->ss_ver ? DAV1D_PIXEL_LAYOUT_I420 : fr->ss_hor ? DAV1D_PIXEL_LAYOUT_I422 : DAV1D_PIXEL_LAYOUT_I444;
    f->bw = ((fr->w + 7) >> 3) << 1; f->bh = ((fr->h + 7) >> 3) << 1;
    f->sb128w = (f->bw + 31) >> 5; f->sb128h = (f->bh + 31) >> 5;
    f->sb_shift = 4; f->sb_step = 16;
    f->sbh = (f->bh + f->sb_step - 1) >> f->sb_shift;
    fr->sb128w = f->sb128w; fr->sb128h = f->sb128h; fr->bw = f->bw; fr->bh = f->bh;
    Av1Filter *const masks = fr->masks;
    memset(masks, 0, sizeof(Av1Filter) * f->sb128w * f->sb128h);
    f->lf.mask = masks;
    f->lf.p[0] = fr->dst[0]; f->lf.p[1] = fr->dst[1]; f->lf.p[2] = fr->dst[2];
    tc->f = f;
    int ret = 0;
    pixel *lines[2][3] = { { NULL } };
    for (int i = 0; i < 2; i++)
        for (int pl = 0; pl < 3; pl++) {
            const ptrdiff_t st = f->cur.stride[!!pl] < 0 ? -f->cur.stride[!!pl] : f->cur.stride[!!pl];
            lines[i][pl] = calloc(2 * st + 64, 1);
            if (!lines[i][pl]) ret = -12;
            f->lf.cdef_line[i][pl] = lines[i][pl];
        }
    if (ret) goto done;

    /* cdef_idx per 64x64 (decode.c:982-988 reads it per superblock), noskip_mask per block (decode.c:1990-1999) */
    uint64_t rng = fr->seed * 2 + 1;
    for (int i = 0; i < f->sb128w * f->sb128h; i++)
        for (int k = 0; k < 4; k++) {
            const uint64_t r = next_u64(&rng);
            masks[i].cdef_idx[k] = (int) (r % 1000) < fr->p_unset ? -1 : (int8_t) ((r >> 20) & 7);
        }
    for (int i = 0; i < fr->n_blocks; i++) {
        const D1SynthBlock *const s = &fr->blocks[i];
        if (s->skip) continue;
        Av1Filter *const lf = &masks[(s->by4 >> 5) * f->sb128w + (s->bx4 >> 5)];
        const int bx4 = s->bx4 & 31, by4 = s->by4 & 31, bw4 = s->w4, bh4 = s->h4;
        uint16_t (*noskip_mask)[2] = &lf->noskip_mask[by4 >> 1];
        const unsigned mask = (~0U >> (32 - bw4)) << (bx4 & 15);
        const int bx_idx = (bx4 & 16) >> 4;
        for (int y = 0; y < bh4; y += 2, noskip_mask++) {
            (*noskip_mask)[bx_idx] |= mask;
            if (bw4 == 32) (*noskip_mask)[1] |= mask;
        }
    }
    if (fr->run)
        for (int sby = 0; sby < f->sbh; sby++) SUFFIX(dav1d_filter_sbrow_cdef)(tc, sby);
done:
    for (int i = 0; i < 2; i++)
        for (int pl = 0; pl < 3; pl++) free(lines[i][pl]);
    free(tc); free(f); free(c);
    return ret;
}
