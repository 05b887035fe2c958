// The recorder: dav1d's block reconstruction drivers as descriptor emitters (host code only).
//
// dav1d_cuda_record_b_intra() walks an intra block exactly like dav1d_recon_b_intra()
// (src/recon_tmpl.c:1195-1596) and appends one Dav1dCudaIntraDesc where the reference calls
// pal_pred / prepare_intra_edges + intra_pred[m] / cfl_ac + cfl_pred / itxfm_add.  What it decides is
// what the reference's driver decides: which transform blocks exist (block clipped to the frame,
// 64x64 units), the per-transform-block edge flags from the block-level ones, the smooth-neighbour
// and edge-filter bits of the angle argument, the CfL padding, what carries a residual.  It is
// validated through the descriptors it produces: tests/test_recorder.py (same descriptors as the
// generator's, which reproduce the reference driver's pixels bit for bit).
#include <string.h>
#include "ctx.h"
#include "itx_geom.cuh"

using namespace d1;

namespace {

struct Rec {
    Dav1dCudaRecorder *r;
    const Dav1dCudaBlockIntra *b;
    const Dav1dCudaTxCoef *tx;
    int n_tx, next_tx, err;
    int ss_hor, ss_ver;

    // one operation; `res`: the transform block consumes a cbi / cf entry (recon_tmpl.c:1318-1330)
    void emit(int pl, int x4, int y4, int tw4, int th4, int mode, int angle_delta, int flags, int edge_flags, bool res,
              int txsz, uint32_t aux, uint32_t idx_off)
    {
        if (err) return;
        if (r->n_intra >= r->cap_intra) { err = -28; return; }
        Dav1dCudaIntraDesc d;
        memset(&d, 0, sizeof(d));
        const int sh = pl ? ss_hor : 0, sv = pl ? ss_ver : 0;
        d.x4 = (uint16_t)x4; d.y4 = (uint16_t)y4;
        // ts->tiling in this plane's units (the chroma calls shift start and end, :1398-1407)
        d.tile_x4_start = (uint16_t)(r->tile_col_start >> sh); d.tile_y4_start = (uint16_t)(r->tile_row_start >> sv);
        d.tile_x4_end = (uint16_t)(imin(r->tile_col_end, r->bw4) + sh >> sh);
        d.tile_y4_end = (uint16_t)(imin(r->tile_row_end, r->bh4) + sv >> sv);
        d.plane = (uint8_t)pl; d.tw4 = (uint8_t)tw4; d.th4 = (uint8_t)th4;
        d.mode = (uint8_t)mode; d.angle_delta = (int8_t)angle_delta;
        d.flags = (uint16_t)flags; d.edge_flags = (uint8_t)edge_flags;
        d.aux = aux;
        d.eob = -1;
        if (mode == DAV1D_CUDA_INTRA_PAL) {
            d.coef_off = idx_off;
        } else if (res) {
            if (next_tx >= n_tx) { err = -22; return; }
            const Dav1dCudaTxCoef &t = tx[next_tx++];
            d.tx = (uint8_t)txsz; d.txtp = t.txtp; d.eob = t.eob;
            d.coef_off = t.coef_off; d.cw4 = t.cw4; d.ch4 = t.ch4;
        }
        r->intra[r->n_intra++] = d;
    }
};

}  // namespace

extern "C" int dav1d_cuda_record_b_intra(Dav1dCudaRecorder *r, const Dav1dCudaBlockIntra *b,
                                         const Dav1dCudaTxCoef *tx, int n_tx)
{
    if (!r || !b || !r->intra || (n_tx > 0 && !tx) || b->tx >= DAV1D_CUDA_N_RECT_TX_SIZES ||
        b->uvtx >= DAV1D_CUDA_N_RECT_TX_SIZES || r->layout < 0 || r->layout > 3) return -22;
    Rec R;
    R.r = r; R.b = b; R.tx = tx; R.n_tx = n_tx; R.next_tx = 0; R.err = 0;
    const int n0 = r->n_intra;
    const int has_uv = r->layout != 0;
    const int ss_ver = R.ss_ver = r->layout == 1, ss_hor = R.ss_hor = has_uv && r->layout != 3;
    const int bx = b->bx4, by = b->by4, bw4 = b->bw4, bh4 = b->bh4;
    // the block clipped to the frame, and its chroma counterpart (:1207-1211)
    const int w4 = imin(bw4, r->bw4 - bx), h4 = imin(bh4, r->bh4 - by);
    const int cw4 = (w4 + ss_hor) >> ss_hor, ch4 = (h4 + ss_ver) >> ss_ver;
    const int cbw4 = (bw4 + ss_hor) >> ss_hor, cbh4 = (bh4 + ss_ver) >> ss_ver;
    const bool has_chroma = has_uv && (bw4 > ss_hor || (bx & 1)) && (bh4 > ss_ver || (by & 1));
    const TxDim td = tx_dim(b->tx), uvd = tx_dim(b->uvtx);
    const int tw4 = td.w >> 2, th4 = td.h >> 2, utw4 = uvd.w >> 2, uth4 = uvd.h >> 2;
    const int ef = r->intra_edge_filter ? 1024 : 0;                     // ANGLE_USE_EDGE_FILTER_FLAG
    const int y_flags = ((b->sm_flags & 1) ? 512 : 0) | ef;             // intra_flags (:1250-1252)
    const int uv_flags = ((b->sm_flags & 2) ? 512 : 0) | ef;            // sm_uv_fl | intra_edge_filter_flag (:1455,1497)
    const bool res = !b->skip;
    const int TR = 1, BL = 8;                                           // EDGE_I444_TOP_HAS_RIGHT / LEFT_HAS_BOTTOM
    // the block-level flags of the chroma layout: EDGE_I420_* >> (layout - 1) (:1457-1462)
    const int uv_tr_bit = 4 >> (r->layout ? r->layout - 1 : 0), uv_bl_bit = 32 >> (r->layout ? r->layout - 1 : 0);

    for (int init_y = 0; init_y < h4; init_y += 16) {
        const int sub_h4 = imin(h4, 16 + init_y), sub_ch4 = imin(ch4, (init_y + 16) >> ss_ver);
        for (int init_x = 0; init_x < w4; init_x += 16) {
            const int sub_w4 = imin(w4, init_x + 16);
            // ---- luma (:1226-1347)
            if (b->pal_sz[0])
                R.emit(0, bx, by, bw4, bh4, DAV1D_CUDA_INTRA_PAL, 0, 0, 0, false, 0, b->pal_off[0], b->pal_idx_off[0]);
            const bool sb_has_tr = init_x + 16 < w4 ? true : init_y ? false : (b->edge_flags & TR) != 0;
            const bool sb_has_bl = init_x ? false : init_y + 16 < h4 ? true : (b->edge_flags & BL) != 0;
            for (int y = init_y; y < sub_h4; y += th4)
                for (int x = init_x; x < sub_w4; x += tw4) {
                    if (b->pal_sz[0]) {                                  // goto skip_y_pred: residual on the palette pixels
                        if (res) R.emit(0, bx + x, by + y, tw4, th4, DAV1D_CUDA_INTRA_NONE, 0, 0, 0, true, b->tx, 0, 0);
                        continue;
                    }
                    const int eflags = (((y > init_y || !sb_has_tr) && x + tw4 >= sub_w4) ? 0 : TR) |
                                       ((x > init_x || (!sb_has_bl && y + th4 >= sub_h4)) ? 0 : BL);
                    R.emit(0, bx + x, by + y, tw4, th4, b->y_mode, b->y_angle, y_flags, eflags, res, b->tx, 0, 0);
                }
            if (!has_chroma) continue;
            // ---- chroma (:1349-1594).  Planes one after the other (they do not touch each other):
            // palette prediction of the plane, then its transform blocks.
            const int cx = bx >> ss_hor, cy = by >> ss_ver;
            const bool cfl = b->uv_mode == DAV1D_CUDA_CFL_PRED;
            const bool uv_sb_has_tr = ((init_x + 16) >> ss_hor) < cw4 ? true : init_y ? false : (b->edge_flags & uv_tr_bit) != 0;
            const bool uv_sb_has_bl = init_x ? false : ((init_y + 16) >> ss_ver) < ch4 ? true : (b->edge_flags & uv_bl_bit) != 0;
            const int sub_cw4 = imin(cw4, (init_x + 16) >> ss_hor);
            // CfL padding: what of the block lies outside the frame, in chroma 4-px units (:1388-1396)
            const int furthest_r = ((cw4 << ss_hor) + tw4 - 1) & ~(tw4 - 1), furthest_b = ((ch4 << ss_ver) + th4 - 1) & ~(th4 - 1);
            const uint32_t pads = (uint32_t)(cbw4 - (furthest_r >> ss_hor)) | ((uint32_t)(cbh4 - (furthest_b >> ss_ver)) << 8);
            for (int pl = 1; pl <= 2; pl++) {
                if (b->pal_sz[1])
                    R.emit(pl, cx, cy, cbw4, cbh4, DAV1D_CUDA_INTRA_PAL, 0, 0, 0, false, 0, b->pal_off[pl], b->pal_idx_off[1]);
                for (int y = init_y >> ss_ver; y < sub_ch4; y += uth4)
                    for (int x = init_x >> ss_hor; x < sub_cw4; x += utw4) {
                        if (b->pal_sz[1]) {
                            if (res) R.emit(pl, cx + x, cy + y, utw4, uth4, DAV1D_CUDA_INTRA_NONE, 0, 0, 0, true, b->uvtx, 0, 0);
                        } else if (cfl && b->cfl_alpha[pl - 1]) {
                            // cfl_ac + DC edges + cfl_pred of the (single) transform block, then its residual
                            R.emit(pl, cx + x, cy + y, utw4, uth4, DAV1D_CUDA_INTRA_CFL, b->cfl_alpha[pl - 1], 0, 0, res, b->uvtx, pads, 0);
                        } else {
                            const int eflags = (((y > (init_y >> ss_ver) || !uv_sb_has_tr) && x + utw4 >= sub_cw4) ? 0 : TR) |
                                               ((x > (init_x >> ss_hor) || (!uv_sb_has_bl && y + uth4 >= sub_ch4)) ? 0 : BL);
                            R.emit(pl, cx + x, cy + y, utw4, uth4, cfl ? 0 : b->uv_mode, cfl ? 0 : b->uv_angle, uv_flags, eflags,
                                   res, b->uvtx, 0, 0);
                        }
                    }
            }
        }
    }
    if (!R.err && R.next_tx != n_tx) R.err = -22;
    if (R.err) { r->n_intra = n0; return R.err; }
    return r->n_intra - n0;
}
