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
        if (mode == DAV1D_CUDA_INTRA_PAL || mode == DAV1D_CUDA_INTRA_II) {
            d.coef_off = idx_off;                       // packed palette indices / inter-intra blend mask in the byte pool
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

// ---------------------------------------------------------------- inter blocks
namespace {

struct RecInter {
    Dav1dCudaInterRecorder *r;
    const Dav1dCudaTxCoef *tx;
    int n_tx, next_tx, err, n_emitted;
    int ss_hor, ss_ver;

    template <typename T> T *slot(T *arr, int32_t &n, const int32_t cap) {
        if (err) return nullptr;
        if (!arr || n >= cap) { err = -28; return nullptr; }
        n_emitted++;
        return &arr[n++];
    }

    bool ref_scaled(const int ref) const {
        return (r->ref_w[ref] && r->ref_w[ref] != r->w) || (r->ref_h[ref] && r->ref_h[ref] != r->h);
    }

    // mc() up to the point where it picks the source: position and phase of the block's top-left in the
    // reference plane (recon_tmpl.c:969-977, same-size branch)
    Dav1dCudaMcSrc src_of(const int pl, const int bx, const int by, const int ref, const int mvx, const int mvy,
                          const int filter_2d) const
    {
        const int sh = pl ? ss_hor : 0, sv = pl ? ss_ver : 0;
        Dav1dCudaMcSrc s;
        memset(&s, 0, sizeof(s));
        s.ref = (uint8_t)ref; s.filter_2d = (uint8_t)filter_2d;
        s.mx = (uint8_t)((mvx & (15 >> !sh)) << !sh);
        s.my = (uint8_t)((mvy & (15 >> !sv)) << !sv);
        s.x = bx * (4 >> sh) + (mvx >> (3 + sh));
        s.y = by * (4 >> sv) + (mvy >> (3 + sv));
        return s;
    }
    // the scaled branch (:1013-1021): orig_pos in 1/16 sample -> scale_mv -> 1/1024 position and step; a
    // reference of the frame's own size inside a scaled compound is the same position at step 1024
    Dav1dCudaMcScaledSrc scaled_of(const Dav1dCudaMcSrc &u) const {
        Dav1dCudaMcScaledSrc o;
        memset(&o, 0, sizeof(o));
        o.ref = u.ref; o.filter_2d = u.filter_2d;
        const int orig[2] = { u.x * 16 + u.mx, u.y * 16 + u.my };
        int32_t *const pos[2] = { &o.pos_x, &o.pos_y }, *const step[2] = { &o.step_x, &o.step_y };
        for (int k = 0; k < 2; k++) {
            if (!ref_scaled(u.ref)) { *pos[k] = orig[k] * 64; *step[k] = 1024; continue; }
            const int ref_sz = k ? (r->ref_h[u.ref] ? r->ref_h[u.ref] : r->h) : (r->ref_w[u.ref] ? r->ref_w[u.ref] : r->w);
            const int cur_sz = k ? r->h : r->w;
            const int scale = ((ref_sz << 14) + (cur_sz >> 1)) / cur_sz;           // decode.c:3517
            const long long tmp = (long long)orig[k] * scale + (long long)(scale - 0x4000) * 8;
            const int mag = (int)(((tmp < 0 ? -tmp : tmp) + 128) >> 8);
            *pos[k] = (tmp < 0 ? -mag : mag) + 32;
            *step[k] = (scale + 8) >> 4;
        }
        return o;
    }

    // one prediction: to the same-size list `arr` or, when a reference it reads has another size, to the
    // scaled section `sec`
    void emit_mc(const Dav1dCudaMcDesc &d, Dav1dCudaMcDesc *arr, int32_t &n, const int32_t cap, const int sec) {
        const bool two = d.kind != DAV1D_CUDA_MC_PUT && d.kind != DAV1D_CUDA_MC_OBMC_H && d.kind != DAV1D_CUDA_MC_OBMC_V;
        if (!ref_scaled(d.src[0].ref) && !(two && ref_scaled(d.src[1].ref))) {
            if (Dav1dCudaMcDesc *o = slot(arr, n, cap)) *o = d;
            return;
        }
        Dav1dCudaMcScaledDesc *o = slot(r->scaled[sec], r->n_scaled[sec], r->cap_scaled[sec]);
        if (!o) return;
        memset(o, 0, sizeof(*o));
        o->x = d.x; o->y = d.y; o->w = d.w; o->h = d.h; o->plane = d.plane; o->kind = d.kind;
        o->src[0] = scaled_of(d.src[0]);
        if (two) o->src[1] = scaled_of(d.src[1]);
        o->weight = d.weight; o->mask_ss = d.mask_ss; o->aux16 = d.aux16; o->aux_off = d.aux_off;
    }

    // obmc() (:1071-1132)
    void obmc(const Dav1dCudaBlockInter *b, const int pl, const int w4, const int h4) {
        const int sh = pl ? ss_hor : 0, sv = pl ? ss_ver : 0;
        const int h_mul = 4 >> sh, v_mul = 4 >> sv;
        const int bx = b->bx4, by = b->by4;
        auto ilog2 = [](int v) { int l = 0; while (v > 1) { v >>= 1; l++; } return l; };
        auto lap = [&](const Dav1dCudaNbMv &nb, const int x4, const int y4, const int ow4, const int oh4, const int kind,
                       const int blend_h) {
            Dav1dCudaMcDesc d;
            memset(&d, 0, sizeof(d));
            d.plane = (uint8_t)pl; d.kind = (uint8_t)kind;
            d.x = (uint16_t)(((bx * 4) >> sh) + (x4 - bx) * h_mul);
            d.y = (uint16_t)(((by * 4) >> sv) + (y4 - by) * v_mul);
            d.w = (uint8_t)(ow4 * h_mul); d.h = (uint8_t)(oh4 * v_mul);
            d.aux16 = (uint16_t)blend_h;
            d.src[0] = src_of(pl, x4, y4, nb.ref, nb.mvx, nb.mvy, nb.filter2d);
            const int k = kind == DAV1D_CUDA_MC_OBMC_V;
            emit_mc(d, r->obmc[k], r->n_obmc[k], r->cap_obmc[k], 2 + k);
        };
        if (by > r->tile_row_start && (!pl || b->bw4 * h_mul + b->bh4 * v_mul >= 16)) {
            for (int i = 0, x = 0; x < w4 && i < imin(ilog2(b->bw4), 4);) {
                const Dav1dCudaNbMv &a = r->above[bx + x + 1];             // only odd blocks are considered (:1088)
                const int step4 = imin(imax(a.bw4, 2), 16);
                if (a.ref >= 0) {
                    const int ow4 = imin(step4, b->bw4), oh4 = imin(b->bh4, 16) >> 1;
                    lap(a, bx + x, by, ow4, (oh4 * 3 + 3) >> 2, DAV1D_CUDA_MC_OBMC_H, v_mul * oh4);
                    i++;
                }
                x += step4;
            }
        }
        if (bx > r->tile_col_start) {
            for (int i = 0, y = 0; y < h4 && i < imin(ilog2(b->bh4), 4);) {
                const Dav1dCudaNbMv &l = r->left[by + y + 1];
                const int step4 = imin(imax(l.bh4, 2), 16);
                if (l.ref >= 0) {
                    const int ow4 = imin(b->bw4, 16) >> 1, oh4 = imin(step4, b->bh4);
                    lap(l, bx, by + y, ow4, oh4, DAV1D_CUDA_MC_OBMC_V, 0);
                    i++;
                }
                y += step4;
            }
        }
    }

    // a transform block: consumes one cbi / cf entry, emits itxfm_add when it carries coefficients
    void emit_tx(const int pl, const int x_px, const int y_px, const int txsz) {
        if (err) return;
        if (next_tx >= n_tx) { err = -22; return; }
        const Dav1dCudaTxCoef &t = tx[next_tx++];
        if (t.eob < 0) return;
        Dav1dCudaItxDesc *o = slot(r->itx, r->n_itx, r->cap_itx);
        if (!o) return;
        memset(o, 0, sizeof(*o));
        o->coef_off = t.coef_off; o->x = (uint16_t)x_px; o->y = (uint16_t)y_px; o->eob = t.eob;
        o->plane = (uint8_t)pl; o->tx = (uint8_t)txsz; o->txtp = t.txtp; o->cw4 = t.cw4; o->ch4 = t.ch4;
    }

    // read_coef_tree() (:726-823) in pass 2: the split decisions and the frame-edge rules
    void coef_tree(const int bx, const int by, const int txsz, const int depth, const uint16_t *tx_split, const int x_off,
                   const int y_off)
    {
        static const uint8_t sub_of[19] = { 0, 0, 1, 2, 3, 0, 0, 1, 1, 2, 2, 3, 3, 5, 6, 7, 8, 9, 10 };
        const TxDim td = tx_dim(txsz);
        const int txw = td.w >> 2, txh = td.h >> 2;
        if (depth < 2 && tx_split[depth] && (tx_split[depth] & (1 << (y_off * 4 + x_off)))) {
            const int sub = sub_of[txsz];
            const TxDim sd = tx_dim(sub);
            const int txsw = sd.w >> 2, txsh = sd.h >> 2;
            coef_tree(bx, by, sub, depth + 1, tx_split, x_off * 2, y_off * 2);
            if (txw >= txh && bx + txsw < r->bw4) coef_tree(bx + txsw, by, sub, depth + 1, tx_split, x_off * 2 + 1, y_off * 2);
            if (txh >= txw && by + txsh < r->bh4) {
                coef_tree(bx, by + txsh, sub, depth + 1, tx_split, x_off * 2, y_off * 2 + 1);
                if (txw >= txh && bx + txsw < r->bw4)
                    coef_tree(bx + txsw, by + txsh, sub, depth + 1, tx_split, x_off * 2 + 1, y_off * 2 + 1);
            }
        } else {
            emit_tx(0, bx * 4, by * 4, txsz);
        }
    }
};

void nb_splat(Dav1dCudaInterRecorder *r, const int bx, const int by, const int bw4, const int bh4, const Dav1dCudaNbMv &n) {
    for (int x = bx; x < imin(bx + bw4, r->bw4); x++) r->above[x] = n;
    for (int y = by; y < imin(by + bh4, r->bh4); y++) r->left[y] = n;
    // the 4x4 cells of the block's 8x8 (the sub8x8 chroma of a later block of the same 8x8 reads them)
    if (r->sub8_x != (bx & ~1) || r->sub8_y != (by & ~1)) {
        r->sub8_x = bx & ~1; r->sub8_y = by & ~1;
        Dav1dCudaNbMv none;
        memset(&none, 0, sizeof(none));
        none.ref = -1;
        for (int k = 0; k < 4; k++) r->sub8[k] = none;
    }
    for (int y = by & 1; y < imin((by & 1) + bh4, 2); y++)
        for (int x = bx & 1; x < imin((bx & 1) + bw4, 2); x++) r->sub8[y * 2 + x] = n;
}

}  // namespace

extern "C" int dav1d_cuda_record_nb_intra(Dav1dCudaInterRecorder *r, int bx4, int by4, int bw4, int bh4) {
    if (!r || !r->above || !r->left || bx4 < 0 || by4 < 0 || bx4 >= r->bw4 || by4 >= r->bh4) return -22;
    Dav1dCudaNbMv n;
    memset(&n, 0, sizeof(n));
    n.ref = -1; n.bw4 = (uint8_t)bw4; n.bh4 = (uint8_t)bh4;
    nb_splat(r, bx4, by4, bw4, bh4, n);
    return 0;
}

extern "C" int dav1d_cuda_record_b_inter(Dav1dCudaInterRecorder *r, const Dav1dCudaBlockInter *b,
                                         const Dav1dCudaTxCoef *tx, int n_tx)
{
    if (!r || !b || !r->above || !r->left || (n_tx > 0 && !tx) || r->layout < 0 || r->layout > 3 ||
        b->max_ytx >= DAV1D_CUDA_N_RECT_TX_SIZES || b->uvtx >= DAV1D_CUDA_N_RECT_TX_SIZES ||
        b->bx4 >= r->bw4 || b->by4 >= r->bh4 || b->comp_type > 4) return -22;
    const bool comp = b->comp_type != 0;
    for (int i = 0; i < (comp ? 2 : 1); i++)
        if (b->ref[i] < 0 || b->ref[i] > 6) return -22;
    const int has_uv = r->layout != 0;
    const int ss_ver = r->layout == 1, ss_hor = has_uv && r->layout != 3;
    const int bx = b->bx4, by = b->by4, bw4 = b->bw4, bh4 = b->bh4;
    const bool has_chroma = has_uv && (bw4 > ss_hor || (bx & 1)) && (bh4 > ss_ver || (by & 1));
    // not transcribed: the 4-MV chroma of sub-8x8 blocks
    if (b->motion_mode > 2 || b->warp > 1) return -22;
    const bool warp = b->warp != 0;
    if (warp && (comp || b->motion_mode == 1 || !r->warp)) return -22;
    const bool ii = b->interintra_type != 0, wedge = b->comp_type == 4;
    // key / intra-only frames: the block is an intrabc block (:1624-1637) - integer-pel vector into the picture
    // being decoded, bilinear filter, no compound / OBMC / warp / inter-intra (decode.c:1262-1330)
    const bool ibc = r->intrabc != 0;
    if (ibc && (comp || ii || warp || b->motion_mode || ((b->mvx[0] | b->mvy[0]) & 7) || !r->intra || !r->intra->intra)) return -22;
    if (b->interintra_type > 2 || b->interintra_mode > 3 || (ii && (comp || b->motion_mode))) return -22;
    if ((ii || wedge) && (bw4 < 2 || bh4 < 2 || bw4 > 8 || bh4 > 8)) return -22;          // BS_8x8 .. BS_32x32 (wedge.h:37)
    if (ii && (!r->intra || !r->intra->intra)) return -22;
    if (wedge && (!r->masks || !b->wedge_mask[0] || (has_chroma && (!b->wedge_mask[1] || !b->wedge_mask[2])))) return -22;
    // chroma of a 4-px-wide / -high block: predicted part by part with the partners' vectors when they are inter
    // blocks (:1685-1751), else as one block with this block's vector from the 8x8's origin (:1764-1769)
    const bool narrow = has_chroma && (bw4 == ss_hor || bh4 == ss_ver);
    if (narrow && !ibc && (comp || warp || ii || b->motion_mode)) return -22;
    if (b->motion_mode == 1 && (comp || (bx & 1) || (by & 1))) return -22;     // obmc(): assert(!(t->bx & 1) && !(t->by & 1))

    RecInter R;
    R.r = r; R.tx = tx; R.n_tx = n_tx; R.next_tx = 0; R.err = 0; R.n_emitted = 0;
    R.ss_hor = ss_hor; R.ss_ver = ss_ver;
    const Dav1dCudaInterRecorder saved = *r;                                   // counters to roll back to
    const int saved_n_intra = r->intra ? r->intra->n_intra : 0;
    const int w4 = imin(bw4, r->bw4 - bx), h4 = imin(bh4, r->bh4 - by);
    const int lay = !has_uv ? 0 : ss_hor ? (ss_ver ? 2 : 1) : 0;               // w_mask[chr_layout_idx]: 444 0, 422 1, 420 2
    uint32_t seg_off = 0;

    for (int pl = 0; pl < (has_chroma && !ibc ? 3 : ibc ? 0 : 1); pl++) {
        const int sh = pl ? ss_hor : 0, sv = pl ? ss_ver : 0;
        Dav1dCudaMcDesc d;
        memset(&d, 0, sizeof(d));
        d.plane = (uint8_t)pl;
        d.x = (uint16_t)((bx * 4) >> sh); d.y = (uint16_t)((by * 4) >> sv);
        d.w = (uint8_t)((bw4 * 4) >> sh); d.h = (uint8_t)((bh4 * 4) >> sv);
        d.src[0] = R.src_of(pl, bx, by, b->ref[0], b->mvx[0], b->mvy[0], b->filter2d);
        if (!comp && warp && imin(pl ? (bw4 + ss_hor) >> ss_hor : bw4, pl ? (bh4 + ss_ver) >> ss_ver : bh4) > 1) {
            // warp_affine() (:1134-1193): one warp8x8 per 8x8 of the plane block, positions and fractions from the model
            if ((d.w | d.h) & 7) { R.err = -22; break; }
            const int32_t *const mat = b->warp_matrix;
            for (int y = 0; y < d.h && !R.err; y += 8) {
                const int src_y = by * 4 + ((y + 4) << sv);
                const int64_t mat3_y = (int64_t)mat[3] * src_y + mat[0];
                const int64_t mat5_y = (int64_t)mat[5] * src_y + mat[1];
                for (int x = 0; x < d.w; x += 8) {
                    const int src_x = bx * 4 + ((x + 4) << sh);
                    const int64_t wx = ((int64_t)mat[2] * src_x + mat3_y) >> sh;
                    const int64_t wy = ((int64_t)mat[4] * src_x + mat5_y) >> sv;
                    Dav1dCudaWarpDesc *o = R.slot(r->warp, r->n_warp, r->cap_warp);
                    if (!o) break;
                    memset(o, 0, sizeof(*o));
                    o->plane = (uint8_t)pl; o->ref = (uint8_t)b->ref[0];
                    o->x = (uint16_t)(d.x + x); o->y = (uint16_t)(d.y + y);
                    o->sx = (int)(wx >> 16) - 4; o->sy = (int)(wy >> 16) - 4;
                    o->mx = (((int)wx & 0xffff) - b->warp_abcd[0] * 4 - b->warp_abcd[1] * 7) & ~0x3f;
                    o->my = (((int)wy & 0xffff) - b->warp_abcd[2] * 4 - b->warp_abcd[3] * 4) & ~0x3f;
                    memcpy(o->abcd, b->warp_abcd, sizeof(o->abcd));
                }
            }
            continue;
        }
        if (!comp && pl && narrow) {
            if (pl == 2) continue;                                             // both chroma planes are emitted at pl == 1
            const Dav1dCudaNbMv &L = r->sub8[(by & 1) * 2], &T = r->sub8[bx & 1], &TL = r->sub8[0];
            bool sub = r->sub8_x == (bx & ~1) && r->sub8_y == (by & ~1);
            if (bw4 == 1) sub = sub && L.ref >= 0;
            if (bh4 == ss_ver) sub = sub && T.ref >= 0;
            if (bw4 == 1 && bh4 == ss_ver) sub = sub && TL.ref >= 0;
            const int cx0 = ((bx & ~ss_hor) * 4) >> ss_hor, cy0 = ((by & ~ss_ver) * 4) >> ss_ver;
            auto part = [&](const int nbx, const int nby, const int ref, const int mvx, const int mvy, const int filter,
                            const int ox, const int oy, const int w, const int h) {
                for (int p2 = 1; p2 <= 2; p2++) {
                    Dav1dCudaMcDesc c;
                    memset(&c, 0, sizeof(c));
                    c.plane = (uint8_t)p2; c.kind = DAV1D_CUDA_MC_PUT;
                    c.x = (uint16_t)(cx0 + ox); c.y = (uint16_t)(cy0 + oy); c.w = (uint8_t)w; c.h = (uint8_t)h;
                    c.src[0] = R.src_of(p2, nbx, nby, ref, mvx, mvy, filter);
                    R.emit_mc(c, r->put, r->n_put, r->cap_put, 0);
                }
            };
            if (sub) {
                const int cw = (bw4 * 4) >> ss_hor, ch = (bh4 * 4) >> ss_ver;
                int h_off = 0, v_off = 0;
                if (bw4 == 1 && bh4 == ss_ver) { part(bx - 1, by - 1, TL.ref, TL.mvx, TL.mvy, TL.filter2d, 0, 0, cw, ch); v_off = 2; h_off = 2; }
                if (bw4 == 1) { part(bx - 1, by, L.ref, L.mvx, L.mvy, L.filter2d, 0, v_off, cw, ch); h_off = 2; }
                if (bh4 == ss_ver) { part(bx, by - 1, T.ref, T.mvx, T.mvy, T.filter2d, h_off, 0, cw, ch); v_off = 2; }
                part(bx, by, b->ref[0], b->mvx[0], b->mvy[0], b->filter2d, h_off, v_off, cw, ch);
            } else {
                part(bx & ~ss_hor, by & ~ss_ver, b->ref[0], b->mvx[0], b->mvy[0], b->filter2d, 0, 0,
                     ((bw4 << (bw4 == ss_hor)) * 4) >> ss_hor, ((bh4 << (bh4 == ss_ver)) * 4) >> ss_ver);
            }
            continue;
        }
        if (!comp) {                                                           // :1638-1657, 1764-1778
            d.kind = DAV1D_CUDA_MC_PUT;
            R.emit_mc(d, r->put, r->n_put, r->cap_put, 0);
            if (b->motion_mode == 1) R.obmc(b, pl, w4, h4);
            continue;
        }
        d.src[1] = R.src_of(pl, bx, by, b->ref[1], b->mvx[1], b->mvy[1], b->filter2d);
        int wave = 0;
        switch (b->comp_type) {                                                // :1842-1868, 1890-1906
        case 2: d.kind = DAV1D_CUDA_MC_AVG; break;
        case 4: {                                                              // COMP_INTER_WEDGE: the plane's mask into the pool
            if (b->mask_sign) { const Dav1dCudaMcSrc t = d.src[0]; d.src[0] = d.src[1]; d.src[1] = t; }
            d.kind = DAV1D_CUDA_MC_MASK;
            const uint32_t n = (uint32_t)d.w * d.h;
            if (r->masks_bytes + n > r->cap_masks) { R.err = -28; break; }
            memcpy(r->masks + r->masks_bytes, b->wedge_mask[pl], n);
            d.aux_off = r->masks_bytes;
            r->masks_bytes += n;
            break;
        }
        case 1:
            d.kind = DAV1D_CUDA_MC_W_AVG;
            d.weight = r->jnt_weights[b->ref[0]][b->ref[1]];
            break;
        default:                                                               // COMP_INTER_SEG
            if (b->mask_sign) { const Dav1dCudaMcSrc t = d.src[0]; d.src[0] = d.src[1]; d.src[1] = t; }   // tmp[mask_sign], tmp[!mask_sign]
            if (pl == 0) {
                d.kind = DAV1D_CUDA_MC_W_MASK;
                seg_off = r->masks_bytes;
                r->masks_bytes += (uint32_t)((d.w >> (lay >= 1)) * (d.h >> (lay == 2)));
                d.mask_ss = (uint8_t)lay; d.weight = b->mask_sign;
            } else {
                d.kind = DAV1D_CUDA_MC_MASK;                                   // the mask the luma call emitted
                wave = 1;
            }
            d.aux_off = seg_off;
            break;
        }
        R.emit_mc(d, r->comp[wave], r->n_comp[wave], r->cap_comp[wave], wave);
    }

    // inter-intra (:1658-1681, 1779-1817): per plane the intra prediction of the whole block (no edge flags, no
    // edge filter) blended onto the inter prediction - an intra-class operation, in decode order with the
    // frame's other intra-class operations; the block's residuals follow as residual-only operations
    Rec RI;
    RI.r = r->intra; RI.b = nullptr; RI.tx = tx; RI.n_tx = n_tx; RI.next_tx = 0; RI.err = 0;
    RI.ss_hor = ss_hor; RI.ss_ver = ss_ver;
    if (ibc) {
        // mc() with &f->sr_cur (:957-1008): position bx * h_mul + (mv >> (3 + ss)), fraction mv & (15 >> !ss) - 0, or
        // half a chroma pixel for an odd luma vector - as ONE intra-class operation per plane: it reads pixels
        // earlier operations of this frame write, so it belongs to the wavefront
        for (int pl = 0; pl < (has_chroma ? 3 : 1); pl++) {
            const int sh = pl ? ss_hor : 0, sv = pl ? ss_ver : 0;
            const int sx = ((bx & ~sh) * 4 >> sh) + (b->mvx[0] >> (3 + sh)), sy = ((by & ~sv) * 4 >> sv) + (b->mvy[0] >> (3 + sv));
            const int mx = b->mvx[0] & (15 >> !sh), my = b->mvy[0] & (15 >> !sv);
            // chroma: bw4 << (bw4 == ss_hor), bh4 << (bh4 == ss_ver) (:1631-1635) - the whole 8x8 for a 4-px-wide / -high block
            RI.emit(pl, bx >> sh, by >> sv, pl ? (bw4 + sh) >> sh : bw4, pl ? (bh4 + sv) >> sv : bh4, DAV1D_CUDA_INTRA_IBC, mx, my, 0,
                    false, 0, (uint32_t)(sx & 0xffff) | ((uint32_t)(sy & 0xffff) << 16), 0);
        }
    }
    if ((ii || ibc) && !R.err) {
        static const uint8_t ii_pred[4] = { 0, 1, 2, 9 };                      // DC_PRED, VERT_PRED, HOR_PRED, SMOOTH_PRED
        for (int pl = 0; pl < (!ii ? 0 : has_chroma ? 3 : 1); pl++) {
            const int sh = pl ? ss_hor : 0, sv = pl ? ss_ver : 0;
            RI.emit(pl, bx >> sh, by >> sv, bw4 >> sh, bh4 >> sv, DAV1D_CUDA_INTRA_II, ii_pred[b->interintra_mode], 0, 0, false,
                    0, 0, b->ii_mask_off[pl]);
        }
        if (!b->skip) {                                                        // the transform tree (one split level) as operations
            const TxDim yd = tx_dim(b->max_ytx), ud = tx_dim(b->uvtx);
            int ytx = b->max_ytx, ytw = yd.w >> 2, yth = yd.h >> 2;
            if (b->tx_split[0] & 1) {
                static const uint8_t sub_of[19] = { 0, 0, 1, 2, 3, 0, 0, 1, 1, 2, 2, 3, 3, 5, 6, 7, 8, 9, 10 };
                ytx = sub_of[ytx];
                const TxDim sd = tx_dim(ytx);
                ytw = sd.w >> 2; yth = sd.h >> 2;
            }
            if (b->tx_split[0] & ~1 || b->tx_split[1]) RI.err = -38;          // deeper / partial splits of such a block
            const int w4c = imin(bw4, r->bw4 - bx), h4c = imin(bh4, r->bh4 - by);
            for (int y = 0; y < h4c; y += yth)
                for (int x = 0; x < w4c; x += ytw) RI.emit(0, bx + x, by + y, ytw, yth, DAV1D_CUDA_INTRA_NONE, 0, 0, 0, true, ytx, 0, 0);
            if (has_chroma) {
                const int utw = ud.w >> 2, uth = ud.h >> 2;
                const int cw4 = (w4c + ss_hor) >> ss_hor, ch4 = (h4c + ss_ver) >> ss_ver;
                for (int pl = 1; pl <= 2; pl++)
                    for (int y = 0; y < ch4; y += uth)
                        for (int x = 0; x < cw4; x += utw)
                            RI.emit(pl, (bx >> ss_hor) + x, (by >> ss_ver) + y, utw, uth, DAV1D_CUDA_INTRA_NONE, 0, 0, 0, true, b->uvtx, 0, 0);
            }
        }
        R.next_tx = RI.next_tx;
        if (RI.err) R.err = RI.err;
        R.n_emitted += r->intra->n_intra - saved_n_intra;
    }
    if (!b->skip && !ii && !ibc) {                                             // :1951-2033
        const TxDim yd = tx_dim(b->max_ytx), ud = tx_dim(b->uvtx);
        const int ytw = yd.w >> 2, yth = yd.h >> 2, utw = ud.w >> 2, uth = ud.h >> 2;
        const int cw4 = (w4 + ss_hor) >> ss_hor, ch4 = (h4 + ss_ver) >> ss_ver;
        for (int init_y = 0; init_y < bh4; init_y += 16)
            for (int init_x = 0; init_x < bw4; init_x += 16) {
                int y_off = !!init_y;
                for (int y = init_y; y < imin(h4, init_y + 16); y += yth, y_off++) {
                    int x_off = !!init_x;
                    for (int x = init_x; x < imin(w4, init_x + 16); x += ytw, x_off++)
                        R.coef_tree(bx + x, by + y, b->max_ytx, 0, b->tx_split, x_off, y_off);
                }
                if (has_chroma)
                    for (int pl = 1; pl <= 2; pl++)
                        for (int y = init_y >> ss_ver; y < imin(ch4, (init_y + 16) >> ss_ver); y += uth)
                            for (int x = init_x >> ss_hor; x < imin(cw4, (init_x + 16) >> ss_hor); x += utw)
                                R.emit_tx(pl, ((bx >> ss_hor) + x) * 4, ((by >> ss_ver) + y) * 4, b->uvtx);
            }
    }
    if (!R.err && R.next_tx != n_tx) R.err = -22;
    if (R.err) {
        *r = saved;
        if (r->intra) r->intra->n_intra = saved_n_intra;
        return R.err;
    }
    // decode.c:815-826 + :808-814: what later blocks' obmc() reads of this one
    Dav1dCudaNbMv n;
    n.mvx = b->mvx[0]; n.mvy = b->mvy[0]; n.ref = b->ref[0]; n.bw4 = (uint8_t)bw4; n.bh4 = (uint8_t)bh4; n.filter2d = b->filter2d;
    nb_splat(r, bx, by, bw4, bh4, n);
    return R.n_emitted;
}

// ---------------------------------------------------------------- compact coefficient stream
extern "C" int dav1d_cuda_pack_coefs(const int32_t *cf32, size_t n, int16_t *out16, Dav1dCudaCoefEsc *esc, int cap_esc) {
    if ((n && (!cf32 || !out16)) || cap_esc < 0 || (cap_esc && !esc) || n > 0xffffffffu) return -22;
    int ne = 0;
    for (size_t i = 0; i < n; i++) {
        const int32_t v = cf32[i];
        if (v > -32768 && v <= 32767) { out16[i] = (int16_t)v; continue; }
        if (ne >= cap_esc) return -28;
        out16[i] = -32768;
        esc[ne].off = (uint32_t)i; esc[ne].value = v;
        ne++;
    }
    return ne;
}
