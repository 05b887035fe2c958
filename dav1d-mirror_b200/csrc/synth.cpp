// Synthetic block-descriptor source (host only, no CUDA): stands in for
// dav1d's pass 1 (entropy decode + mode/MV parse, src/decode.c:717-1250) when
// no bitstreams are available.  For one frame it produces, in decode order,
// exactly the DSP calls src/recon_tmpl.c would make for a random partition /
// mode / MV / coefficient assignment, as the descriptor arrays of
// include/dav1d_cuda.h (the recorder's output) plus the decode order the
// sequential oracle replays.
//
// Workload shape follows SURVEY.md 8(d) configs 2-4: random partition tree per
// 64x64 superblock (NONE/H/V/SPLIT/H4/V4, blocks 8x8..64x64 luma), intra vs
// inter per block, 13 intra modes + angle deltas, filter-intra, palette, CfL,
// 10 MC filters, random MVs (+-mv_range px at 1/8 pel, so some windows leave
// the frame), put / avg / w_avg / wedge mask / segmentation w_mask / warp,
// largest transform or one split, transform type uniform over the legal types
// of the size, three eob classes (dc only / low frequency / full).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../../include/dav1d_cuda.h"

extern "C" {

typedef struct D1SynthParams {
    int32_t w, h;               // luma size, multiples of 8
    int32_t ss_hor, ss_ver;     // chroma subsampling (both 1 = 4:2:0, both 0 = 4:4:4, or luma only with no_chroma)
    int32_t bitdepth_max;
    int32_t no_chroma;          // 1 = luma plane only
    uint64_t seed;
    float p_intra;              // fraction of blocks coded intra
    float p_residual;           // fraction of blocks carrying a residual
    float p_tx_split;           // probability of one transform split
    float p_filter_intra, p_palette, p_cfl;
    float p_avg, p_w_avg, p_wedge, p_seg, p_warp;   // of inter blocks; remainder = single-reference put
    int32_t mv_range;           // pixels
    int32_t n_refs;             // 1..7
    int32_t edge_filter;        // sequence-level intra_edge_filter
    int32_t only_tx;            // >= 0: frame tiled with inter blocks of this tx size only (config 2); else -1
    int32_t only_txtp;          // with only_tx: >= 0 fixes the type
    int32_t eob_class;          // -1 random, 0 dc-only, 1 low-frequency, 2 full
    int32_t dense_coefs;        // 1: dense coefficient blocks (reference layout) instead of packed ones
    float p_obmc;               // single-reference blocks (>= 8x8) that get OBMC blends
    float p_ii;                 // single-reference blocks (8x8..32x32) with inter-intra prediction
    float p_ibc;                // intra-coded blocks predicted by intrabc (copy from the current picture)
    int32_t tile_cols, tile_rows;   // uniform tile grid in superblock units (recon_tmpl.c:1283-1287: nothing
                                    // is predicted across a tile edge); 0 or 1 = one tile
    int32_t real_blocks;            // 1: intra blocks exactly as dav1d_recon_b_intra() would reconstruct an
                                    // Av1Block (smooth-neighbour flags from the above / left block contexts,
                                    // chroma of 4-px-wide / -high blocks with the odd partner, no CfL padding
                                    // inside the picture) + a block record per block for the reference driver
    int32_t ref_w[7], ref_h[7];     // luma size of reference i when it differs from the frame's (0 = same size):
                                    // predictions from it go through the scaled branch of mc() (recon_tmpl.c:1010-1065)
    uint64_t mask_tab;              // real_blocks: address of a D1SynthMaskTab with the decoder's wedge / inter-intra mask
                                    // tables (the tests fill it from the reference's dav1d_masks); 0 = wedge and
                                    // inter-intra blocks are not generated in real-block mode
    uint64_t warp_tab;              // real_blocks: address of n_warp_tab D1SynthWarp entries - valid local-warp models
    int32_t n_warp_tab;             // (matrix + the shear parameters the decoder derives from it: the tests make them
                                    // with the reference's dav1d_get_shear_params); 0 = no warped blocks in real-block mode
    float p_sub8x8;                 // real_blocks, 4:2:0: share of the 8x8 partitions split into 8x4 / 4x8 / 4x4 blocks
                                    // (the chroma of such inter blocks is predicted with the vectors of up to three
                                    // neighbours, recon_tmpl.c:1683-1751)
} D1SynthParams;
// Dav1dWarpedMotionParams as recon_b_inter hands it to warp_affine (t->warpmv of an MM_WARP block)
typedef struct D1SynthWarp {
    int32_t matrix[6];
    int16_t abcd[4];                // alpha, beta, gamma, delta
} D1SynthWarp;

// Wedge and inter-intra masks per chroma layout (0 = 4:4:4 / luma, 1 = 4:2:2, 2 = 4:2:0) and block size
// (w4, h4 in {2, 4, 8}: index log2 - 1): byte offsets into `base` (wedge.h WEDGE_MASK / II_MASK).
typedef struct D1SynthMaskTab {
    const uint8_t *base;
    uint32_t wedge[3][3][3][2][16];     // [layout][w][h][sign][wedge_idx]
    uint32_t ii[3][3][3][4];            // [layout][w][h][II_DC / VERT / HOR / SMOOTH]
} D1SynthMaskTab;

// One coded block as the reference's reconstruction driver sees it (the Av1Block fields
// dav1d_recon_b_intra reads, src/levels.h:262-287, plus what decode_b() hands over).
typedef struct D1SynthBlock {
    uint16_t bx4, by4;            // t->bx, t->by
    uint8_t  w4, h4;              // dav1d_block_dimensions[bs]
    uint8_t  intra, has_chroma, skip, tile;
    uint8_t  edge_tr, edge_bl;    // bit 0 luma, bit 1 chroma: block-level EDGE_*_TOP_HAS_RIGHT / LEFT_HAS_BOTTOM
    uint8_t  y_mode, uv_mode;     // IntraPredMode; 13 = FILTER_PRED (luma) / CFL_PRED (chroma)
    int8_t   y_angle, uv_angle;
    uint8_t  tx, uvtx;            // RectTxfmSize
    uint8_t  pal_sz[2];
    int8_t   cfl_alpha[2];
    uint16_t tile_x0, tile_y0, tile_x1, tile_y1;   // ts->tiling in luma 4-px units
    uint32_t pal_off[3];          // palettes in the palette pool (pixels)
    uint32_t pal_idx_off[2];      // packed indices in the index pool (bytes): luma, chroma
    uint32_t first_op, n_ops;     // the block's operations in `intra` (the order the reference consumes cbi / cf in)
    uint8_t  sm_flags, pad[3];    // bit 0 / 1: smooth neighbour of the luma / chroma block (what sm_flag / sm_uv_flag
                                  // return for the contexts above; the reference driver derives it itself)
    // inter blocks (intra = 0): what dav1d_recon_b_inter reads
    int16_t  mvx[2], mvy[2];      // b->mv[i] (1/8 luma pixel)
    uint8_t  ref[2];              // b->ref[i]
    uint8_t  comp_kind;           // enum Dav1dCudaMcKind: PUT / AVG / W_AVG / W_MASK (segmentation mask)
    uint8_t  filter2d, mask_sign, max_ytx, tx_split, jnt_weight;   // tx_split: b->tx_split0 (one level), b->max_ytx
    uint32_t first_tx, n_tx;      // the block's cbi / cf entries in `tx_recs` (consumption order)
    D1SynthWarp warp;             // comp_kind == 255 (b->motion_mode == MM_WARP): t->warpmv
} D1SynthBlock;
// One cbi / cf entry of an inter block: (eob << 5) | txtp and where the coefficients are
typedef struct D1SynthTx {
    uint32_t coef_off;
    int16_t  eob;
    uint8_t  txtp, cw4, ch4, tx, plane, pad;
} D1SynthTx;

typedef struct D1SynthFrame {
    Dav1dCudaMcDesc *mc_put;   int32_t n_mc_put;   uint32_t *mc_put_tiles;  int32_t n_mc_put_tiles;
    Dav1dCudaMcDesc *mc_comp;  int32_t n_mc_comp;  uint32_t *mc_comp_tiles; int32_t n_mc_comp_tiles[2];
    int32_t n_mc_put_small;    int32_t n_mc_comp_small[2];   // leading tiles of blocks <= 8x8 per list / wave
    Dav1dCudaWarpDesc *warp;   int32_t n_warp;
    Dav1dCudaItxDesc *itx;     int32_t n_itx;      int32_t itx_class_count[19];
    Dav1dCudaIntraDesc *intra; int32_t n_intra;    // decode order
    void *cf;                  uint64_t cf_elems;  // int16 (8 bpc) / int32 coefficients
    uint8_t *masks;            uint64_t masks_bytes;
    void *pal;                 uint64_t pal_px;
    uint8_t *pal_idx;          uint64_t pal_idx_bytes;
    uint32_t *order;           int32_t n_order;    // (class << 28) | index, decode order
    int32_t bw4, bh4;
    double algo_bytes;         // algorithmic HBM bytes of the frame (SURVEY 8d accounting)
    double algo_class[5];      // same, split by launch class: put, compound, warp, itx, intra
    double luma_px;            // luma pixels covered
    int64_t n_blocks, n_intra_blocks;
    Dav1dCudaMcDesc *mc_obmc;  int32_t n_mc_obmc;  uint32_t *mc_obmc_tiles; int32_t n_mc_obmc_tiles[2];
    Dav1dCudaItxDesc *intra_itx; int32_t n_intra_itx; int32_t intra_itx_class_count[19];   // the intra residuals as transforms
    double dense_coef_bytes;   // part of algo_bytes that counts DENSE coefficient blocks; the packed stream is cf_elems
    D1SynthBlock *blocks;      int32_t n_block_recs;   // real_blocks: every block in decode order
    D1SynthTx *tx_recs;        int32_t n_tx_recs;      // real_blocks: cbi / cf entries of the inter blocks
    Dav1dCudaMcScaledDesc *mc_scaled; int32_t n_mc_scaled[4];   // sections: wave 0, wave 1, OBMC_H, OBMC_V
} D1SynthFrame;

}  // extern "C"

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) { next(); next(); }
    uint64_t next() { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; return s * 0x2545F4914F6CDD1Dull; }
    uint32_t u32() { return (uint32_t)(next() >> 32); }
    int range(int n) { return (int)(((uint64_t)u32() * (uint64_t)n) >> 32); }      // [0, n)
    int irange(int lo, int hi) { return lo + range(hi - lo + 1); }                 // [lo, hi]
    float unit() { return (u32() >> 8) * (1.0f / 16777216.0f); }
    bool chance(float p) { return unit() < p; }
};

// enum RectTxfmSize -> dims in 4-px units (levels.h:44-78)
const uint8_t TXW4[19] = { 1, 2, 4, 8, 16, 1, 2, 2, 4, 4, 8, 8, 16, 1, 4, 2, 8, 4, 16 };
const uint8_t TXH4[19] = { 1, 2, 4, 8, 16, 2, 1, 4, 2, 8, 4, 16, 8, 4, 1, 8, 2, 16, 4 };

int tx_from_dims(int w4, int h4) {
    for (int t = 0; t < 19; t++)
        if (TXW4[t] == w4 && TXH4[t] == h4) return t;
    return -1;
}

// populated itxfm_add slots per size (itx_tmpl.c:248-268)
int pick_txtp(Rng &r, int tx) {
    const int w4 = TXW4[tx], h4 = TXH4[tx], m = std::max(w4, h4);
    if (m == 16) return 0;
    if (m == 8) return r.chance(0.5f) ? 0 : 9;
    if (w4 == 4 && h4 == 4) { static const int t[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 }; return t[r.range(12)]; }
    if (tx == 0 && r.chance(0.04f)) return 16;   // WHT_WHT
    return r.range(16);
}

struct Gen {
    const D1SynthParams &P;
    Rng rng;
    int bw4, bh4, hbd;
    std::vector<Dav1dCudaMcDesc> put, comp0, comp1, obmc_h, obmc_v;
    std::vector<Dav1dCudaWarpDesc> warp;
    std::vector<Dav1dCudaItxDesc> itx;
    std::vector<Dav1dCudaIntraDesc> intra;
    std::vector<int32_t> cf32;
    std::vector<uint8_t> masks, pal_idx;
    std::vector<uint16_t> pal;
    struct Ord { uint8_t cls; uint32_t idx; };   // cls 1 = comp wave 0, 5 = comp wave 1 (remapped later)
    std::vector<Ord> order;
    std::vector<uint8_t> decoded[3];             // per plane, 4x4 cells
    int pw4[3], ph4[3];
    double algo = 0, luma_px = 0, dense_coef_bytes = 0;
    double algo_cls[5] = { 0, 0, 0, 0, 0 };
    int cur_cls = 0;
    void add_bytes(int cls, double b) { algo += b; algo_cls[cls] += b; }
    int64_t n_blocks = 0, n_intra_blocks = 0;
    int Bp, Bc;

    explicit Gen(const D1SynthParams &p) : P(p), rng(p.seed) {
        bw4 = p.w / 4; bh4 = p.h / 4; hbd = p.bitdepth_max > 0xff;
        Bp = hbd ? 2 : 1; Bc = hbd ? 4 : 2;
        for (int pl = 0; pl < 3; pl++) {
            const int sh = pl ? p.ss_hor : 0, sv = pl ? p.ss_ver : 0;
            pw4[pl] = (bw4 + sh) >> sh; ph4[pl] = (bh4 + sv) >> sv;
            decoded[pl].assign((size_t)pw4[pl] * ph4[pl], 0);
        }
        ctx_init();
    }
    int nplanes() const { return P.no_chroma ? 1 : 3; }

    // ---- coefficients: the bounding box of the non-zero coefficients (rounded up to 4 x 4)
    // of the column-major sw x sh block is appended to the stream (packed format of
    // Dav1dCudaItxDesc: cw4 columns x ch4 rows of four, stride 4 * ch4); P.dense_coefs keeps
    // the reference's dense sw x sh layout (cw4 = ch4 = 0)
    uint32_t emit_coefs(int tx, int txtp, int16_t *eob_out, uint8_t *cw4_out, uint8_t *ch4_out) {
        const int w = TXW4[tx] * 4, h = TXH4[tx] * 4, sw = std::min(w, 32), sh = std::min(h, 32);
        int32_t c[32 * 32];
        memset(c, 0, sizeof(int32_t) * sw * sh);
        const int cls = P.eob_class >= 0 ? P.eob_class : (rng.chance(0.2f) ? 0 : rng.chance(0.6f) ? 1 : 2);
        // amplitude so that the reconstructed residual spans a good part of the pixel range
        const double amp = (double)P.bitdepth_max * 8.0 * sqrt((double)(w * h)) / 16.0;
        const int cmax = hbd ? (P.bitdepth_max > 1023 ? 524287 : 131071) : 32767;   // recon_tmpl.c:594
        int last = 0;
        if (cls == 0) {
            c[0] = rng.irange(-std::min(cmax, (int)amp), std::min(cmax, (int)amp));
            if (txtp == 16) c[0] = rng.irange(-P.bitdepth_max * 4, P.bitdepth_max * 4);
        } else {
            const int lw = cls == 1 ? std::min(sw, 8) : sw, lh = cls == 1 ? std::min(sh, 8) : sh;
            for (int x = 0; x < lw; x++)
                for (int y = 0; y < lh; y++) {
                    if (cls == 1 && rng.chance(0.5f)) continue;
                    const double decay = 1.0 / (1.0 + 0.75 * (x + y));
                    int a = (int)(amp * decay);
                    if (txtp == 16) a = P.bitdepth_max * 2;
                    a = std::max(1, std::min(a, cmax));
                    const int v = rng.irange(-a, a);
                    c[y + x * sh] = v;
                    if (v) last = y + x * sh;
                }
            if (!last) { c[1 % (sw * sh)] = 1; last = 1; }
        }
        *eob_out = (int16_t)(cls == 0 ? 0 : std::min(last, sw * sh - 1));
        add_bytes(cur_cls, (double)sw * sh * Bc);      // algorithmic bytes: the dense block
        dense_coef_bytes += (double)sw * sh * Bc;
        const uint32_t off = (uint32_t)cf32.size();
        if (P.dense_coefs) {
            *cw4_out = *ch4_out = 0;
            cf32.insert(cf32.end(), c, c + sw * sh);
            return off;
        }
        int nzw = 1, nzh = 1;
        for (int x = 0; x < sw; x++)
            for (int y = 0; y < sh; y++)
                if (c[y + x * sh]) { nzw = std::max(nzw, x + 1); nzh = std::max(nzh, y + 1); }
        const int cw = (nzw + 3) & ~3, ch = (nzh + 3) & ~3;
        *cw4_out = (uint8_t)(cw / 4); *ch4_out = (uint8_t)(ch / 4);
        for (int x = 0; x < cw; x++)
            for (int y = 0; y < ch; y++) cf32.push_back(c[y + x * sh]);
        return off;
    }

    void mark(int pl, int x4, int y4, int w4, int h4) {
        for (int y = y4; y < std::min(y4 + h4, ph4[pl]); y++)
            for (int x = x4; x < std::min(x4 + w4, pw4[pl]); x++) decoded[pl][(size_t)y * pw4[pl] + x] = 1;
    }
    bool all_decoded(int pl, int x0, int y0, int x1, int y1) const {
        if (x0 < 0 || y0 < 0 || x1 > pw4[pl] || y1 > ph4[pl] || x0 >= x1 || y0 >= y1) return false;
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++)
                if (!decoded[pl][(size_t)y * pw4[pl] + x]) return false;
        return true;
    }

    void add_itx(int pl, int x4, int y4, int tx) {
        Dav1dCudaItxDesc d;
        memset(&d, 0, sizeof(d));
        d.plane = (uint8_t)pl; d.x = (uint16_t)(x4 * 4); d.y = (uint16_t)(y4 * 4);
        d.tx = (uint8_t)tx;
        d.txtp = (uint8_t)(P.only_txtp >= 0 ? P.only_txtp : pick_txtp(rng, tx));
        cur_cls = 3;
        d.coef_off = emit_coefs(tx, d.txtp, &d.eob, &d.cw4, &d.ch4);
        add_bytes(3, 2.0 * Bp * TXW4[tx] * TXH4[tx] * 16);
        order.push_back({ 3, (uint32_t)itx.size() });
        itx.push_back(d);
        if (P.real_blocks) {
            D1SynthTx t;
            memset(&t, 0, sizeof(t));
            t.coef_off = d.coef_off; t.eob = d.eob; t.txtp = d.txtp; t.cw4 = d.cw4; t.ch4 = d.ch4;
            t.tx = (uint8_t)tx; t.plane = (uint8_t)pl;
            tx_recs.push_back(t);
        }
    }

    // one tx-sized intra-class op
    void add_intra(int pl, int x4, int y4, int tw4, int th4, int mode, int angle_delta, int flags,
                   bool residual, uint32_t aux = 0, uint32_t idx_off = 0)
    {
        Dav1dCudaIntraDesc d;
        memset(&d, 0, sizeof(d));
        d.x4 = (uint16_t)x4; d.y4 = (uint16_t)y4;
        {   // the current tile in this plane's 4-px units (ts->tiling.col_start .. row_end >> ss)
            const int sh = pl ? P.ss_hor : 0, sv = pl ? P.ss_ver : 0;
            d.tile_x4_start = (uint16_t)(tile_x0 >> sh); d.tile_y4_start = (uint16_t)(tile_y0 >> sv);
            d.tile_x4_end = (uint16_t)std::min((tile_x1 + sh) >> sh, pw4[pl]);
            d.tile_y4_end = (uint16_t)std::min((tile_y1 + sv) >> sv, ph4[pl]);
        }
        d.plane = (uint8_t)pl; d.tw4 = (uint8_t)tw4; d.th4 = (uint8_t)th4;
        d.mode = (uint8_t)mode; d.angle_delta = (int8_t)angle_delta;
        d.flags = (uint16_t)flags;
        d.aux = aux;
        d.eob = -1;
        if (mode <= 13) {
            const bool tr = y4 > 0 && all_decoded(pl, x4 + tw4, y4 - 1, std::min(x4 + 2 * tw4, pw4[pl]), y4);
            const bool bl = x4 > 0 && all_decoded(pl, x4 - 1, y4 + th4, x4, std::min(y4 + 2 * th4, ph4[pl]));
            d.edge_flags = (uint8_t)((tr ? 1 : 0) | (bl ? 8 : 0));
        }
        const int w = tw4 * 4, h = th4 * 4;
        cur_cls = 4;
        if (mode == DAV1D_CUDA_INTRA_PAL) {
            d.coef_off = idx_off;
            add_bytes(4, (double)Bp * w * h + 0.5 * w * h);
        } else if (mode == DAV1D_CUDA_INTRA_IBC) {
            add_bytes(4, 2.0 * Bp * w * h);
        } else if (mode == DAV1D_CUDA_INTRA_II) {
            d.coef_off = idx_off;                                   // blend mask in the byte pool
            add_bytes(4, (double)Bp * (2 * w + 2 * h + 1) + 2.0 * Bp * w * h + (double)w * h);
        } else {
            if (mode != DAV1D_CUDA_INTRA_NONE) add_bytes(4, (double)Bp * (2 * w + 2 * h + 1));
            if (mode == DAV1D_CUDA_INTRA_CFL) add_bytes(4, (double)Bp * (w << P.ss_hor) * (h << P.ss_ver));
            if (mode != DAV1D_CUDA_INTRA_NONE || residual) add_bytes(4, (double)Bp * w * h);   // final write
            if (mode == DAV1D_CUDA_INTRA_NONE) add_bytes(4, (double)Bp * w * h);               // read back pal pred
            if (residual) {
                d.tx = (uint8_t)tx_from_dims(tw4, th4);
                d.txtp = (uint8_t)pick_txtp(rng, d.tx);
                d.coef_off = emit_coefs(d.tx, d.txtp, &d.eob, &d.cw4, &d.ch4);
            }
        }
        order.push_back({ 4, (uint32_t)intra.size() });
        intra.push_back(d);
        if (mode != DAV1D_CUDA_INTRA_PAL) mark(pl, x4, y4, tw4, th4);
    }

    // transform shapes are at most 4:1 (4:2:2 chroma of a 1:2 block would be 1:8)
    static void fit_tx(int &tw4, int &th4) {
        while (tw4 > 4 * th4) tw4 >>= 1;
        while (th4 > 4 * tw4) th4 >>= 1;
    }
    static void split_tx(int &tw4, int &th4) {   // halve the longer side (both if square)
        if (tw4 == th4) { if (tw4 > 1) { tw4 >>= 1; th4 >>= 1; } }
        else if (tw4 > th4) tw4 >>= 1;
        else th4 >>= 1;
    }

    // intrabc block (recon_tmpl.c:1624-1637): prediction = bilinear mc() from the already decoded
    // part of the CURRENT picture (any position in the superblock rows above; now and then past the
    // right edge, which the reference pads with emu_edge), residuals as residual-only operations
    bool ibc_block(int bx4, int by4, int w4, int h4) {
        const int sb_top = (by4 & ~15) * 4 - tile_y0 * 4, wpx = w4 * 4, hpx = h4 * 4;   // rows of the tile above this SB row
        // 4-px-wide / -high blocks (real-block mode with p_sub8x8): the block at the odd position of its 8x8 predicts
        // the chroma of the whole 8x8 with its own vector (recon_tmpl.c:1631-1635: bw4 << (bw4 == ss_hor) from t->bx & ~ss_hor)
        const bool narrow = (w4 < 2 || h4 < 2);
        if (narrow && !(P.real_blocks && P.p_sub8x8 > 0.f && P.ss_hor && !P.no_chroma)) return false;
        const int pad = narrow ? 4 : 0;                    // the widened chroma source stays inside the tile / above the SB row
        if (w4 > 16 || h4 > 16 || sb_top < hpx + 2 + 2 * pad) return false;
        const int W = std::min(tile_x1, bw4) * 4 - tile_x0 * 4;                      // the source stays inside the tile
        if (W < wpx + 8 + 2 * pad) return false;
        const bool one_tile = tile_x0 == 0 && tile_x1 >= bw4;
        const int sxl = tile_x0 * 4 + pad + ((!narrow && one_tile && rng.chance(0.1f)) ? W - wpx + 2 * rng.range(4) : 2 * rng.range((W - wpx - 2 * pad) / 2 + 1));
        const int syl = tile_y0 * 4 + pad + 2 * rng.range((sb_top - hpx - 2 - 2 * pad) / 2 + 1);
        const bool half = rng.chance(0.5f);          // odd luma vector components: half-pel chroma
        // ... but never one pixel into the tile to the right, which is decoded later (a stream may not do that)
        const bool halfx = half && (one_tile || sxl + 1 + wpx <= std::min(tile_x1, bw4) * 4);
        n_intra_blocks++;
        // real-block mode: the block record dav1d_recon_b_inter reads for an intrabc block of a key / intra-only
        // frame (b->intra == 0, integer-pel b->mv[0], FILTER_2D_BILINEAR; decode.c:1262-1330)
        D1SynthBlock rec;
        memset(&rec, 0, sizeof(rec));
        rec.bx4 = (uint16_t)bx4; rec.by4 = (uint16_t)by4; rec.w4 = (uint8_t)w4; rec.h4 = (uint8_t)h4;
        const bool hc = !P.no_chroma && (w4 > P.ss_hor || (bx4 & 1)) && (h4 > P.ss_ver || (by4 & 1));
        rec.intra = 0; rec.has_chroma = hc ? 1 : 0; rec.skip = 1; rec.tile = (uint8_t)tile_no;
        rec.tile_x0 = (uint16_t)tile_x0; rec.tile_y0 = (uint16_t)tile_y0;
        rec.tile_x1 = (uint16_t)std::min(tile_x1, bw4); rec.tile_y1 = (uint16_t)std::min(tile_y1, bh4);
        rec.comp_kind = 254; rec.filter2d = 9;
        rec.mvx[0] = (int16_t)((sxl + (halfx ? 1 : 0) - bx4 * 4) * 8);
        rec.mvy[0] = (int16_t)((syl + (half ? 1 : 0) - by4 * 4) * 8);
        rec.first_tx = (uint32_t)tx_recs.size();
        {
            int a = std::min(w4, 16), b2 = std::min(h4, 16);
            rec.max_ytx = (uint8_t)tx_from_dims(a, b2);
            int ua = std::min(std::max(1, w4 >> P.ss_hor), 8), ub = std::min(std::max(1, h4 >> P.ss_ver), 8);
            fit_tx(ua, ub);
            rec.uvtx = (uint8_t)tx_from_dims(ua, ub);
        }
        // chroma block of the (odd) block: the widened one of a narrow block
        const int cbw4 = (w4 + P.ss_hor) >> P.ss_hor, cbh4 = (h4 + P.ss_ver) >> P.ss_ver;
        if (narrow) rec.uvtx = (uint8_t)tx_from_dims(cbw4, cbh4);
        auto rec_tx = [&]() {                             // the cbi / cf entry of the residual-only operation just added
            if (!P.real_blocks) return;
            const Dav1dCudaIntraDesc &o = intra.back();
            D1SynthTx t;
            memset(&t, 0, sizeof(t));
            t.coef_off = o.coef_off; t.eob = o.eob; t.txtp = o.txtp; t.cw4 = o.cw4; t.ch4 = o.ch4; t.tx = o.tx; t.plane = o.plane;
            tx_recs.push_back(t);
        };
        for (int pl = 0; pl < (narrow ? (hc ? 3 : 1) : nplanes()); pl++) {
            const int sh = pl ? P.ss_hor : 0, sv = pl ? P.ss_ver : 0;
            // chroma of a narrow block: from the 8x8's origin (t->bx & ~ss_hor, t->by & ~ss_ver)
            const int lx = sxl + (halfx ? 1 : 0) - (narrow && pl ? 4 * (bx4 & sh) : 0);
            const int ly = syl + (half ? 1 : 0) - (narrow && pl ? 4 * (by4 & sv) : 0);
            const int sx = lx >> sh, sy = ly >> sv;
            const int mx = (sh && (lx & 1)) ? 8 : 0, my = (sv && (ly & 1)) ? 8 : 0;
            add_intra(pl, bx4 >> sh, by4 >> sv, narrow && pl ? cbw4 : w4 >> sh, narrow && pl ? cbh4 : h4 >> sv,
                      DAV1D_CUDA_INTRA_IBC, mx, my, false, (uint32_t)(sx & 0xffff) | ((uint32_t)(sy & 0xffff) << 16));
        }
        if (rng.chance(P.p_residual)) {
            int tw4 = std::min(w4, 16), th4 = std::min(h4, 16);
            rec.skip = 0;
            if (rng.chance(P.p_tx_split) && !(narrow && tw4 == th4)) { split_tx(tw4, th4); rec.tx_split = tw4 * th4 < std::min(w4, 16) * std::min(h4, 16); }
            for (int y = 0; y < h4; y += th4)
                for (int x = 0; x < w4; x += tw4) {
                    add_intra(0, bx4 + x, by4 + y, tw4, th4, DAV1D_CUDA_INTRA_NONE, 0, 0, true);
                    rec_tx();
                }
            if (narrow) {
                if (hc)
                    for (int pl = 1; pl <= 2; pl++) {
                        add_intra(pl, bx4 >> P.ss_hor, by4 >> P.ss_ver, cbw4, cbh4, DAV1D_CUDA_INTRA_NONE, 0, 0, true);
                        rec_tx();
                    }
            } else if (!P.no_chroma) {
                const int cw4 = w4 >> P.ss_hor, ch4 = h4 >> P.ss_ver;
                int utw4 = std::min(cw4, 8), uth4 = std::min(ch4, 8);
                fit_tx(utw4, uth4);
                for (int pl = 1; pl <= 2; pl++)
                    for (int y = 0; y < ch4; y += uth4)
                        for (int x = 0; x < cw4; x += utw4) {
                            add_intra(pl, (bx4 >> P.ss_hor) + x, (by4 >> P.ss_ver) + y, utw4, uth4,
                                      DAV1D_CUDA_INTRA_NONE, 0, 0, true);
                            rec_tx();
                        }
            }
        }
        if (P.real_blocks) {
            rec.n_tx = (uint32_t)tx_recs.size() - rec.first_tx;
            blocks.push_back(rec);
        }
        return true;
    }

    void intra_block(int bx4, int by4, int w4, int h4) {
        if (P.p_ibc > 0.f && rng.chance(P.p_ibc) && ibc_block(bx4, by4, w4, h4)) {
            if (P.real_blocks) ctx_set(bx4, by4, w4, h4, 0, 0, true, 0);      // intrabc is coded as an inter block
            return;
        }
        n_intra_blocks++;
        const bool real = P.real_blocks != 0;
        const int sh = P.ss_hor, sv = P.ss_ver;
        D1SynthBlock rec;
        memset(&rec, 0, sizeof(rec));
        rec.bx4 = (uint16_t)bx4; rec.by4 = (uint16_t)by4; rec.w4 = (uint8_t)w4; rec.h4 = (uint8_t)h4;
        rec.intra = 1; rec.tile = (uint8_t)tile_no;
        rec.tile_x0 = (uint16_t)tile_x0; rec.tile_y0 = (uint16_t)tile_y0;
        rec.tile_x1 = (uint16_t)std::min(tile_x1, bw4); rec.tile_y1 = (uint16_t)std::min(tile_y1, bh4);
        rec.first_op = (uint32_t)intra.size();
        // block-level edge availability (what the partition tree of src/intra_edge.c tells decode_b):
        // here simply whether the area is decoded
        rec.edge_tr = (by4 > 0 && all_decoded(0, bx4 + w4, by4 - 1, std::min(bx4 + 2 * w4, pw4[0]), by4)) ? 1 : 0;
        rec.edge_bl = (bx4 > 0 && all_decoded(0, bx4 - 1, by4 + h4, bx4, std::min(by4 + 2 * h4, ph4[0]))) ? 1 : 0;
        const bool residual = rng.chance(P.p_residual);
        rec.skip = residual ? 0 : 1;
        // smooth-neighbour flag: sm_flag(t->a, bx4) | sm_flag(&t->l, by4) (ipred_prepare.h:95-101)
        const bool sm = real ? ((a_intra[bx4] && smooth_mode(a_mode[bx4])) || (l_intra[by4] && smooth_mode(l_mode[by4])))
                             : rng.chance(0.25f);
        const int flags = (P.edge_filter ? 1024 : 0) | (sm ? 512 : 0);
        rec.sm_flags = sm ? 1 : 0;
        int tw4 = std::min(w4, 16), th4 = std::min(h4, 16);
        if (rng.chance(P.p_tx_split)) split_tx(tw4, th4);
        rec.tx = (uint8_t)tx_from_dims(tw4, th4);
        const bool pal = w4 <= 16 && h4 <= 16 && rng.chance(P.p_palette);
        int mode = rng.range(13), delta = 0;
        if (!pal && w4 <= 8 && h4 <= 8 && rng.chance(P.p_filter_intra)) { mode = DAV1D_CUDA_INTRA_FILTER; delta = rng.range(5); }
        else if (mode >= 1 && mode <= 8) delta = rng.irange(-3, 3);
        if (pal && real) { mode = 0; delta = 0; }               // a palette block is coded as DC_PRED
        rec.y_mode = (uint8_t)mode; rec.y_angle = (int8_t)delta; rec.pal_sz[0] = pal ? 8 : 0;
        uint32_t pal_off = 0, idx_off = 0;
        if (pal) {
            pal_off = (uint32_t)pal_px_alloc();
            idx_off = (uint32_t)pal_idx_alloc(w4 * 4 * h4 * 4 / 2);
            rec.pal_off[0] = pal_off; rec.pal_idx_off[0] = idx_off;
            add_intra(0, bx4, by4, w4, h4, DAV1D_CUDA_INTRA_PAL, 0, 0, false, pal_off, idx_off);
        }
        for (int y = 0; y < h4; y += th4)
            for (int x = 0; x < w4; x += tw4) {
                if (pal) {
                    if (residual) add_intra(0, bx4 + x, by4 + y, tw4, th4, DAV1D_CUDA_INTRA_NONE, 0, 0, true);
                    else mark(0, bx4 + x, by4 + y, tw4, th4);
                } else {
                    add_intra(0, bx4 + x, by4 + y, tw4, th4, mode, delta, flags, residual);
                }
            }
        if (pal) mark(0, bx4, by4, w4, h4);
        // ---- chroma.  The reference (recon_tmpl.c:1209-1211): a block that is 4 luma pixels wide /
        // high carries the chroma of the pair it closes, i.e. only the odd partner has chroma.
        bool has_chroma = !P.no_chroma;
        int cx4 = bx4 >> sh, cy4 = by4 >> sv, cw4 = w4 >> sh, ch4 = h4 >> sv;
        if (real && has_chroma) {
            has_chroma = (w4 > sh || (bx4 & 1)) && (h4 > sv || (by4 & 1));
            cw4 = (w4 + sh) >> sh; ch4 = (h4 + sv) >> sv;
        }
        rec.has_chroma = has_chroma ? 1 : 0;
        const int ymode_ctx = mode == DAV1D_CUDA_INTRA_FILTER ? 0 : mode;   // y_mode_nofilt (decode.c:745-746)
        if (!has_chroma || cw4 <= 0 || ch4 <= 0) {
            if (real) {
                rec.n_ops = (uint32_t)intra.size() - rec.first_op;
                blocks.push_back(rec);
                ctx_set(bx4, by4, w4, h4, 1, ymode_ctx, false, 0);
            }
            return;
        }
        if (all_decoded(1, cx4 + cw4, cy4 - 1, std::min(cx4 + 2 * cw4, pw4[1]), cy4) && cy4 > 0) rec.edge_tr |= 2;
        if (all_decoded(1, cx4 - 1, cy4 + ch4, cx4, std::min(cy4 + 2 * ch4, ph4[1])) && cx4 > 0) rec.edge_bl |= 2;
        int uvtw4 = std::min(cw4, 8), uvth4 = std::min(ch4, 8);
        fit_tx(uvtw4, uvth4);
        rec.uvtx = (uint8_t)tx_from_dims(uvtw4, uvth4);
        const bool cfl = !pal && w4 <= 8 && h4 <= 8 && rng.chance(P.p_cfl);
        // sm_uv_flag(t->a, cbx4) | sm_uv_flag(&t->l, cby4) (ipred_prepare.h:103-107)
        const bool uvsm = real ? (smooth_mode(a_uvmode[cx4]) || smooth_mode(l_uvmode[cy4])) : rng.chance(0.25f);
        const int uvflags = (P.edge_filter ? 1024 : 0) | (uvsm ? 512 : 0);
        rec.sm_flags |= uvsm ? 2 : 0;
        int uvmode = rng.range(13), uvdelta = 0;
        if (uvmode >= 1 && uvmode <= 8) uvdelta = rng.irange(-3, 3);
        if (real && (pal || cfl)) { uvmode = 0; uvdelta = 0; }
        rec.uv_mode = (uint8_t)(cfl ? 13 : uvmode); rec.uv_angle = (int8_t)uvdelta; rec.pal_sz[1] = pal ? 8 : 0;
        uint32_t uvpal_off[2] = { 0, 0 }, uvidx_off = 0;
        if (pal) {
            uvpal_off[0] = (uint32_t)pal_px_alloc();
            uvpal_off[1] = (uint32_t)pal_px_alloc();
            uvidx_off = (uint32_t)pal_idx_alloc(cw4 * 4 * ch4 * 4 / 2);
            rec.pal_off[1] = uvpal_off[0]; rec.pal_off[2] = uvpal_off[1]; rec.pal_idx_off[1] = uvidx_off;
        }
        int wpad = 0, hpad = 0;
        // CfL padding (recon_tmpl.c:1388-1396) only arises where a block sticks out of the picture;
        // the random padding exercises the operator beyond what a stream can produce
        if (!real && cfl && rng.chance(0.1f)) { wpad = rng.range(uvtw4); hpad = rng.range(uvth4); }   // per CfL operation (tx block)
        for (int pl = 1; pl <= 2; pl++) {
            if (pal) add_intra(pl, cx4, cy4, cw4, ch4, DAV1D_CUDA_INTRA_PAL, 0, 0, false, uvpal_off[pl - 1], uvidx_off);
            int alpha = 0;
            if (cfl) alpha = (rng.range(16) + 1) * (rng.chance(0.5f) ? -1 : 1);
            if (cfl && pl == 2 && rng.chance(0.2f)) alpha = 0;    // alpha 0 -> plain DC_PRED (recon_tmpl.c:1479)
            rec.cfl_alpha[pl - 1] = (int8_t)alpha;
            for (int y = 0; y < ch4; y += uvth4)
                for (int x = 0; x < cw4; x += uvtw4) {
                    if (pal) {
                        if (residual) add_intra(pl, cx4 + x, cy4 + y, uvtw4, uvth4, DAV1D_CUDA_INTRA_NONE, 0, 0, true);
                    } else if (cfl && alpha) {
                        add_intra(pl, cx4 + x, cy4 + y, uvtw4, uvth4, DAV1D_CUDA_INTRA_CFL, alpha, 0, residual,
                                  (uint32_t)(wpad | (hpad << 8)));
                    } else {
                        add_intra(pl, cx4 + x, cy4 + y, uvtw4, uvth4, cfl ? 0 : uvmode, cfl ? 0 : uvdelta,
                                  uvflags, residual);
                    }
                }
            mark(pl, cx4, cy4, cw4, ch4);
        }
        if (real) {
            rec.n_ops = (uint32_t)intra.size() - rec.first_op;
            blocks.push_back(rec);
            ctx_set(bx4, by4, w4, h4, 1, ymode_ctx, true, cfl ? 13 : uvmode);
        }
    }

    size_t pal_px_alloc() {
        const size_t o = pal.size();
        for (int i = 0; i < 8; i++) pal.push_back((uint16_t)(rng.u32() & P.bitdepth_max));
        return o;
    }
    size_t pal_idx_alloc(int bytes) {
        const size_t o = pal_idx.size();
        for (int i = 0; i < bytes; i++) pal_idx.push_back((uint8_t)(rng.u32() & 0x77));
        return o;
    }

    Dav1dCudaMcSrc make_src(int pl, int bx4, int by4, int ref, int mvx, int mvy, int filter) {
        const int sh = pl ? P.ss_hor : 0, sv = pl ? P.ss_ver : 0;   // recon_tmpl.c:966-977,1003
        Dav1dCudaMcSrc s;
        s.ref = (uint8_t)ref; s.filter_2d = (uint8_t)filter;
        const int mx = mvx & (15 >> !sh), my = mvy & (15 >> !sv);
        s.mx = (uint8_t)(mx << !sh); s.my = (uint8_t)(my << !sv);
        s.x = bx4 * (4 >> sh) + (mvx >> (3 + sh));
        s.y = by4 * (4 >> sv) + (mvy >> (3 + sv));
        return s;
    }

    // ---- references of another size (f->svc[ref][0/1], decode.c:3517-3524)
    std::vector<Dav1dCudaMcScaledDesc> sc[4];
    bool ref_scaled(int r) const {
        return (P.ref_w[r] && P.ref_w[r] != P.w) || (P.ref_h[r] && P.ref_h[r] != P.h);
    }
    static int scale_fac(int ref_sz, int cur_sz) { return ((ref_sz << 14) + (cur_sz >> 1)) / cur_sz; }
    // the unscaled source -> what the scaled branch computes for it (recon_tmpl.c:1013-1021): orig_pos in
    // 1/16 sample = (integer position << 4) + phase
    Dav1dCudaMcScaledSrc scaled_src(const Dav1dCudaMcSrc &u) const {
        Dav1dCudaMcScaledSrc o;
        memset(&o, 0, sizeof(o));
        o.ref = u.ref; o.filter_2d = u.filter_2d;
        const int orig[2] = { u.x * 16 + u.mx, u.y * 16 + u.my };
        int32_t *pos[2] = { &o.pos_x, &o.pos_y }, *step[2] = { &o.step_x, &o.step_y };
        for (int k = 0; k < 2; k++) {
            if (!ref_scaled(u.ref)) { *pos[k] = orig[k] * 64; *step[k] = 1024; continue; }
            const int ref_sz = k ? (P.ref_h[u.ref] ? P.ref_h[u.ref] : P.h) : (P.ref_w[u.ref] ? P.ref_w[u.ref] : P.w);
            const int scale = scale_fac(ref_sz, k ? P.h : P.w);
            const int64_t tmp = (int64_t)orig[k] * scale + (int64_t)(scale - 0x4000) * 8;
            const int mag = (int)(((tmp < 0 ? -tmp : tmp) + 128) >> 8);
            *pos[k] = (tmp < 0 ? -mag : mag) + 32;
            *step[k] = (scale + 8) >> 4;
        }
        return o;
    }
    // a prediction goes to the scaled list when a reference it reads has another size
    bool route_scaled(const Dav1dCudaMcDesc &d, int section) {
        const bool two = d.kind != DAV1D_CUDA_MC_PUT && d.kind != DAV1D_CUDA_MC_OBMC_H && d.kind != DAV1D_CUDA_MC_OBMC_V;
        if (!ref_scaled(d.src[0].ref) && !(two && ref_scaled(d.src[1].ref))) return false;
        Dav1dCudaMcScaledDesc o;
        memset(&o, 0, sizeof(o));
        o.x = d.x; o.y = d.y; o.w = d.w; o.h = d.h; o.plane = d.plane; o.kind = d.kind;
        o.src[0] = scaled_src(d.src[0]);
        if (two) o.src[1] = scaled_src(d.src[1]);
        o.weight = d.weight; o.mask_ss = d.mask_ss; o.aux16 = d.aux16; o.aux_off = d.aux_off;
        order.push_back({ (uint8_t)(8 + section), (uint32_t)sc[section].size() });
        sc[section].push_back(o);
        return true;
    }

    // obmc() (recon_tmpl.c:1071-1131) with random neighbours: per top / left neighbour (width or
    // height step4 in {2,4,8,16}, inter with probability 0.7, at most min(log2(dim), 4) of them) a
    // prediction with the neighbour's motion vector + blend_h / blend_v
    void add_obmc(int pl, int bx4, int by4, int w4, int h4) {
        const int sh = pl ? P.ss_hor : 0, sv = pl ? P.ss_ver : 0;
        const int h_mul = 4 >> sh, v_mul = 4 >> sv;
        const int R = P.mv_range * 8;
        auto ilog2 = [](int v) { int l = 0; while (v > 1) { v >>= 1; l++; } return l; };
        struct NbMv { int ref, mvx, mvy, filter; };
        auto neighbour = [&](int x4, int y4, int ow4, int mh4, int kind, int blend_h, const NbMv *nb = nullptr) {
            Dav1dCudaMcDesc d;
            memset(&d, 0, sizeof(d));
            d.plane = (uint8_t)pl; d.kind = (uint8_t)kind;
            d.x = (uint16_t)(((bx4 * 4) >> sh) + (x4 - bx4) * h_mul);
            d.y = (uint16_t)(((by4 * 4) >> sv) + (y4 - by4) * v_mul);
            d.w = (uint8_t)(ow4 * h_mul); d.h = (uint8_t)(mh4 * v_mul);
            d.aux16 = (uint16_t)blend_h;
            if (nb) d.src[0] = make_src(pl, x4, y4, nb->ref, nb->mvx, nb->mvy, nb->filter);
            else d.src[0] = make_src(pl, x4, y4, rng.range(P.n_refs), rng.irange(-R, R), rng.irange(-R, R), rng.range(10));
            add_bytes(0, 4.0 * Bp * d.w * d.h);     // reference read + blend read-modify-write
            if (route_scaled(d, kind == DAV1D_CUDA_MC_OBMC_H ? 2 : 3)) return;
            if (kind == DAV1D_CUDA_MC_OBMC_H) { order.push_back({ 6, (uint32_t)obmc_h.size() }); obmc_h.push_back(d); }
            else { order.push_back({ 7, (uint32_t)obmc_v.size() }); obmc_v.push_back(d); }
        };
        if (P.real_blocks) {
            // obmc() (recon_tmpl.c:1071-1131) over the ACTUAL neighbours: the blocks whose bottom row /
            // right column lies along this block's top / left edge (what the refmvs rows hold), looked at
            // at the odd 4x4 position of every step; inter neighbours only
            if (by4 > tile_y0 && (!pl || w4 * h_mul + h4 * v_mul >= 16)) {
                for (int i = 0, x = 0; x < w4 && i < std::min(ilog2(w4), 4);) {
                    const Nb &a = nb_above[bx4 + x + 1];
                    const int step4 = std::min(std::max((int)a.w4, 2), 16);
                    if (a.inter) {
                        const int ow4 = std::min(step4, w4), oh4 = std::min(h4, 16) >> 1;
                        const NbMv m = { a.ref, a.mvx, a.mvy, a.filter };
                        neighbour(bx4 + x, by4, ow4, (oh4 * 3 + 3) >> 2, DAV1D_CUDA_MC_OBMC_H, v_mul * oh4, &m);
                        i++;
                    }
                    x += step4;
                }
            }
            if (bx4 > tile_x0) {
                for (int i = 0, y = 0; y < h4 && i < std::min(ilog2(h4), 4);) {
                    const Nb &l = nb_left[by4 + y + 1];
                    const int step4 = std::min(std::max((int)l.h4, 2), 16);
                    if (l.inter) {
                        const int ow4 = std::min(w4, 16) >> 1, oh4 = std::min(step4, h4);
                        const NbMv m = { l.ref, l.mvx, l.mvy, l.filter };
                        neighbour(bx4, by4 + y, ow4, oh4, DAV1D_CUDA_MC_OBMC_V, 0, &m);
                        i++;
                    }
                    y += step4;
                }
            }
            return;
        }
        if (by4 > tile_y0 && (!pl || w4 * h_mul + h4 * v_mul >= 16)) {
            for (int i = 0, x = 0; x < w4 && i < std::min(ilog2(w4), 4);) {
                int step4 = 2 << rng.range(4);
                while (x > 0 && x % step4) step4 >>= 1;      // neighbours are aligned to their own size
                if (rng.chance(0.7f)) {
                    const int ow4 = std::min(step4, w4), oh4 = std::min(h4, 16) >> 1;
                    neighbour(bx4 + x, by4, ow4, (oh4 * 3 + 3) >> 2, DAV1D_CUDA_MC_OBMC_H, v_mul * oh4);
                    i++;
                }
                x += step4;
            }
        }
        if (bx4 > tile_x0) {
            for (int i = 0, y = 0; y < h4 && i < std::min(ilog2(h4), 4);) {
                int step4 = 2 << rng.range(4);
                while (y > 0 && y % step4) step4 >>= 1;
                if (rng.chance(0.7f)) {
                    const int ow4 = std::min(w4, 16) >> 1, oh4 = std::min(step4, h4);
                    neighbour(bx4, by4 + y, ow4, oh4, DAV1D_CUDA_MC_OBMC_V, 0);
                    i++;
                }
                y += step4;
            }
        }
    }

    void inter_block(int bx4, int by4, int w4, int h4) {
        const float u = rng.unit();
        float acc = P.p_avg;
        int kind = DAV1D_CUDA_MC_PUT;
        bool is_warp = false;
        if (u < acc) kind = DAV1D_CUDA_MC_AVG;
        else if (u < (acc += P.p_w_avg)) kind = DAV1D_CUDA_MC_W_AVG;
        else if (u < (acc += P.p_wedge)) kind = DAV1D_CUDA_MC_MASK;
        else if (u < (acc += P.p_seg)) kind = DAV1D_CUDA_MC_W_MASK;
        else if (u < (acc += P.p_warp) && w4 >= 4 && h4 >= 4) is_warp = true;
        const D1SynthMaskTab *const MT = (const D1SynthMaskTab *)(uintptr_t)P.mask_tab;
        auto l2 = [](int v) { return v == 2 ? 0 : v == 4 ? 1 : 2; };
        const bool mask_block = w4 >= 2 && h4 >= 2 && w4 <= 8 && h4 <= 8;        // BS_8x8 .. BS_32x32
        if (P.real_blocks && kind == DAV1D_CUDA_MC_MASK && !(MT && mask_block)) kind = DAV1D_CUDA_MC_AVG;
        const int wedge_idx = (P.real_blocks && MT) ? rng.range(16) : 0;      // no draw otherwise: older frames keep their streams
        bool any_scaled = false;
        for (int i = 0; i < 7; i++) any_scaled |= ref_scaled(i);
        const int filter = rng.range(10);
        const int R = P.mv_range * 8;
        int ref[2], mvx[2], mvy[2];
        for (int i = 0; i < 2; i++) {
            ref[i] = rng.range(P.n_refs);
            mvx[i] = rng.irange(-R, R); mvy[i] = rng.irange(-R, R);
            if (rng.chance(0.1f)) mvx[i] &= ~7;    // integer-pel columns / rows now and then
            if (rng.chance(0.1f)) mvy[i] &= ~7;
        }
        if (is_warp && any_scaled && ref_scaled(ref[0])) is_warp = false;   // allow_warp needs a same-size reference (decode.c:1828)
        const D1SynthWarp *const WT = (const D1SynthWarp *)(uintptr_t)P.warp_tab;
        if (is_warp && P.real_blocks && !(WT && P.n_warp_tab > 0)) is_warp = false;
        D1SynthWarp wm;
        memset(&wm, 0, sizeof(wm));
        int warp_global = 0;
        if (is_warp && P.real_blocks) {
            // one of the caller's models, translated so that the block lands where its vector points
            const int wk = rng.range(P.n_warp_tab);
            wm = WT[wk];
            warp_global = wk & 1;      // odd entries: coded as GLOBALMV with the model as the reference's global motion
            const int cx = bx4 * 4 + w4 * 2, cy = by4 * 4 + h4 * 2;
            wm.matrix[0] = mvx[0] * 8192 - (int32_t)(((int64_t)(wm.matrix[2] - 0x10000) * cx + (int64_t)wm.matrix[3] * cy));
            wm.matrix[1] = mvy[0] * 8192 - (int32_t)(((int64_t)wm.matrix[4] * cx + (int64_t)(wm.matrix[5] - 0x10000) * cy));
        }
        int weight = rng.irange(1, 15);
        const int sign = rng.range(2);
        if (P.real_blocks) weight = jnt_weight_of(ref[0], ref[1]);    // COMP_INTER_WEIGHTED_AVG takes it from the frame
        D1SynthBlock rec_inter;
        memset(&rec_inter, 0, sizeof(rec_inter));
        rec_inter.bx4 = (uint16_t)bx4; rec_inter.by4 = (uint16_t)by4; rec_inter.w4 = (uint8_t)w4; rec_inter.h4 = (uint8_t)h4;
        rec_inter.intra = 0; rec_inter.has_chroma = P.no_chroma ? 0 : 1; rec_inter.skip = 1; rec_inter.tile = (uint8_t)tile_no;
        rec_inter.tile_x0 = (uint16_t)tile_x0; rec_inter.tile_y0 = (uint16_t)tile_y0;
        rec_inter.tile_x1 = (uint16_t)std::min(tile_x1, bw4); rec_inter.tile_y1 = (uint16_t)std::min(tile_y1, bh4);
        for (int i = 0; i < 2; i++) { rec_inter.mvx[i] = (int16_t)mvx[i]; rec_inter.mvy[i] = (int16_t)mvy[i]; rec_inter.ref[i] = (uint8_t)ref[i]; }
        rec_inter.comp_kind = (uint8_t)(is_warp ? 255 : kind); rec_inter.filter2d = (uint8_t)filter;
        rec_inter.mask_sign = (uint8_t)sign; rec_inter.jnt_weight = (uint8_t)weight;
        rec_inter.first_tx = (uint32_t)tx_recs.size();
        {   // the largest transforms of the block (b->max_ytx, b->uvtx) also when it carries no residual
            int a = std::min(w4, 16), b2 = std::min(h4, 16);
            rec_inter.max_ytx = (uint8_t)tx_from_dims(a, b2);
            int ua = std::min(std::max(1, w4 >> P.ss_hor), 8), ub = std::min(std::max(1, h4 >> P.ss_ver), 8);
            fit_tx(ua, ub);
            rec_inter.uvtx = (uint8_t)tx_from_dims(ua, ub);
        }
        // OBMC: single-reference, translational blocks of at least 8x8 on even 4x4 coordinates
        const bool do_obmc = P.p_obmc > 0.f && kind == DAV1D_CUDA_MC_PUT && !is_warp && w4 >= 2 && h4 >= 2 &&
                             !(bx4 & 1) && !(by4 & 1) && rng.chance(P.p_obmc);
        // inter-intra: single-reference translational blocks of 8x8..32x32 without OBMC
        // (chroma prediction blocks must keep an aspect ratio of at most 4: an 8x32 block in 4:2:2 would
        // need a 4x32 predictor, which the reference does not have - recon_tmpl.c:1780-1782)
        const int ii_cw4 = std::max(1, w4 >> P.ss_hor), ii_ch4 = std::max(1, h4 >> P.ss_ver);
        const bool do_ii = P.p_ii > 0.f && kind == DAV1D_CUDA_MC_PUT && !is_warp && !do_obmc && w4 >= 2 && h4 >= 2 &&
                           w4 <= 8 && h4 <= 8 && (P.no_chroma || (ii_cw4 <= 4 * ii_ch4 && ii_ch4 <= 4 * ii_cw4)) &&
                           (!P.real_blocks || MT) && rng.chance(P.p_ii);
        uint32_t seg_off = 0, wedge_off[3] = { 0, 0, 0 };
        for (int pl = 0; pl < nplanes(); pl++) {
            const int sh = pl ? P.ss_hor : 0, sv = pl ? P.ss_ver : 0;
            const int w = (w4 * 4) >> sh, h = (h4 * 4) >> sv;
            const int x = (bx4 * 4) >> sh, y = (by4 * 4) >> sv;
            if (is_warp && P.real_blocks) {
                // warp_affine() (recon_tmpl.c:1134-1193) over the block's 8x8 units of this plane
                for (int yy = 0; yy < h; yy += 8) {
                    const int src_y = by4 * 4 + ((yy + 4) << sv);
                    const int64_t mat3_y = (int64_t)wm.matrix[3] * src_y + wm.matrix[0];
                    const int64_t mat5_y = (int64_t)wm.matrix[5] * src_y + wm.matrix[1];
                    for (int xx = 0; xx < w; xx += 8) {
                        const int src_x = bx4 * 4 + ((xx + 4) << sh);
                        const int64_t wx = ((int64_t)wm.matrix[2] * src_x + mat3_y) >> sh;
                        const int64_t wy = ((int64_t)wm.matrix[4] * src_x + mat5_y) >> sv;
                        Dav1dCudaWarpDesc d;
                        memset(&d, 0, sizeof(d));
                        d.plane = (uint8_t)pl; d.ref = (uint8_t)ref[0];
                        d.x = (uint16_t)(x + xx); d.y = (uint16_t)(y + yy);
                        d.sx = (int)(wx >> 16) - 4; d.sy = (int)(wy >> 16) - 4;
                        d.mx = (((int)wx & 0xffff) - wm.abcd[0] * 4 - wm.abcd[1] * 7) & ~0x3f;
                        d.my = (((int)wy & 0xffff) - wm.abcd[2] * 4 - wm.abcd[3] * 4) & ~0x3f;
                        memcpy(d.abcd, wm.abcd, sizeof(d.abcd));
                        order.push_back({ 2, (uint32_t)warp.size() });
                        warp.push_back(d);
                    }
                }
                add_bytes(2, 2.0 * Bp * w * h);
                continue;
            }
            if (is_warp) {
                // per-8x8 calls of warp_affine() (recon_tmpl.c:1151-1191) with random shear parameters
                int16_t abcd[4];
                for (int k = 0; k < 4; k++) abcd[k] = (int16_t)(rng.irange(0, 0x1fff) - 0xa00);
                for (int yy = 0; yy < h; yy += 8)
                    for (int xx = 0; xx < w; xx += 8) {
                        Dav1dCudaWarpDesc d;
                        memset(&d, 0, sizeof(d));
                        d.plane = (uint8_t)pl; d.ref = (uint8_t)ref[0];
                        d.x = (uint16_t)(x + xx); d.y = (uint16_t)(y + yy);
                        d.sx = x + xx + ((mvx[0] >> (3 + sh))); d.sy = y + yy + ((mvy[0] >> (3 + sv)));
                        d.mx = (rng.irange(0, 0x1fff) - 0xa00) & ~0x3f;
                        d.my = (rng.irange(0, 0x1fff) - 0xa00) & ~0x3f;
                        memcpy(d.abcd, abcd, sizeof(abcd));
                        order.push_back({ 2, (uint32_t)warp.size() });
                        warp.push_back(d);
                    }
                add_bytes(2, 2.0 * Bp * w * h);
                continue;
            }
            Dav1dCudaMcDesc d;
            memset(&d, 0, sizeof(d));
            d.x = (uint16_t)x; d.y = (uint16_t)y; d.w = (uint8_t)w; d.h = (uint8_t)h;
            d.plane = (uint8_t)pl; d.kind = (uint8_t)kind;
            d.src[0] = make_src(pl, bx4, by4, ref[0], mvx[0], mvy[0], filter);
            d.src[1] = make_src(pl, bx4, by4, ref[1], mvx[1], mvy[1], filter);
            bool wave1 = false;
            if (kind == DAV1D_CUDA_MC_PUT) {
                add_bytes(0, 2.0 * Bp * w * h);
                if (!route_scaled(d, 0)) {
                    order.push_back({ 0, (uint32_t)put.size() });
                    put.push_back(d);
                }
                if (do_obmc) add_obmc(pl, bx4, by4, w4, h4);
                continue;
            }
            add_bytes(1, 3.0 * Bp * w * h);
            if (kind == DAV1D_CUDA_MC_W_AVG) d.weight = (uint8_t)weight;
            if (kind == DAV1D_CUDA_MC_MASK || kind == DAV1D_CUDA_MC_W_MASK) {
                if (sign) std::swap(d.src[0], d.src[1]);   // tmp[mask_sign], tmp[!mask_sign]
            }
            if (kind == DAV1D_CUDA_MC_MASK) {               // wedge: one mask table per plane size
                wedge_off[pl] = (uint32_t)masks.size();
                if (P.real_blocks) {
                    // recon_tmpl.c:1861-1866: luma WEDGE_MASK(0, bs, 0, idx), chroma WEDGE_MASK(layout, bs, sign, idx)
                    const int lay = pl ? (P.ss_hor ? (P.ss_ver ? 2 : 1) : 0) : 0;
                    const uint8_t *m = MT->base + MT->wedge[lay][l2(w4)][l2(h4)][pl ? sign : 0][wedge_idx];
                    masks.insert(masks.end(), m, m + w * h);
                } else
                for (int i = 0; i < w * h; i++) masks.push_back((uint8_t)rng.range(65));
                d.aux_off = wedge_off[pl];
                add_bytes(1, (double)w * h);
            } else if (kind == DAV1D_CUDA_MC_W_MASK) {
                if (pl == 0) {
                    const int lay = P.no_chroma ? 0 : (P.ss_hor ? (P.ss_ver ? 2 : 1) : 0);
                    const int mw = w >> (lay >= 1), mh = h >> (lay == 2);
                    seg_off = (uint32_t)masks.size();
                    masks.resize(masks.size() + (size_t)mw * mh, 0);
                    d.aux_off = seg_off; d.mask_ss = (uint8_t)lay; d.weight = (uint8_t)sign;
                    add_bytes(1, (double)mw * mh);
                } else {
                    d.kind = DAV1D_CUDA_MC_MASK;            // chroma reuses the luma-derived mask
                    d.aux_off = seg_off;
                    wave1 = true;
                    add_bytes(1, (double)w * h);
                }
            }
            if (route_scaled(d, wave1 ? 1 : 0)) continue;
            if (wave1) { order.push_back({ 5, (uint32_t)comp1.size() }); comp1.push_back(d); }
            else { order.push_back({ 1, (uint32_t)comp0.size() }); comp0.push_back(d); }
        }
        if (do_ii) {
            // intra prediction of the whole block blended onto the inter prediction; the block then
            // belongs to the wavefront: its residuals follow as residual-only intra-class operations
            static const int ii_modes[4] = { 0, 1, 2, 9 };          // DC, VERT, HOR, SMOOTH
            const int ii_mode = rng.range(4), ii_wedge = P.real_blocks ? rng.chance(0.4f) : 0;
            const int m = ii_modes[ii_mode];
            rec_inter.pad[1] = (uint8_t)wedge_idx;
            rec_inter.pad[2] = (uint8_t)((ii_wedge ? 2 : 1) | (ii_mode << 2));   // b->interintra_type | b->interintra_mode << 2
            for (int pl = 0; pl < nplanes(); pl++) {
                const int sh = pl ? P.ss_hor : 0, sv = pl ? P.ss_ver : 0;
                const int cw4 = w4 >> sh, ch4 = h4 >> sv;
                const size_t off = pal_idx.size();
                if (P.real_blocks) {
                    // II_MASK(layout of the plane, bs, b) (wedge.h:88-93): the smooth / wedge blend mask of the block size
                    const int lay = pl ? (P.ss_hor ? (P.ss_ver ? 2 : 1) : 0) : 0;
                    const uint8_t *mk = MT->base + (ii_wedge ? MT->wedge[lay][l2(w4)][l2(h4)][0][wedge_idx]
                                                             : MT->ii[lay][l2(w4)][l2(h4)][ii_mode]);
                    pal_idx.insert(pal_idx.end(), mk, mk + cw4 * 4 * ch4 * 4);
                } else
                for (int i = 0; i < cw4 * 4 * ch4 * 4; i++) pal_idx.push_back((uint8_t)rng.range(65));
                add_intra(pl, bx4 >> sh, by4 >> sv, cw4, ch4, DAV1D_CUDA_INTRA_II, m, 0, false, 0, (uint32_t)off);
            }
        }
        // ---- residual
        auto rec_tx_of_last_op = [&]() {                  // the cbi / cf entry of the residual-only operation just added
            if (!P.real_blocks) return;
            const Dav1dCudaIntraDesc &o = intra.back();
            D1SynthTx t;
            memset(&t, 0, sizeof(t));
            t.coef_off = o.coef_off; t.eob = o.eob; t.txtp = o.txtp; t.cw4 = o.cw4; t.ch4 = o.ch4; t.tx = o.tx; t.plane = o.plane;
            tx_recs.push_back(t);
        };
        if (do_ii) {
            if (rng.chance(P.p_residual)) {
                int tw4 = std::min(w4, 16), th4 = std::min(h4, 16);
                rec_inter.skip = 0;
                if (rng.chance(P.p_tx_split)) { split_tx(tw4, th4); rec_inter.tx_split = tw4 * th4 < std::min(w4, 16) * std::min(h4, 16); }
                for (int y = 0; y < h4; y += th4)
                    for (int x = 0; x < w4; x += tw4) {
                        add_intra(0, bx4 + x, by4 + y, tw4, th4, DAV1D_CUDA_INTRA_NONE, 0, 0, true);
                        rec_tx_of_last_op();
                    }
                if (!P.no_chroma) {
                    const int cw4 = w4 >> P.ss_hor, ch4 = h4 >> P.ss_ver;
                    int utw4 = std::min(cw4, 8), uth4 = std::min(ch4, 8);
                    fit_tx(utw4, uth4);
                    for (int pl = 1; pl <= 2; pl++)
                        for (int y = 0; y < ch4; y += uth4)
                            for (int x = 0; x < cw4; x += utw4)
                            {
                                add_intra(pl, (bx4 >> P.ss_hor) + x, (by4 >> P.ss_ver) + y, utw4, uth4,
                                          DAV1D_CUDA_INTRA_NONE, 0, 0, true);
                                rec_tx_of_last_op();
                            }
                }
            }
        } else if (rng.chance(P.p_residual)) {
            int tw4 = std::min(w4, 16), th4 = std::min(h4, 16);
            rec_inter.max_ytx = (uint8_t)tx_from_dims(tw4, th4);
            rec_inter.skip = 0;
            if (rng.chance(P.p_tx_split)) { split_tx(tw4, th4); rec_inter.tx_split = tw4 * th4 < std::min(w4, 16) * std::min(h4, 16); }
            const int tx = tx_from_dims(tw4, th4);
            for (int y = 0; y < h4; y += th4)
                for (int x = 0; x < w4; x += tw4) add_itx(0, bx4 + x, by4 + y, tx);
            if (!P.no_chroma) {
                const int cw4 = w4 >> P.ss_hor, ch4 = h4 >> P.ss_ver;
                int utw4 = std::min(cw4, 8), uth4 = std::min(ch4, 8);
                fit_tx(utw4, uth4);
                const int utx = tx_from_dims(utw4, uth4);
                rec_inter.uvtx = (uint8_t)utx;
                for (int pl = 1; pl <= 2; pl++)
                    for (int y = 0; y < ch4; y += uth4)
                        for (int x = 0; x < cw4; x += utw4)
                            add_itx(pl, (bx4 >> P.ss_hor) + x, (by4 >> P.ss_ver) + y, utx);
            }
        }
        if (P.real_blocks) {
            rec_inter.n_tx = (uint32_t)tx_recs.size() - rec_inter.first_tx;
            rec_inter.pad[0] = do_obmc ? 1 : 0;                 // b->motion_mode == MM_OBMC
            if (kind == DAV1D_CUDA_MC_MASK) rec_inter.pad[1] = (uint8_t)wedge_idx;   // b->wedge_idx
            rec_inter.warp = wm;
            if (is_warp) rec_inter.pad[1] = (uint8_t)warp_global;      // 1: inter_mode == GLOBALMV + gmv_warp_allowed, 0: MM_WARP
            blocks.push_back(rec_inter);
            nb_set(bx4, by4, w4, h4, 1, ref[0], mvx[0], mvy[0], filter);
            grid_set(bx4, by4, w4, h4, Nb{ 1, (uint8_t)ref[0], (uint8_t)filter, (uint8_t)w4, (uint8_t)h4, (int16_t)mvx[0], (int16_t)mvy[0] });
        }
        for (int pl = 0; pl < nplanes(); pl++) {
            const int sh = pl ? P.ss_hor : 0, sv = pl ? P.ss_ver : 0;
            mark(pl, bx4 >> sh, by4 >> sv, w4 >> sh, h4 >> sv);
        }
    }

    void block(int bx4, int by4, int w4, int h4) {
        n_blocks++;
        luma_px += 16.0 * w4 * h4;
        if (rng.chance(P.p_intra)) {
            intra_block(bx4, by4, w4, h4);
            if (P.real_blocks) { nb_set(bx4, by4, w4, h4, 0, 0, 0, 0, 0); grid_set(bx4, by4, w4, h4, Nb{ 0, 0, 0, (uint8_t)w4, (uint8_t)h4, 0, 0 }); }
        } else {
            if (P.real_blocks && (w4 == 1 || h4 == 1)) inter_block_sub8x8(bx4, by4, w4, h4);
            else inter_block(bx4, by4, w4, h4);
            if (P.real_blocks) {            // decode.c:810-830: intra = 0, uvmode = DC_PRED
                const bool hc = (w4 > P.ss_hor || (bx4 & 1)) && (h4 > P.ss_ver || (by4 & 1));
                ctx_set(bx4, by4, w4, h4, 0, 0, hc, 0);
            }
        }
    }

    // recursive partition of an s4 x s4 (4-px units) square at (bx4, by4), decode (Z) order
    void partition(int bx4, int by4, int s4) {
        if (bx4 >= bw4 || by4 >= bh4) return;
        const bool fits = bx4 + s4 <= bw4 && by4 + s4 <= bh4;
        int choice;   // 0 NONE 1 H 2 V 3 SPLIT 4 H4 5 V4
        const bool sub_ok = P.real_blocks && P.p_sub8x8 > 0.f && P.ss_hor && !P.no_chroma;
        if (s4 <= 1) choice = 0;
        else if (!fits) choice = 3;
        else if (s4 == 2) choice = (sub_ok && rng.chance(P.p_sub8x8)) ? (P.ss_ver ? 1 + rng.range(3) : 1 + 2 * rng.range(2)) : 0;   // 4:2:2: 8x4 or 4x4
        else {
            const int r = rng.range(100);
            if (s4 == 16) choice = r < 8 ? 0 : r < 16 ? 1 : r < 24 ? 2 : r < 28 ? 4 : r < 32 ? 5 : 3;
            else if (s4 == 8) choice = r < 25 ? 0 : r < 38 ? 1 : r < 51 ? 2 : r < 56 ? 4 : r < 61 ? 5 : 3;
            else choice = r < 40 ? 0 : r < 55 ? 1 : r < 70 ? 2 : 3;      // s4 == 4 (16x16)
        }
        // 4:2:2 has no block whose chroma would be 1:4 or narrower: no vertical two- / four-way split
        // (the 0 entries of dav1d_max_txfm_size_for_bs, tables.c:171-195)
        if (P.real_blocks && P.ss_hor && !P.ss_ver && !P.no_chroma) choice = choice == 2 ? 1 : choice == 5 ? 4 : choice;
        // frames with inter blocks: no 4-pixel-wide / -high blocks (their chroma is predicted with the
        // neighbours' motion vectors from the refmvs rows, recon_tmpl.c:1683-1751, which the records do not carry)
        if (P.real_blocks && P.p_intra < 1.f) choice = choice == 4 ? 1 : choice == 5 ? (P.ss_hor && !P.ss_ver && !P.no_chroma ? 1 : 2) : choice;
        const int hs = s4 >> 1, q = s4 >> 2;
        switch (choice) {
        case 0: block(bx4, by4, s4, s4); break;
        case 1: block(bx4, by4, s4, hs); block(bx4, by4 + hs, s4, hs); break;
        case 2: block(bx4, by4, hs, s4); block(bx4 + hs, by4, hs, s4); break;
        case 4: for (int i = 0; i < 4; i++) block(bx4, by4 + i * q, s4, q); break;
        case 5: for (int i = 0; i < 4; i++) block(bx4 + i * q, by4, q, s4); break;
        default:
            partition(bx4, by4, hs); partition(bx4 + hs, by4, hs);
            partition(bx4, by4 + hs, hs); partition(bx4 + hs, by4 + hs, hs);
            break;
        }
    }

    void run() {
        if (P.only_tx >= 0) {   // config 2: plane tiled with one transform size, residual on top of a put
            const int tw4 = TXW4[P.only_tx], th4 = TXH4[P.only_tx];
            for (int y = 0; y + th4 <= bh4; y += th4)
                for (int x = 0; x + tw4 <= bw4; x += tw4) { add_itx(0, x, y, P.only_tx); luma_px += 16.0 * tw4 * th4; }
            return;
        }
        // tiles in raster order, superblocks in raster order inside a tile (the decode order of
        // dav1d_decode_tile_sbrow over the tiles of a frame)
        const int sbw = (bw4 + 15) >> 4, sbh = (bh4 + 15) >> 4;
        const int tc = std::max(1, std::min(P.tile_cols, sbw)), tr = std::max(1, std::min(P.tile_rows, sbh));
        const int tw = (sbw + tc - 1) / tc, th = (sbh + tr - 1) / tr;
        for (int ty = 0; ty * th < sbh; ty++)
            for (int tx = 0; tx * tw < sbw; tx++) {
                tile_x0 = tx * tw * 16; tile_y0 = ty * th * 16;
                tile_x1 = std::min((tx + 1) * tw * 16, bw4); tile_y1 = std::min((ty + 1) * th * 16, bh4);
                tile_no = ty * tc + tx;
                if (P.real_blocks) ctx_reset_above();
                for (int y = tile_y0; y < tile_y1; y += 16) {
                    if (P.real_blocks) ctx_reset_left(y);
                    for (int x = tile_x0; x < tile_x1; x += 16) partition(x, y, 16);
                }
            }
    }
    int tile_x0 = 0, tile_y0 = 0, tile_x1 = 1 << 20, tile_y1 = 1 << 20;   // current tile, luma 4-px units
    int tile_no = 0;
    // above / left block contexts (BlockContext mode / intra / uvmode, src/env.h), reset like
    // dav1d_reset_context() at a tile's top edge and at the left edge of every superblock row of a tile
    std::vector<uint8_t> a_intra, a_mode, a_uvmode, l_intra, l_mode, l_uvmode;
    std::vector<D1SynthBlock> blocks;
    std::vector<D1SynthTx> tx_recs;
    // what the refmvs rows and the filter contexts tell obmc() about the block above a column / left of a row
    struct Nb { uint8_t inter, ref, filter, w4, h4; int16_t mvx, mvy; };
    std::vector<Nb> nb_above, nb_left;
    std::vector<Nb> nb_grid;            // per 4x4 of the frame: what the refmvs rows / f->frame_thread.b hold (sub8x8 chroma)
    void grid_set(int bx4, int by4, int w4, int h4, const Nb &n) {
        if (nb_grid.empty()) return;
        for (int y = by4; y < std::min(by4 + h4, bh4); y++)
            for (int x = bx4; x < std::min(bx4 + w4, bw4); x++) nb_grid[(size_t)y * bw4 + x] = n;
    }

    // A 4-px-wide and / or 4-px-high inter block of a 4:2:0 / 4:2:2 frame (recon_tmpl.c:1638-1657 luma, :1683-1751 chroma):
    // single reference, translation.  The block at the odd position of its 8x8 carries the chroma of the whole
    // 8x8: when the other blocks of the 8x8 it looks at are inter blocks too, every 2x2 / 2x4 / 4x2 chroma part is
    // predicted with the vector, reference and filter of the luma block above it; otherwise the whole 4x4 chroma
    // block with this block's vector.
    void inter_block_sub8x8(int bx4, int by4, int w4, int h4) {
        const int R = P.mv_range * 8;
        const int filter = rng.range(10), ref = rng.range(P.n_refs);
        int mvx = rng.irange(-R, R), mvy = rng.irange(-R, R);
        if (rng.chance(0.1f)) mvx &= ~7;
        if (rng.chance(0.1f)) mvy &= ~7;
        const int sv = P.ss_ver;
        const bool hc = (w4 > 1 || (bx4 & 1)) && (h4 > sv || (by4 & 1));
        if (hc && !(w4 == 1 || h4 == sv)) {          // nothing special about its chroma (4:2:2: an 8x4 block)
            inter_block(bx4, by4, w4, h4);
            return;
        }
        D1SynthBlock rec;
        memset(&rec, 0, sizeof(rec));
        rec.bx4 = (uint16_t)bx4; rec.by4 = (uint16_t)by4; rec.w4 = (uint8_t)w4; rec.h4 = (uint8_t)h4;
        rec.intra = 0; rec.has_chroma = hc ? 1 : 0; rec.skip = 1; rec.tile = (uint8_t)tile_no;
        rec.tile_x0 = (uint16_t)tile_x0; rec.tile_y0 = (uint16_t)tile_y0;
        rec.tile_x1 = (uint16_t)std::min(tile_x1, bw4); rec.tile_y1 = (uint16_t)std::min(tile_y1, bh4);
        rec.mvx[0] = (int16_t)mvx; rec.mvy[0] = (int16_t)mvy; rec.ref[0] = (uint8_t)ref;
        rec.comp_kind = DAV1D_CUDA_MC_PUT; rec.filter2d = (uint8_t)filter;
        rec.max_ytx = (uint8_t)tx_from_dims(w4, h4); rec.uvtx = 0;          // TX_4X4 chroma
        rec.first_tx = (uint32_t)tx_recs.size();
        auto put_desc = [&](int pl, int x, int y, int w, int h, const Dav1dCudaMcSrc &src) {
            Dav1dCudaMcDesc d;
            memset(&d, 0, sizeof(d));
            d.x = (uint16_t)x; d.y = (uint16_t)y; d.w = (uint8_t)w; d.h = (uint8_t)h;
            d.plane = (uint8_t)pl; d.kind = DAV1D_CUDA_MC_PUT;
            d.src[0] = src;
            add_bytes(0, 2.0 * Bp * w * h);
            if (!route_scaled(d, 0)) { order.push_back({ 0, (uint32_t)put.size() }); put.push_back(d); }
        };
        put_desc(0, bx4 * 4, by4 * 4, w4 * 4, h4 * 4, make_src(0, bx4, by4, ref, mvx, mvy, filter));
        if (hc) {
            const Nb own = { 1, (uint8_t)ref, (uint8_t)filter, (uint8_t)w4, (uint8_t)h4, (int16_t)mvx, (int16_t)mvy };
            auto at = [&](int x, int y) -> const Nb & { return nb_grid[(size_t)y * bw4 + x]; };
            bool sub = true;
            if (w4 == 1) sub = sub && at(bx4 - 1, by4).inter;
            if (h4 == sv) sub = sub && at(bx4, by4 - 1).inter;
            if (w4 == 1 && h4 == sv) sub = sub && at(bx4 - 1, by4 - 1).inter;
            const int cx0 = (bx4 >> 1) * 4, cy0 = ((by4 & ~sv) * 4) >> sv, cw = w4 * 2, ch = (h4 * 4) >> sv;
            if (sub) {
                int h_off = 0, v_off = 0;
                auto part = [&](int nbx, int nby, const Nb &n, int ox, int oy) {
                    for (int pl = 1; pl <= 2; pl++)
                        put_desc(pl, cx0 + ox, cy0 + oy, cw, ch, make_src(pl, nbx, nby, n.ref, n.mvx, n.mvy, n.filter));
                };
                if (w4 == 1 && h4 == sv) { part(bx4 - 1, by4 - 1, at(bx4 - 1, by4 - 1), 0, 0); v_off = 2; h_off = 2; }
                if (w4 == 1) { part(bx4 - 1, by4, at(bx4 - 1, by4), 0, v_off); h_off = 2; }
                if (h4 == sv) { part(bx4, by4 - 1, at(bx4, by4 - 1), h_off, 0); v_off = 2; }
                part(bx4, by4, own, h_off, v_off);
            } else {
                for (int pl = 1; pl <= 2; pl++)
                    put_desc(pl, cx0, cy0, ((w4 << (w4 == 1)) * 4) >> 1, ((h4 << (h4 == sv)) * 4) >> sv,
                             make_src(pl, bx4 & ~1, by4 & ~sv, ref, mvx, mvy, filter));
            }
        }
        const int cbh4 = (h4 + sv) >> sv;                 // the chroma block of the odd partner: 4 px wide, 4 * cbh4 high
        const int uvtx = tx_from_dims(1, cbh4);
        rec.uvtx = (uint8_t)uvtx;
        if (rng.chance(P.p_residual)) {
            int tw4 = w4, th4 = h4;
            rec.skip = 0;
            if (tw4 != th4 && rng.chance(P.p_tx_split)) { split_tx(tw4, th4); rec.tx_split = 1; }
            const int tx = tx_from_dims(tw4, th4);
            for (int y = 0; y < h4; y += th4)
                for (int x = 0; x < w4; x += tw4) add_itx(0, bx4 + x, by4 + y, tx);
            if (hc)
                for (int pl = 1; pl <= 2; pl++) add_itx(pl, bx4 >> 1, by4 >> sv, uvtx);
        }
        rec.n_tx = (uint32_t)tx_recs.size() - rec.first_tx;
        blocks.push_back(rec);
        nb_set(bx4, by4, w4, h4, 1, ref, mvx, mvy, filter);
        grid_set(bx4, by4, w4, h4, Nb{ 1, (uint8_t)ref, (uint8_t)filter, (uint8_t)w4, (uint8_t)h4, (int16_t)mvx, (int16_t)mvy });
        mark(0, bx4, by4, w4, h4);
        if (hc) { mark(1, bx4 >> 1, by4 >> sv, 1, (h4 + sv) >> sv); mark(2, bx4 >> 1, by4 >> sv, 1, (h4 + sv) >> sv); }
    }
    void nb_set(int bx4, int by4, int w4, int h4, int inter, int ref, int mvx, int mvy, int filter) {
        const Nb n = { (uint8_t)inter, (uint8_t)ref, (uint8_t)filter, (uint8_t)w4, (uint8_t)h4, (int16_t)mvx, (int16_t)mvy };
        for (int x = bx4; x < std::min(bx4 + w4, bw4); x++) nb_above[x] = n;
        for (int y = by4; y < std::min(by4 + h4, bh4); y++) nb_left[y] = n;
    }
    // f->jnt_weights[ref0][ref1] of the synthetic frame header (real_blocks)
    static int jnt_weight_of(int r0, int r1) { return 1 + (r0 * 7 + r1 * 3 + 4) % 15; }
    static bool smooth_mode(int m) { return m >= 9 && m <= 11; }
    void ctx_init() {
        a_intra.assign(bw4 + 1, 0); a_mode.assign(bw4 + 1, 0); a_uvmode.assign(bw4 + 1, 0);
        l_intra.assign(bh4 + 1, 0); l_mode.assign(bh4 + 1, 0); l_uvmode.assign(bh4 + 1, 0);
        nb_above.assign(bw4 + 2, Nb{ 0, 0, 0, 1, 1, 0, 0 }); nb_left.assign(bh4 + 2, Nb{ 0, 0, 0, 1, 1, 0, 0 });
        if (P.real_blocks && P.p_sub8x8 > 0.f) nb_grid.assign((size_t)bw4 * bh4, Nb{ 0, 0, 0, 1, 1, 0, 0 });
    }
    void ctx_reset_above() {
        for (int x = tile_x0; x < std::min(tile_x1, bw4); x++) { a_intra[x] = 0; a_mode[x] = 0; }
        for (int x = tile_x0 >> P.ss_hor; x < (std::min(tile_x1, bw4) + P.ss_hor) >> P.ss_hor; x++) a_uvmode[x] = 0;
    }
    void ctx_reset_left(int y0) {
        for (int y = y0; y < std::min(y0 + 16, bh4); y++) { l_intra[y] = 0; l_mode[y] = 0; }
        for (int y = y0 >> P.ss_ver; y < (std::min(y0 + 16, bh4) + P.ss_ver) >> P.ss_ver; y++) l_uvmode[y] = 0;
    }
    void ctx_set(int bx4, int by4, int w4, int h4, int intra, int ymode, bool has_chroma, int uvmode) {
        for (int x = bx4; x < std::min(bx4 + w4, bw4); x++) { a_intra[x] = (uint8_t)intra; if (intra) a_mode[x] = (uint8_t)ymode; }
        for (int y = by4; y < std::min(by4 + h4, bh4); y++) { l_intra[y] = (uint8_t)intra; if (intra) l_mode[y] = (uint8_t)ymode; }
        if (has_chroma && !P.no_chroma) {
            const int sh = P.ss_hor, sv = P.ss_ver;
            for (int x = bx4 >> sh; x < (bx4 >> sh) + ((w4 + sh) >> sh); x++) a_uvmode[x] = (uint8_t)uvmode;
            for (int y = by4 >> sv; y < (by4 >> sv) + ((h4 + sv) >> sv); y++) l_uvmode[y] = (uint8_t)uvmode;
        }
    }
};

int mc_tiles(uint32_t desc_index, int w, int h, uint32_t *out) {   // == dav1d_cuda_mc_tiles()
    int n = 0;
    for (int ty = 0; ty * 32 < h; ty++)
        for (int tx = 0; tx * 32 < w; tx++) out[n++] = desc_index * 16 + ty * 4 + tx;
    return n;
}

template <typename T> T *dup(const std::vector<T> &v) {
    T *p = (T *)malloc(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) void d1synth_default_params(D1SynthParams *p, int w, int h, int bitdepth_max, uint64_t seed) {
    memset(p, 0, sizeof(*p));
    p->w = w; p->h = h; p->ss_hor = 1; p->ss_ver = 1; p->bitdepth_max = bitdepth_max; p->seed = seed;
    p->p_intra = 0.3f; p->p_residual = 0.6f; p->p_tx_split = 0.5f;
    p->p_filter_intra = 0.05f; p->p_palette = 0.02f; p->p_cfl = 0.25f;
    p->p_avg = 0.2f; p->p_w_avg = 0.1f; p->p_wedge = 0.1f; p->p_seg = 0.05f; p->p_warp = 0.05f;
    p->mv_range = 128; p->n_refs = 2; p->edge_filter = 1; p->only_tx = -1; p->only_txtp = -1; p->eob_class = -1; p->dense_coefs = 0; p->p_obmc = 0.f; p->p_ii = 0.f; p->p_ibc = 0.f; p->tile_cols = 1; p->tile_rows = 1; p->real_blocks = 0;
    for (int i = 0; i < 7; i++) p->ref_w[i] = p->ref_h[i] = 0;
    p->mask_tab = 0; p->warp_tab = 0; p->n_warp_tab = 0; p->p_sub8x8 = 0.f;
}

__attribute__((visibility("default"))) int d1synth_generate(const D1SynthParams *p, D1SynthFrame *f) {
    if (!p || !f || (p->w & 7) || (p->h & 7) || p->w <= 0 || p->h <= 0) return -22;
    Gen g(*p);
    g.run();
    memset(f, 0, sizeof(*f));
    // compound: wave 0 then wave 1
    std::vector<Dav1dCudaMcDesc> comp(g.comp0);
    comp.insert(comp.end(), g.comp1.begin(), g.comp1.end());
    // itx: stable sort by class and, inside a class, by transform type (dc-only blocks first):
    // neighbouring warps then run the same 1-D transforms, which keeps the groups of a warp
    // convergent and the instruction working set of an SM small
    std::vector<uint32_t> itx_new(g.itx.size());
    std::vector<Dav1dCudaItxDesc> itx_sorted(g.itx.size());
    {
        for (auto &d : g.itx) f->itx_class_count[d.tx]++;
        std::vector<uint32_t> idx(g.itx.size());
        for (size_t i = 0; i < idx.size(); i++) idx[i] = (uint32_t)i;
        auto key = [&](uint32_t i) {
            const Dav1dCudaItxDesc &d = g.itx[i];
            return ((int)d.tx << 8) | (d.eob == 0 && d.txtp == 0 ? 0 : 1 + d.txtp);
        };
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return key(x) < key(y); });
        for (size_t n = 0; n < idx.size(); n++) {
            itx_new[idx[n]] = (uint32_t)n;
            itx_sorted[n] = g.itx[idx[n]];
        }
    }
    // the residuals of the intra-class operations as transform descriptors, ordered like `itx`
    {
        std::vector<Dav1dCudaItxDesc> v;
        for (auto &d : g.intra) {
            if (d.eob < 0 || d.mode == DAV1D_CUDA_INTRA_PAL) continue;
            Dav1dCudaItxDesc t;
            memset(&t, 0, sizeof(t));
            t.coef_off = d.coef_off; t.x = (uint16_t)(d.x4 * 4); t.y = (uint16_t)(d.y4 * 4);
            t.eob = d.eob; t.plane = d.plane; t.tx = d.tx; t.txtp = d.txtp; t.cw4 = d.cw4; t.ch4 = d.ch4;
            v.push_back(t);
            f->intra_itx_class_count[d.tx]++;
        }
        std::stable_sort(v.begin(), v.end(), [](const Dav1dCudaItxDesc &x, const Dav1dCudaItxDesc &y) {
            const int kx = ((int)x.tx << 8) | (x.eob == 0 && x.txtp == 0 ? 0 : 1 + x.txtp);
            const int ky = ((int)y.tx << 8) | (y.eob == 0 && y.txtp == 0 ? 0 : 1 + y.txtp);
            return kx < ky;
        });
        f->intra_itx = dup(v); f->n_intra_itx = (int32_t)v.size();
    }
    // tiles: per list (put, compound wave 0, compound wave 1) the tiles of blocks of at most
    // 8x8 samples first (they are processed four per warp), then the others
    std::vector<uint32_t> put_tiles, comp_tiles;
    uint32_t buf[16];
    auto emit = [&](const std::vector<Dav1dCudaMcDesc> &v, size_t lo, size_t hi, std::vector<uint32_t> &out,
                    int32_t *n_small) {
        const size_t base = out.size();
        for (int pass = 0; pass < 2; pass++)
            for (size_t i = lo; i < hi; i++) {
                const bool small = v[i].w <= 8 && v[i].h <= 8;
                if (small != (pass == 0)) continue;
                const int n = mc_tiles((uint32_t)i, v[i].w, v[i].h, buf);
                out.insert(out.end(), buf, buf + n);
                if (small) *n_small += n;
            }
        return (int32_t)(out.size() - base);
    };
    emit(g.put, 0, g.put.size(), put_tiles, &f->n_mc_put_small);
    f->n_mc_comp_tiles[0] = emit(comp, 0, g.comp0.size(), comp_tiles, &f->n_mc_comp_small[0]);
    f->n_mc_comp_tiles[1] = emit(comp, g.comp0.size(), comp.size(), comp_tiles, &f->n_mc_comp_small[1]);
    // OBMC: blend_h wave, then blend_v wave (one descriptor array, tiles per wave)
    std::vector<Dav1dCudaMcDesc> obmc(g.obmc_h);
    obmc.insert(obmc.end(), g.obmc_v.begin(), g.obmc_v.end());
    std::vector<uint32_t> obmc_tiles;
    for (int wave = 0; wave < 2; wave++) {
        const size_t lo = wave ? g.obmc_h.size() : 0, hi = wave ? obmc.size() : g.obmc_h.size();
        const size_t base = obmc_tiles.size();
        for (size_t i = lo; i < hi; i++) {
            const int n = mc_tiles((uint32_t)i, obmc[i].w, obmc[i].h, buf);
            obmc_tiles.insert(obmc_tiles.end(), buf, buf + n);
        }
        f->n_mc_obmc_tiles[wave] = (int32_t)(obmc_tiles.size() - base);
    }
    f->mc_obmc = dup(obmc); f->n_mc_obmc = (int32_t)obmc.size();
    f->mc_obmc_tiles = dup(obmc_tiles);
    std::vector<uint32_t> order(g.order.size());
    for (size_t i = 0; i < g.order.size(); i++) {
        uint32_t cls = g.order[i].cls, idx = g.order[i].idx;
        if (cls == 5) { cls = 1; idx += (uint32_t)g.comp0.size(); }
        if (cls == 7) { cls = 6; idx += (uint32_t)g.obmc_h.size(); }
        if (cls >= 8) { for (uint32_t k = 8; k < cls; k++) idx += (uint32_t)g.sc[k - 8].size(); cls = 8; }
        if (cls == 3) idx = itx_new[idx];
        order[i] = (cls << 28) | idx;
    }
    f->mc_put = dup(g.put); f->n_mc_put = (int32_t)g.put.size();
    f->mc_put_tiles = dup(put_tiles); f->n_mc_put_tiles = (int32_t)put_tiles.size();
    f->mc_comp = dup(comp); f->n_mc_comp = (int32_t)comp.size();
    f->mc_comp_tiles = dup(comp_tiles);
    f->warp = dup(g.warp); f->n_warp = (int32_t)g.warp.size();
    f->itx = dup(itx_sorted); f->n_itx = (int32_t)itx_sorted.size();
    f->intra = dup(g.intra); f->n_intra = (int32_t)g.intra.size();
    f->cf_elems = g.cf32.size();
    if (g.hbd) f->cf = dup(g.cf32);
    else {
        std::vector<int16_t> c16(g.cf32.begin(), g.cf32.end());
        f->cf = dup(c16);
    }
    f->masks = dup(g.masks); f->masks_bytes = g.masks.size();
    if (g.hbd) f->pal = dup(g.pal);
    else {
        std::vector<uint8_t> p8(g.pal.begin(), g.pal.end());
        f->pal = dup(p8);
    }
    f->pal_px = g.pal.size();
    f->pal_idx = dup(g.pal_idx); f->pal_idx_bytes = g.pal_idx.size();
    f->order = dup(order); f->n_order = (int32_t)order.size();
    f->bw4 = g.bw4; f->bh4 = g.bh4;
    f->algo_bytes = g.algo; f->luma_px = g.luma_px; f->dense_coef_bytes = g.dense_coef_bytes;
    for (int i = 0; i < 5; i++) f->algo_class[i] = g.algo_cls[i];
    f->n_blocks = g.n_blocks; f->n_intra_blocks = g.n_intra_blocks;
    f->blocks = dup(g.blocks); f->n_block_recs = (int32_t)g.blocks.size();
    f->tx_recs = dup(g.tx_recs); f->n_tx_recs = (int32_t)g.tx_recs.size();
    std::vector<Dav1dCudaMcScaledDesc> scaled;
    for (int k = 0; k < 4; k++) {
        scaled.insert(scaled.end(), g.sc[k].begin(), g.sc[k].end());
        f->n_mc_scaled[k] = (int32_t)g.sc[k].size();
    }
    f->mc_scaled = dup(scaled);
    return 0;
}

__attribute__((visibility("default"))) void d1synth_free(D1SynthFrame *f) {
    if (!f) return;
    free(f->mc_put); free(f->mc_put_tiles); free(f->mc_comp); free(f->mc_comp_tiles); free(f->warp);
    free(f->mc_obmc); free(f->mc_obmc_tiles); free(f->blocks); free(f->tx_recs); free(f->mc_scaled);
    free(f->intra_itx);
    free(f->itx); free(f->intra); free(f->cf); free(f->masks); free(f->pal); free(f->pal_idx); free(f->order);
    memset(f, 0, sizeof(*f));
}

}  // extern "C"
