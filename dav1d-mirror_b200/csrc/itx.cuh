// 2-D inverse transform + add for one transform block, executed by a group
// of G consecutive lanes of a warp (G = 4..32, several blocks per warp for the
// small sizes).  Row pass: one lane per coefficient row, the whole W-point
// 1-D transform in registers; transpose through a per-group shared-memory
// tile; column pass: one lane per column.  Follows inv_txfm_add_c
// (reference src/itx_tmpl.c:40-100) step for step, including the dc-only
// shortcut (:53-65), rect2 pre-scaling (:80-82), the per-size inter-pass
// shift (:142-160), the bit-depth dependent clamps (:68-76) and WHT 4x4
// (:166-185).
#pragma once
#include "itx_geom.cuh"

namespace d1 {

template <int W, int H> struct ItxGeom {
    static constexpr int SW = W < 32 ? W : 32;
    static constexpr int SH = H < 32 ? H : 32;
    static constexpr int GMIN = SH > SW ? SH : SW;      // lanes that have work
    static constexpr int TSTRIDE = W + 1;               // odd stride: conflict-free transpose
    static constexpr int TILE_INTS = SH * TSTRIDE;
    static constexpr int SHIFT =
        (W == 4 && H == 4) || (W == 4 && H == 8) || (W == 8 && H == 4) ? 0 :
        (W == 8 && H == 32) || (W == 32 && H == 8) || (W == 16 && H == 16) || (W == 32 && H == 32) ||
        (W == 16 && H == 64) || (W == 64 && H == 16) || (W == 64 && H == 64) ? 2 : 1;
    static constexpr bool RECT2 = (W * 2 == H) || (H * 2 == W);
};

// The two passes are NOT inlined and depend only on the transform length (5
// variants each) instead of on the 19 (W, H) pairs: the fused intra kernel
// would otherwise carry ~1.5 MB of SASS and thrash the instruction cache.
//
// Row pass: lane gl (< SH, G >= SH) owns coefficient row gl (coeff[gl + x*SH],
// column-major).  While loading, each lane records which coefficients are
// non-zero; OR-reductions inside the group give the bounding box of the
// non-zero coefficients, which selects reduced 1-D transforms (inputs beyond
// the box are literal zeros).  Writes rows [0, rows_used) of the tile and
// returns rows_used (8, 16 or SH).
// (inlined in the per-size / task kernels of itx.cu, shared out of line in the fused intra
// kernel of recon.cu, which defines D1_ITX_PASS_NOINLINE)
#if defined(D1_ITX_PASS_NOINLINE)
#define D1_ITX_PASS __device__ __noinline__
#else
#define D1_ITX_PASS __device__ __forceinline__
#endif
template <typename coef, int W>
D1_ITX_PASS int itx_row_pass(const bool full, const int gl, const int G, coef *cf, const int SH,
                                         const bool rect2, const int shift, const int rk, const Clamp rowcl,
                                         const Clamp colcl, int *tile, const bool zero_coefs,
                                         const int cw, const int ch, const bool box)
{
    constexpr int SW = W < 32 ? W : 32, TS = W + 1;
    const bool wht = W == 4 && rk == K_WHT;      // WHT exists for 4x4 only
    const int rnd = (1 << shift) >> 1;
    int c[W];
    unsigned rowmask = 0;
    // stored coefficients: cw columns x ch rows, stride ch (dense: cw = SW, ch = SH)
    if (full && gl < ch) {
        if (box) {
            // packed descriptor: the non-zero box is given, no need to look at the values
#pragma unroll
            for (int x = 0; x < SW; x++) {
                int v = x < cw ? (int)cf[gl + x * ch] : 0;
                if (wht) v >>= 2;
                else if (rect2) v = (v * 181 + 128) >> 8;
                c[x] = v;
            }
            rowmask = cw >= 32 ? ~0u : (1u << cw) - 1u;
        } else {
#pragma unroll
            for (int x = 0; x < SW; x++) {
                int v = x < cw ? (int)cf[gl + x * ch] : 0;
                rowmask |= (unsigned)(v != 0) << x;
                if (wht) v >>= 2;
                else if (rect2) v = (v * 181 + 128) >> 8;
                c[x] = v;
            }
        }
#pragma unroll
        for (int x = SW; x < W; x++) c[x] = 0;
        if (zero_coefs) {
#pragma unroll
            for (int x = 0; x < SW; x++)
                if (x < cw) cf[gl + x * ch] = 0;
        }
    }
    unsigned colbits = rowmask, rowbits = rowmask ? 1u << gl : 0u;
    for (int o = G >> 1; o > 0; o >>= 1) {
        colbits |= __shfl_xor_sync(0xffffffffu, colbits, o);
        rowbits |= __shfl_xor_sync(0xffffffffu, rowbits, o);
    }
    const int nzw = 32 - __clz(colbits), nzh = 32 - __clz(rowbits);    // 0 if the block is all zero
    const int rows_used = SH <= 8 ? SH : nzh <= 8 ? 8 : (SH <= 16 || nzh <= 16) ? (SH < 16 ? SH : 16) : SH;
    if (full && gl < rows_used) {
        int *trow = tile + gl * TS;
        if (rowmask == 0) {
            // an all-zero row transforms to zeros (every stage is v*c+rnd>>s and clamp)
#pragma unroll
            for (int x = 0; x < W; x++) trow[x] = 0;
        } else {
            itx1d_dispatch<W>(c, rk, rowcl, nzw);
            if (wht) {
#pragma unroll
                for (int x = 0; x < W; x++) trow[x] = c[x];
            } else {
#pragma unroll
                for (int x = 0; x < W; x++) trow[x] = colcl((c[x] + rnd) >> shift);
            }
        }
    }
    return rows_used;
}

// Column pass: lane gl owns columns gl, gl+G, ...; only rows < rows_used of the
// tile are non-zero.  Read-modify-write of the destination column in chunks of
// CH rows: the loads of the first chunk are issued BEFORE the 1-D transform and
// those of chunk k+1 before the stores of chunk k, so the global-memory latency
// overlaps the arithmetic.  __ldcg: L2-coherent reads (the intra dataflow
// kernel consumes pixels written by other SMs in the same launch).
template <typename pixel, int H>
D1_ITX_PASS void itx_col_pass(const bool full, const int gl, const int G, const int *tile, const int TS,
                                          const int W, const int rows_used, const int ck, const Clamp colcl,
                                          pixel *dst, const int dstride, const int bdmax)
{
    constexpr int SH = H < 32 ? H : 32;
    const bool wht = H == 4 && ck == K_WHT;      // WHT exists for 4x4 only
    if (!full) return;
    for (int x = gl; x < W; x += G) {
        pixel *p = dst + x;
        constexpr int CH = H < 16 ? H : 16;
        int c[H], dpx[CH];
#pragma unroll
        for (int y = 0; y < CH; y++) dpx[y] = __ldcg(p + y * dstride);
        if (SH > 16 && rows_used > 16) {
#pragma unroll
            for (int y = 0; y < SH; y++) c[y] = tile[y * TS + x];
        } else if (SH > 8 && rows_used > 8) {
#pragma unroll
            for (int y = 0; y < (SH < 16 ? SH : 16); y++) c[y] = tile[y * TS + x];
#pragma unroll
            for (int y = 16; y < SH; y++) c[y] = 0;
        } else {
#pragma unroll
            for (int y = 0; y < (SH < 8 ? SH : 8); y++) c[y] = tile[y * TS + x];
#pragma unroll
            for (int y = 8; y < SH; y++) c[y] = 0;
        }
#pragma unroll
        for (int y = SH; y < H; y++) c[y] = 0;
        itx1d_dispatch<H>(c, ck, colcl, rows_used);
#pragma unroll
        for (int y0 = 0; y0 < H; y0 += CH) {
            int cur[CH];
#pragma unroll
            for (int y = 0; y < CH; y++) cur[y] = dpx[y];
            if (y0 + CH < H) {
#pragma unroll
                for (int y = 0; y < CH; y++) dpx[y] = __ldcg(p + (y0 + CH + y) * dstride);
            }
#pragma unroll
            for (int y = 0; y < CH; y++) {
                const int r = wht ? c[y0 + y] : (c[y0 + y] + 8) >> 4;
                p[(y0 + y) * dstride] = (pixel)clip_px<pixel>(cur[y] + r, bdmax);
            }
        }
    }
}

// `gl` = lane index inside the group, `G` = group size (power of two >= 4),
// `tile` = this group's shared scratch (ItxGeom::TILE_INTS ints).
// `dst`/`dstride` (in pixels) point to global memory.
// All lanes of the warp must call this (it contains __syncwarp()); lanes of a
// group without a block pass active = false.
template <typename pixel, int W, int H, int G>
DEV void itx_block(const bool active, const int gl, int *tile,
                   typename PxTraits<pixel>::coef *cf, const int eob, const int txtp,
                   pixel *dst, const int dstride, const int bdmax, const bool zero_coefs,
                   const int cw4 = 0, const int ch4 = 0)
{
    typedef ItxGeom<W, H> Geo;
    typedef typename PxTraits<pixel>::coef coef;
    constexpr int SH = Geo::SH, TS = Geo::TSTRIDE, SHIFT = Geo::SHIFT;
    constexpr int RND = (1 << SHIFT) >> 1;

    const bool dc_only = active && eob == 0 && txtp == 0;    // has_dconly: DCT_DCT only

    if (dc_only) {
        int dc = 0;
        dc = cf[0];
        if (Geo::RECT2) dc = (dc * 181 + 128) >> 8;
        dc = (dc * 181 + 128) >> 8;
        dc = (dc + RND) >> SHIFT;
        dc = (dc * 181 + 128 + 2048) >> 12;
        // vectorised read-modify-write, 4 vectors in flight per lane
        constexpr int VW = W < 8 ? 4 : 8, SEGS = W / VW, NV = SEGS * H;
        for (int i0 = gl; i0 < NV; i0 += 4 * G) {
            int v[4][VW];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * G;
                if (i < NV) load_px<pixel, VW>(dst + (i / SEGS) * dstride + (i % SEGS) * VW, v[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * G;
                if (i < NV) {
#pragma unroll
                    for (int k = 0; k < VW; k++) v[u][k] = clip_px<pixel>(v[u][k] + dc, bdmax);
                    store_px<pixel, VW>(dst + (i / SEGS) * dstride + (i % SEGS) * VW, v[u]);
                }
            }
        }
    }
    // make sure every lane of the group has read cf[0] before it is cleared
    __syncwarp();
    if (dc_only && gl == 0 && zero_coefs) cf[0] = 0;

    const bool full = active && !dc_only;
    Clamp rowcl, colcl;
    if (PxTraits<pixel>::hbd) {
        rowcl.lo = (int)((unsigned)~bdmax << 7);
        colcl.lo = (int)((unsigned)~bdmax << 5);
    } else {
        rowcl.lo = colcl.lo = -32768;
    }
    rowcl.hi = ~rowcl.lo;
    colcl.hi = ~colcl.lo;
    const int rk = txtp_row_kind(txtp), ck = txtp_col_kind(txtp);

    const int rows_used = itx_row_pass<coef, W>(full, gl, G, cf, SH, Geo::RECT2, SHIFT, rk, rowcl, colcl, tile,
                                                zero_coefs, cw4 ? cw4 * 4 : Geo::SW, ch4 ? ch4 * 4 : SH, cw4 != 0);
    __syncwarp();
    itx_col_pass<pixel, H>(full, gl, G, tile, TS, W, rows_used, ck, colcl, dst, dstride, bdmax);
    __syncwarp();
}

}  // namespace d1
