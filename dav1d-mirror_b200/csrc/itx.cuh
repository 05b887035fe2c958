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
#include "itx_1d.cuh"

namespace d1 {

struct TxDim { uint8_t w, h, shift; };

// enum RectTxfmSize -> (w, h, inter-pass shift), itx_tmpl.c:142-160 + levels.h:44-78
__host__ __device__ inline TxDim tx_dim(int tx) {
    const TxDim t[19] = {
        { 4, 4, 0 }, { 8, 8, 1 }, { 16, 16, 2 }, { 32, 32, 2 }, { 64, 64, 2 },
        { 4, 8, 0 }, { 8, 4, 0 }, { 8, 16, 1 }, { 16, 8, 1 }, { 16, 32, 1 },
        { 32, 16, 1 }, { 32, 64, 1 }, { 64, 32, 1 }, { 4, 16, 1 }, { 16, 4, 1 },
        { 8, 32, 2 }, { 32, 8, 2 }, { 16, 64, 2 }, { 64, 16, 2 } };
    return t[tx];
}

// enum TxfmType is named VERT_HORZ (levels.h:80-100); the row (first) pass is
// the horizontal transform (itx_tmpl.c:212-245).
HD int txtp_row_kind(int txtp) {
    // D=0 A=1 F=2 I=3 W=4, 3 bits each
    const uint64_t rows = 0ull | (0ull << 0) | (0ull << 3) | (1ull << 6) | (1ull << 9) | (0ull << 12) |
                          (2ull << 15) | (2ull << 18) | (2ull << 21) | (1ull << 24) | (3ull << 27) |
                          (3ull << 30) | (0ull << 33) | (3ull << 36) | (1ull << 39) | (3ull << 42) |
                          (2ull << 45) | (4ull << 48);
    return (int)((rows >> (3 * txtp)) & 7);
}
HD int txtp_col_kind(int txtp) {
    const uint64_t cols = 0ull | (0ull << 0) | (1ull << 3) | (0ull << 6) | (1ull << 9) | (2ull << 12) |
                          (0ull << 15) | (2ull << 18) | (1ull << 21) | (2ull << 24) | (3ull << 27) |
                          (0ull << 30) | (3ull << 33) | (1ull << 36) | (3ull << 39) | (2ull << 42) |
                          (3ull << 45) | (4ull << 48);
    return (int)((cols >> (3 * txtp)) & 7);
}

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

// `gl` = lane index inside the group, `G` = group size (power of two >= 4),
// `tile` = this group's shared scratch (ItxGeom::TILE_INTS ints).
// `dst`/`dstride` (in pixels) may point to global or shared memory.
// All lanes of the warp must call this (it contains __syncwarp()); lanes of a
// group without a block pass active = false.
template <typename pixel, int W, int H, int G>
DEV void itx_block(const bool active, const int gl, int *tile,
                   typename PxTraits<pixel>::coef *cf, const int eob, const int txtp,
                   pixel *dst, const int dstride, const int bdmax, const bool zero_coefs)
{
    typedef ItxGeom<W, H> Geo;
    constexpr int SW = Geo::SW, SH = Geo::SH, TS = Geo::TSTRIDE, SHIFT = Geo::SHIFT;
    constexpr int RND = (1 << SHIFT) >> 1;
    constexpr int NMAX = W > H ? W : H;

    const bool dc_only = active && eob == 0 && txtp == 0;    // has_dconly: DCT_DCT only
    const bool wht = (W == 4 && H == 4) && txtp == 16;

    if (dc_only) {
        int dc = 0;
        dc = cf[0];
        if (Geo::RECT2) dc = (dc * 181 + 128) >> 8;
        dc = (dc * 181 + 128) >> 8;
        dc = (dc + RND) >> SHIFT;
        dc = (dc * 181 + 128 + 2048) >> 12;
        for (int i = gl; i < W * H; i += G) {
            const int y = i / W, x = i % W;
            pixel *p = dst + y * dstride + x;
            *p = (pixel)clip_px<pixel>(*p + dc, bdmax);
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

    // ---- row pass: lane y owns coefficient row y (coeff[y + x*SH], column-major)
    if (full) {
        for (int y = gl; y < SH; y += G) {
            int c[NMAX];
            int nz = 0;
#pragma unroll
            for (int x = 0; x < SW; x++) {
                int v = cf[y + x * SH];
                nz |= v;
                if (wht) v >>= 2;
                else if (Geo::RECT2) v = (v * 181 + 128) >> 8;
                c[x] = v;
            }
#pragma unroll
            for (int x = SW; x < NMAX; x++) c[x] = 0;
            if (zero_coefs) {
#pragma unroll
                for (int x = 0; x < SW; x++) cf[y + x * SH] = 0;
            }
            int *trow = tile + y * TS;
            if (nz == 0) {
                // an all-zero row transforms to zeros (every stage is v*c+rnd>>s and clamp)
#pragma unroll
                for (int x = 0; x < W; x++) trow[x] = 0;
            } else {
                itx1d_run<W>(c, rk, rowcl);
                if (wht) {
#pragma unroll
                    for (int x = 0; x < W; x++) trow[x] = c[x];
                } else {
#pragma unroll
                    for (int x = 0; x < W; x++) trow[x] = colcl((c[x] + RND) >> SHIFT);
                }
            }
        }
    }
    __syncwarp();
    // ---- column pass: lane x owns column x
    if (full) {
        for (int x = gl; x < W; x += G) {
            int c[NMAX];
#pragma unroll
            for (int y = 0; y < SH; y++) c[y] = tile[y * TS + x];
#pragma unroll
            for (int y = SH; y < NMAX; y++) c[y] = 0;
            itx1d_run<H>(c, ck, colcl);
            pixel *p = dst + x;
            if (wht) {
#pragma unroll
                for (int y = 0; y < H; y++)
                    p[y * dstride] = (pixel)clip_px<pixel>(p[y * dstride] + c[y], bdmax);
            } else {
#pragma unroll
                for (int y = 0; y < H; y++)
                    p[y * dstride] = (pixel)clip_px<pixel>(p[y * dstride] + ((c[y] + 8) >> 4), bdmax);
            }
        }
    }
    __syncwarp();
}

}  // namespace d1
