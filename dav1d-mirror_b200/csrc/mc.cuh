// Motion-compensation device functions (one warp = one tile of at most
// 32x32 output samples).
//
// Reference being matched, bit for bit: src/mc_tmpl.c
//   put_8tap_c :113-171   prep_8tap_c :223-282   put/prep_bilin_c :395-546
//   avg/w_avg/mask :587-639   w_mask :683-742   blend* :642-681
//   warp_affine_8x8(t) :758-825   emu_edge :827-875   *_scaled :173-221,284-328,452-585
//
// Staging: the (tw+7)x(th+7) reference window is read ONCE from global memory
// with clamped coordinates (= emu_edge followed by a plain read) into shared
// memory; the horizontal pass writes an int16 `mid` tile, the vertical pass
// reads it - both as register-resident sliding windows (8 outputs per lane).
// Bilinear is run through the same machinery as the 2-tap filter
// {16-m, m} with base shift 4 instead of 6 (identical integer results).
#pragma once
#include "common.cuh"
#include "tables.cuh"
#include "ctx.h"

namespace d1 {

constexpr int MC_T = 32;                 // tile edge
constexpr int MC_ROWS = MC_T + 7;        // rows/cols of the staged window
constexpr int MC_MID_STRIDE = MC_T;      // int16
// Staged window row: the 16-byte aligned superset of the (tw+7) needed pixels,
// i.e. up to 39 + (vector width - 1) pixels: 48 u16 (6 vectors) / 64 u8 (4 vectors).
// TMAX = largest tile edge of the kernel variant: 32 (one tile per warp) or 8 (four
// tiles of at most 8x8 per warp, one per group of 8 lanes): rows of 15 + 7 px -> 24 u16 / 32 u8.
template <typename pixel, int TMAX = MC_T> struct McSrcGeo {
    static constexpr int VPX = 16 / (int)sizeof(pixel);      // pixels per 16-byte vector
    static constexpr int STRIDE = TMAX == 32 ? (sizeof(pixel) == 2 ? 48 : 64) : (sizeof(pixel) == 2 ? 24 : 32);
};

template <typename pixel, int TMAX = MC_T> struct __align__(16) McSmem {
    pixel src[(TMAX + 7) * McSrcGeo<pixel, TMAX>::STRIDE];
    int16_t mid[(TMAX + 7) * TMAX];
};
template <typename pixel, int TMAX = MC_T> struct __align__(16) McSmemCompound {
    McSmem<pixel, TMAX> s;
    int16_t ta[TMAX * TMAX];
    int16_t tb[TMAX * TMAX];
};

// 8 taps for one direction (warp-uniform). `dim` is the full block's width
// (horizontal) or height (vertical): dims <= 4 switch to the 4-tap sets
// (mc_tmpl.c:99-107).  Filter2d -> (h, v) filter type: levels.h:184-196,
// mc_tmpl.c:376-384.
DEV void mc_load_taps(int *f, const int filter_2d, const bool vertical, const int frac, const int dim) {
    if (filter_2d == 9) {
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = 0;
        f[3] = 16 - frac;
        f[4] = frac;
        return;
    }
    const int th = (0x15A80 >> (2 * filter_2d)) & 3;   // {0,0,0,2,2,2,1,1,1}
    const int tv = filter_2d % 3;                      // {0,1,2,0,1,2,0,1,2}
    const int t = vertical ? tv : th;
    const int set = dim > 4 ? t : 3 + (t & 1);
    const int8_t *p = g_subpel_filters + (set * 15 + frac - 1) * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) f[k] = p[k];
}

template <typename pixel, bool PREP> struct McOut;
template <typename pixel> struct McOut<pixel, false> {
    typedef pixel type;
    static DEV pixel fin(int sum, int rnd, int sh, int bdmax) {
        return (pixel)clip_px<pixel>((sum + rnd) >> sh, bdmax);
    }
};
template <typename pixel> struct McOut<pixel, true> {
    typedef int16_t type;
    static DEV int16_t fin(int sum, int rnd, int sh, int) {
        return (int16_t)(((sum + rnd) >> sh) - PxTraits<pixel>::prep_bias);
    }
};

// Horizontal pass over rows [r_lo, r_hi) of the staged window. SW outputs per lane.
// FIN = false: write int16 mid ((sum + rnd) >> sh); FIN = true: finish to `out`.
template <typename pixel, bool PREP, int SW, bool FIN, int TMAX, int G>
DEV void mc_hpass(const pixel *s_src, const int *fh, const int tw, const int r_lo, const int r_hi,
                  const int rnd, const int sh, const int bdmax, int16_t *s_mid,
                  typename McOut<pixel, PREP>::type *out, const int ostride, const int lane)
{
    const int nst = tw / SW;
    const int total = (r_hi - r_lo) * nst;
    for (int s = lane; s < total; s += G) {
        const int r = r_lo + s / nst, c0 = (s % nst) * SW;
        const pixel *p = s_src + r * McSrcGeo<pixel, TMAX>::STRIDE + c0;
        int v[SW + 7];
#pragma unroll
        for (int k = 0; k < SW + 7; k++) v[k] = p[k];
#pragma unroll
        for (int o = 0; o < SW; o++) {
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += fh[k] * v[o + k];
            if (FIN) out[(r - 3) * ostride + c0 + o] = McOut<pixel, PREP>::fin(sum, rnd, sh, bdmax);
            else s_mid[r * TMAX + c0 + o] = (int16_t)((sum + rnd) >> sh);
        }
    }
}

// Vertical pass: 8 output rows per lane, column x. SRC is the staged pixel
// window (column offset 3, stride McSrcGeo::STRIDE) or the int16 mid tile.
template <typename pixel, bool PREP, typename SRC, int G>
DEV void mc_vpass(const SRC *src, const int sstride, const int *fv, const int tw, const int th,
                  const int rnd, const int sh, const int bdmax,
                  typename McOut<pixel, PREP>::type *out, const int ostride, const int lane)
{
    const int nvs = (th + 7) >> 3;
    const int total = tw * nvs;
    for (int s = lane; s < total; s += G) {
        const int x = s % tw, y0 = (s / tw) * 8;
        const SRC *p = src + y0 * sstride + x;
        int v[15];
#pragma unroll
        for (int k = 0; k < 15; k++) v[k] = p[k * sstride];
#pragma unroll
        for (int o = 0; o < 8; o++) {
            if (y0 + o < th) {
                int sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) sum += fv[k] * v[o + k];
                out[(y0 + o) * ostride + x] = McOut<pixel, PREP>::fin(sum, rnd, sh, bdmax);
            }
        }
    }
}

// One tile of put (PREP = false) / prep (PREP = true).
//   ref       reference plane (clamped reads)
//   sx, sy    integer sample position of the tile's top-left in the reference
//   tw, th    tile size (<= 32); bw, bh: full block size (filter selection)
//   out       tile's top-left in the destination (pixels, or int16 for prep)
//   lane      lane index inside the group of G lanes that owns this tile (G = 32: the warp;
//             G = 8: four tiles per warp, `gmask` = the group's lanes for the barriers)
template <typename pixel, bool PREP, int TMAX = MC_T, int G = 32>
DEV void mc_tile(const PlaneView &ref, const int sx, const int sy, const int tw, const int th,
                 const int bw, const int bh, const int mx, const int my, const int filter_2d,
                 const int bdmax, McSmem<pixel, TMAX> *sm,
                 typename McOut<pixel, PREP>::type *out, const int ostride, const int lane,
                 const unsigned gmask = 0xffffffffu)
{
    typedef McOut<pixel, PREP> O;
    const int ib = PxTraits<pixel>::inter_bits(bdmax);
    const int bs = filter_2d == 9 ? 4 : 6;
    int fh[8], fv[8];
    if (mx) mc_load_taps(fh, filter_2d, false, mx, bw);
    if (my) mc_load_taps(fv, filter_2d, true, my, bh);

    // ---- stage the window: rows/cols -3..+4 only where a filter needs them.
    // Fast path (window columns inside the plane): 16-byte cp.async copies of
    // the aligned superset of each row, global -> shared without a register
    // round trip; `off` = position of window column 0 inside the staged row.
    // Row indices are clamped in both paths (top/bottom edge emulation); tiles
    // that cross the left/right picture edge take the per-pixel clamped path.
    constexpr int SS = McSrcGeo<pixel, TMAX>::STRIDE, VPX = McSrcGeo<pixel, TMAX>::VPX;
    const int c_lo = mx ? 0 : 3, c_hi = mx ? tw + 7 : tw + 3;
    const int r_lo = my ? 0 : 3, r_hi = my ? th + 7 : th + 3;
    int off = 0;
    {
        const pixel *rp = (const pixel *)ref.data;
        const int64_t rstride = ref.stride / (int64_t)sizeof(pixel);
        if (sx - 3 >= 0 && sx - 3 + c_hi <= ref.w) {
            const int a0 = (sx - 3) & ~(VPX - 1);
            off = (sx - 3) - a0;
            const int v_lo = (off + c_lo) / VPX, nv = (off + c_hi - 1) / VPX - v_lo + 1;
            const int total = (r_hi - r_lo) * nv;
            for (int i = lane; i < total; i += G) {
                const int r = r_lo + i / nv, v = v_lo + i % nv;
                const int yy = iclip(sy - 3 + r, 0, ref.h - 1);
                const pixel *g = rp + yy * rstride + a0 + v * VPX;
                const unsigned sa = (unsigned)__cvta_generic_to_shared(sm->src + r * SS + v * VPX);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(g) : "memory");
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        } else {
            const int ncols = c_hi - c_lo;
            const int lpr = ncols <= 8 || G == 8 ? 8 : ncols <= 16 ? 16 : 32;   // lanes per row
            const int rpi = G / lpr;                                   // rows per iteration
            const int lr = lane / lpr, lc = lane % lpr;
            for (int r = r_lo + lr; r < r_hi; r += rpi) {
                const int yy = iclip(sy - 3 + r, 0, ref.h - 1);
                const pixel *row = rp + yy * rstride;
                for (int c = c_lo + lc; c < c_hi; c += lpr) {
                    const int xx = iclip(sx - 3 + c, 0, ref.w - 1);
                    sm->src[r * SS + c] = row[xx];
                }
            }
        }
    }
    __syncwarp(gmask);
    const pixel *wsrc = sm->src + off;      // window column c lives at wsrc[r * SS + c]

    if (mx && my) {
        const int sh1 = bs - ib, rnd1 = (1 << sh1) >> 1;
        if (tw >= 8)      mc_hpass<pixel, PREP, 8, false, TMAX, G>(wsrc, fh, tw, 0, th + 7, rnd1, sh1, bdmax, sm->mid, nullptr, 0, lane);
        else if (tw == 4) mc_hpass<pixel, PREP, 4, false, TMAX, G>(wsrc, fh, tw, 0, th + 7, rnd1, sh1, bdmax, sm->mid, nullptr, 0, lane);
        else              mc_hpass<pixel, PREP, 2, false, TMAX, G>(wsrc, fh, tw, 0, th + 7, rnd1, sh1, bdmax, sm->mid, nullptr, 0, lane);
        __syncwarp(gmask);
        const int sh2 = PREP ? bs : bs + ib, rnd2 = (1 << sh2) >> 1;
        mc_vpass<pixel, PREP, int16_t, G>(sm->mid, TMAX, fv, tw, th, rnd2, sh2, bdmax, out, ostride, lane);
    } else if (mx) {
        const int sh = PREP ? bs - ib : bs;
        const int rnd = PREP ? (1 << sh) >> 1 : (1 << (bs - 1)) + ((1 << (bs - ib)) >> 1);
        if (tw >= 8)      mc_hpass<pixel, PREP, 8, true, TMAX, G>(wsrc, fh, tw, 3, th + 3, rnd, sh, bdmax, nullptr, out, ostride, lane);
        else if (tw == 4) mc_hpass<pixel, PREP, 4, true, TMAX, G>(wsrc, fh, tw, 3, th + 3, rnd, sh, bdmax, nullptr, out, ostride, lane);
        else              mc_hpass<pixel, PREP, 2, true, TMAX, G>(wsrc, fh, tw, 3, th + 3, rnd, sh, bdmax, nullptr, out, ostride, lane);
    } else if (my) {
        const int sh = PREP ? bs - ib : bs, rnd = (1 << sh) >> 1;
        mc_vpass<pixel, PREP, pixel, G>(wsrc + 3, SS, fv, tw, th, rnd, sh, bdmax, out, ostride, lane);
    } else {
        for (int i = lane; i < tw * th; i += G) {
            const int y = i / tw, x = i % tw;
            const int px = wsrc[(y + 3) * SS + x + 3];
            if (PREP) out[y * ostride + x] = (typename O::type)((px << ib) - PxTraits<pixel>::prep_bias);
            else out[y * ostride + x] = (typename O::type)px;
        }
    }
    __syncwarp(gmask);
}

// ------------------------------------------------------------ compound combine
// avg / w_avg / mask / w_mask over a w x h region, executed by `nthr` threads.
// t1/t2: int16 intermediates (stride ts).  mask: MASK input (stride ms) or
// W_MASK output (stride ms, already offset to the region's first entry).
// (mc_tmpl.c:587-639, 683-742)
template <typename pixel>
DEV void mc_combine(const int kind, const int16_t *t1, const int16_t *t2, const int ts,
                    pixel *dst, const int dstride, const int w, const int h,
                    const int weight_or_sign, uint8_t *mask, const int ms, const int mask_ss,
                    const int bdmax, const int tid, const int nthr)
{
    const int ib = PxTraits<pixel>::inter_bits(bdmax);
    const int bias = PxTraits<pixel>::prep_bias;
    if (kind == DAV1D_CUDA_MC_AVG) {
        const int sh = ib + 1, rnd = (1 << ib) + bias * 2;
        for (int i = tid; i < w * h; i += nthr) {
            const int y = i / w, x = i % w;
            dst[y * dstride + x] = (pixel)clip_px<pixel>((t1[y * ts + x] + t2[y * ts + x] + rnd) >> sh, bdmax);
        }
    } else if (kind == DAV1D_CUDA_MC_W_AVG) {
        const int sh = ib + 4, rnd = (8 << ib) + bias * 16, wt = weight_or_sign;
        for (int i = tid; i < w * h; i += nthr) {
            const int y = i / w, x = i % w;
            dst[y * dstride + x] = (pixel)clip_px<pixel>(
                (t1[y * ts + x] * wt + t2[y * ts + x] * (16 - wt) + rnd) >> sh, bdmax);
        }
    } else if (kind == DAV1D_CUDA_MC_MASK) {
        const int sh = ib + 6, rnd = (32 << ib) + bias * 64;
        for (int i = tid; i < w * h; i += nthr) {
            const int y = i / w, x = i % w;
            const int m = mask[y * ms + x];
            dst[y * dstride + x] = (pixel)clip_px<pixel>(
                (t1[y * ts + x] * m + t2[y * ts + x] * (64 - m) + rnd) >> sh, bdmax);
        }
    } else {  // W_MASK
        const int sh = ib + 6, rnd = (32 << ib) + bias * 64;
        const int bitdepth = PxTraits<pixel>::bitdepth(bdmax);
        const int mask_sh = bitdepth + ib - 4, mask_rnd = 1 << (mask_sh - 5);
        const int sign = weight_or_sign;
        const int ssh = mask_ss >= 1, ssv = mask_ss == 2;
        const int qw = w >> ssh, qh = h >> ssv;
        for (int i = tid; i < qw * qh; i += nthr) {
            const int qy = i / qw, qx = i % qw;
            int msum = 0;
            for (int dy = 0; dy <= ssv; dy++) {
                for (int dx = 0; dx <= ssh; dx++) {
                    const int y = (qy << ssv) + dy, x = (qx << ssh) + dx;
                    const int a = t1[y * ts + x], b = t2[y * ts + x];
                    const int m = imin(38 + ((iabs(a - b) + mask_rnd) >> mask_sh), 64);
                    dst[y * dstride + x] = (pixel)clip_px<pixel>((a * m + b * (64 - m) + rnd) >> sh, bdmax);
                    msum += m;
                }
            }
            // 444: m; 422: (m+n+1-sign)>>1; 420: (m+n+m'+n'+2-sign)>>2
            const int out = !ssh ? msum : !ssv ? (msum + 1 - sign) >> 1 : (msum + 2 - sign) >> 2;
            mask[qy * ms + qx] = (uint8_t)out;
        }
    }
}

// ------------------------------------------------------------ blend (OBMC, inter-intra)
// kind 0: blend (per-pixel mask, stride w); 1: blend_v (obmc_masks[w + x], first 3w/4
// columns); 2: blend_h (obmc_masks[h + y], first 3h/4 rows).  mc_tmpl.c:642-681
template <typename pixel>
DEV void mc_blend(const int kind, pixel *dst, const int dstride, const pixel *tmp, const int w, const int h,
                  const uint8_t *mask, const int tid, const int nthr)
{
    const int ww = kind == 1 ? (w * 3) >> 2 : w;
    const int hh = kind == 2 ? (h * 3) >> 2 : h;
    for (int i = tid; i < ww * hh; i += nthr) {
        const int y = i / ww, x = i % ww;
        const int m = kind == 0 ? mask[y * w + x] : kind == 1 ? g_obmc_masks[w + x] : g_obmc_masks[h + y];
        const int a = dst[y * dstride + x], b = tmp[y * w + x];
        dst[y * dstride + x] = (pixel)((a * (64 - m) + b * m + 32) >> 6);
    }
}

// ------------------------------------------------------------ warp 8x8 (one warp per block)
// src points at the block's (0,0) sample position inside a plane that is read
// with clamping; `mid` is 15*8 int16 of shared memory.  mc_tmpl.c:758-825
template <typename pixel, bool PREP>
DEV void mc_warp8x8(const PlaneView &ref, const int sx, const int sy, const int16_t *abcd,
                    const int mx0, const int my0, const int bdmax, int16_t *mid,
                    typename McOut<pixel, PREP>::type *out, const int ostride, const int lane)
{
    const int ib = PxTraits<pixel>::inter_bits(bdmax);
    const pixel *rp = (const pixel *)ref.data;
    const int64_t rstride = ref.stride / (int64_t)sizeof(pixel);
    const int a0 = abcd[0], a1 = abcd[1], a2 = abcd[2], a3 = abcd[3];
    {
        const int sh = 7 - ib, rnd = (1 << sh) >> 1;
        for (int i = lane; i < 15 * 8; i += 32) {
            const int y = i >> 3, x = i & 7;
            const int tmx = mx0 + y * a1 + x * a0;
            const int8_t *f = g_warp_filter + (64 + ((tmx + 512) >> 10)) * 8;
            const int yy = iclip(sy - 3 + y, 0, ref.h - 1);
            const pixel *row = rp + yy * rstride;
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += f[k] * row[iclip(sx + x - 3 + k, 0, ref.w - 1)];
            mid[i] = (int16_t)((sum + rnd) >> sh);
        }
    }
    __syncwarp();
    {
        const int sh = PREP ? 7 : 7 + ib, rnd = (1 << sh) >> 1;
        for (int i = lane; i < 64; i += 32) {
            const int y = i >> 3, x = i & 7;
            const int tmy = my0 + y * a3 + x * a2;
            const int8_t *f = g_warp_filter + (64 + ((tmy + 512) >> 10)) * 8;
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += f[k] * mid[(y + k) * 8 + x];
            out[y * ostride + x] = McOut<pixel, PREP>::fin(sum, rnd, sh, bdmax);
        }
    }
    __syncwarp();
}

}  // namespace d1
