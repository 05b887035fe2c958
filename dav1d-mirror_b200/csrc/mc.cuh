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
#include <type_traits>
#include "common.cuh"
#include "tables.cuh"
#include "ctx.h"

namespace d1 {

constexpr int MC_T = 32;                 // tile edge
constexpr int MC_ROWS = MC_T + 7;        // rows/cols of the staged window
// Staged window row: the 16-byte aligned superset of the (tw+7) needed pixels,
// i.e. up to 39 + (vector width - 1) pixels: 48 u16 (6 vectors) / 64 u8 (4 vectors).
// TMAX = largest tile edge of the kernel variant: 32 (one tile per warp) or 8 (four
// tiles of at most 8x8 per warp, one per group of 8 lanes): rows of 15 + 7 px -> 24 u16 / 32 u8.
// MID = row stride (int16) of the horizontal pass's output: 36 = 18 words for the 32-wide tile,
// so that the two 8-row blocks a warp reads in the vertical pass fall into disjoint banks.
template <typename pixel, int TMAX = MC_T> struct McSrcGeo {
    static constexpr int VPX = 16 / (int)sizeof(pixel);      // pixels per 16-byte vector
    static constexpr int STRIDE = TMAX == 32 ? (sizeof(pixel) == 2 ? 48 : 64) : (sizeof(pixel) == 2 ? 24 : 32);
    static constexpr int MID = TMAX == 32 ? 36 : 8;
};

template <typename pixel, int TMAX = MC_T> struct __align__(16) McSmem {
    pixel src[(TMAX + 7) * McSrcGeo<pixel, TMAX>::STRIDE];
    int16_t mid[(TMAX + 7) * McSrcGeo<pixel, TMAX>::MID];
};

// ---- TMA staging (32x32-tile kernels, pictures from dav1d_cuda_picture_alloc): the whole window of a tile
// is ONE cp.async.bulk.tensor.2d request against the reference plane's tensor map (box = STRIDE pixels x
// 15 / 23 / 39 rows, chosen by the tile height), completing on an mbarrier in the warp's shared memory.
// Two window buffers per warp: the request of the NEXT window is in flight while the current one is
// filtered.  Windows that touch the picture border keep the clamped cp.async / per-pixel path (TMA fills
// out-of-bounds elements with zeros, edge emulation replicates).
DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
constexpr int MC_TMA_CLASSES = 3;        // tensor maps per plane: box rows 15, 23, 39
DEV int mc_tma_class(const int th) { return th <= 8 ? 0 : th <= 16 ? 1 : 2; }
DEV int mc_tma_rows(const int cls) { return cls == 0 ? 15 : cls == 1 ? 23 : 39; }
template <typename pixel> struct __align__(128) McSmemTma {
    pixel src[2][(MC_T + 8) * McSrcGeo<pixel, MC_T>::STRIDE];      // 3840 / 2560 bytes each: 128-byte multiples
    int16_t mid[(MC_T + 7) * McSrcGeo<pixel, MC_T>::MID + 4];
    unsigned long long bar[2];
};
template <typename pixel> struct __align__(128) McSmemTmaCompound {
    McSmemTma<pixel> s;
    int16_t ta[MC_T * MC_T];
    int16_t tb[MC_T * MC_T];
};
template <typename pixel>
DEV void mc_tma_init(McSmemTma<pixel> *sm, const int lane) {
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm->bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm->bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
}
// one lane: expect the box's bytes, then the tensor request (x, y = element coordinates of the window's
// top-left in the plane)
DEV void mc_tma_issue(const unsigned dst, const void *tmap, const int x, const int y, const unsigned bytes,
                      const unsigned bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic accesses to the buffer are done
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
DEV void mbar_wait(const unsigned bar, const unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
template <typename pixel, int TMAX = MC_T> struct __align__(16) McSmemCompound {
    McSmem<pixel, TMAX> s;
    int16_t ta[TMAX * TMAX];
    int16_t tb[TMAX * TMAX];
};

// 8 taps for one direction (group-uniform), packed as bytes: x = (f0, f1, f2, f3),
// y = (f4, f5, f6, f7) - the operand format of dp2a.  `dim` is the full block's width
// (horizontal) or height (vertical): dims <= 4 switch to the 4-tap sets
// (mc_tmpl.c:99-107).  Filter2d -> (h, v) filter type: levels.h:184-196,
// mc_tmpl.c:376-384.
DEV uint2 mc_load_taps(const int filter_2d, const bool vertical, const int frac, const int dim) {
    if (filter_2d == 9) return make_uint2((unsigned)(16 - frac) << 24, (unsigned)frac);
    const int th = (0x15A80 >> (2 * filter_2d)) & 3;   // {0,0,0,2,2,2,1,1,1}
    const int tv = filter_2d % 3;                      // {0,1,2,0,1,2,0,1,2}
    const int t = vertical ? tv : th;
    const int set = dim > 4 ? t : 3 + (t & 1);
    return *(const uint2 *)(g_subpel_filters + (set * 15 + frac - 1) * 8);
}
DEV void mc_unpack_taps(const uint2 t, int *f) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        f[k] = (int)(int8_t)(t.x >> (8 * k));
        f[4 + k] = (int)(int8_t)(t.y >> (8 * k));
    }
}

// acc + a.lo16 * b.byte0 + a.hi16 * b.byte1 (lo) / ... b.byte2, b.byte3 (hi); a = two unsigned
// (pixels) or signed (intermediates) 16-bit values, b = signed taps.  Full rate on sm_100a
// (IDP.2A, profiles/ubench), i.e. two multiply-adds per issue slot instead of IMAD's one.
template <bool SIGNED> DEV int dp2a_lo(const unsigned a, const unsigned b, int c) {
    if (SIGNED) asm("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(c) : "r"(a), "r"(b));
    else asm("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(c) : "r"(a), "r"(b));
    return c;
}
template <bool SIGNED> DEV int dp2a_hi(const unsigned a, const unsigned b, int c) {
    if (SIGNED) asm("dp2a.hi.s32.s32 %0, %1, %2, %0;" : "+r"(c) : "r"(a), "r"(b));
    else asm("dp2a.hi.u32.s32 %0, %1, %2, %0;" : "+r"(c) : "r"(a), "r"(b));
    return c;
}

template <typename pixel, bool PREP> struct McOut;
template <typename pixel> struct McOut<pixel, false> {
    typedef pixel type;
    static DEV pixel fin(int sum, int rnd, int sh, int bdmax) {
        return (pixel)clip_px<pixel>((sum + rnd) >> sh, bdmax);
    }
    static DEV int fin2(int sum, int sh, int bdmax) { return min(max(sum >> sh, 0), bdmax); }
};
template <typename pixel> struct McOut<pixel, true> {
    typedef int16_t type;
    static DEV int16_t fin(int sum, int rnd, int sh, int) {
        return (int16_t)(((sum + rnd) >> sh) - PxTraits<pixel>::prep_bias);
    }
    static DEV int fin2(int sum, int sh, int) { return (sum >> sh) - PxTraits<pixel>::prep_bias; }
};
// two finished values -> one store (32-bit for 16-bit outputs, 16-bit for 8-bit pixels)
template <typename T> DEV void mc_store2(T *p, const int a, const int b) {
    if (sizeof(T) == 2) *(unsigned *)p = __byte_perm((unsigned)a, (unsigned)b, 0x5410);
    else *(uint16_t *)p = (uint16_t)((a & 0xff) | (b << 8));
}

// Horizontal pass over rows [r_lo, r_hi) of the staged window. SW outputs per lane.
// FIN = false: write int16 mid ((sum + rnd) >> sh); FIN = true: finish to `out`.
// (scalar variant: 8-bit pixels and tiles narrower than 8)
template <typename pixel, bool PREP, int SW, bool FIN, int TMAX, int G>
DEV void mc_hpass(const pixel *s_src, const int *fh, const int tw, const int r_lo, const int r_hi,
                  const int rnd, const int sh, const int bdmax, int16_t *s_mid,
                  typename McOut<pixel, PREP>::type *out, const int ostride, const int lane)
{
    const int ls = SW == 8 ? (tw >> 4) : 0;   // log2(tw / SW): SW = 8: tw 8/16/32; else tw == SW
    const int nst = 1 << ls;
    const int total = (r_hi - r_lo) << ls;
    for (int s = lane; s < total; s += G) {
        const int r = r_lo + (s >> ls), c0 = (s & (nst - 1)) * SW;
        const pixel *p = s_src + r * McSrcGeo<pixel, TMAX>::STRIDE + c0;
        int v[SW + 7];
#pragma unroll
        for (int k = 0; k < SW + 7; k++) v[k] = p[k];
#pragma unroll
        for (int o = 0; o < SW; o++) {
            int sum = rnd;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += fh[k] * v[o + k];
            if (FIN) out[(r - 3) * ostride + c0 + o] = (typename McOut<pixel, PREP>::type)McOut<pixel, PREP>::fin2(sum, sh, bdmax);
            else s_mid[r * McSrcGeo<pixel, TMAX>::MID + c0 + o] = (int16_t)(sum >> sh);
        }
    }
}

// Horizontal pass for 16-bit pixels and tiles at least 8 wide: a lane owns 8 adjacent outputs
// of one row.  It reads the 9 words (18 pixels) that start at the even pixel at or before its
// first tap, funnel-shifts them by the window's parity so that word i = pixels (2i, 2i+1), and
// forms every output from pixel PAIRS: even outputs use the taps (f0,f1)(f2,f3)(f4,f5)(f6,f7),
// odd outputs the same words with (0,f0)(f1,f2)(f3,f4)(f5,f6)(f7,0) - 4.5 dp2a per output
// instead of 8 IMAD, no per-output repacking.
template <typename pixel, bool PREP, bool FIN, int TMAX, int G>
DEV void mc_hpass_dp(const pixel *s_src, const int off, const uint2 t, const int tw, const int r_lo,
                     const int r_hi, const int rnd, const int sh, const int bdmax, int16_t *s_mid,
                     typename McOut<pixel, PREP>::type *out, const int ostride, const bool vec_ok, const int lane)
{
    static_assert(sizeof(pixel) == 2, "16-bit pixels only");
    constexpr int SS = McSrcGeo<pixel, TMAX>::STRIDE, MS = McSrcGeo<pixel, TMAX>::MID;
    const unsigned ta = t.x, tb = t.y;
    const unsigned tc = t.x << 8, td = __funnelshift_r(t.x, t.y, 24), te = t.y >> 24;
    const int ls = tw >> 4, nst = 1 << ls;         // tw = 8, 16, 32 -> 1, 2, 4 segments per row
    const int total = (r_hi - r_lo) << ls;
    const int p16 = (off & 1) * 16;
    for (int s = lane; s < total; s += G) {
        const int r = r_lo + (s >> ls), c0 = (s & (nst - 1)) * 8;
        const unsigned *wp = (const unsigned *)(s_src + r * SS) + ((off + c0) >> 1);
        unsigned w[9], a[8];
#pragma unroll
        for (int k = 0; k < 9; k++) w[k] = wp[k];
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = __funnelshift_r(w[k], w[k + 1], p16);
        int v[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int e = rnd, o = rnd;
            e = dp2a_lo<false>(a[i], ta, e);
            e = dp2a_hi<false>(a[i + 1], ta, e);
            e = dp2a_lo<false>(a[i + 2], tb, e);
            e = dp2a_hi<false>(a[i + 3], tb, e);
            o = dp2a_lo<false>(a[i], tc, o);
            o = dp2a_hi<false>(a[i + 1], tc, o);
            o = dp2a_lo<false>(a[i + 2], td, o);
            o = dp2a_hi<false>(a[i + 3], td, o);
            o = dp2a_lo<false>(a[i + 4], te, o);
            v[2 * i] = FIN ? McOut<pixel, PREP>::fin2(e, sh, bdmax) : e >> sh;
            v[2 * i + 1] = FIN ? McOut<pixel, PREP>::fin2(o, sh, bdmax) : o >> sh;
        }
        unsigned pk[4];
#pragma unroll
        for (int i = 0; i < 4; i++) pk[i] = __byte_perm((unsigned)v[2 * i], (unsigned)v[2 * i + 1], 0x5410);
        if (FIN) {
            typename McOut<pixel, PREP>::type *q = out + (r - 3) * ostride + c0;
            if (vec_ok) {
                *(uint2 *)q = make_uint2(pk[0], pk[1]);
                *(uint2 *)(q + 4) = make_uint2(pk[2], pk[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) q[i] = (typename McOut<pixel, PREP>::type)v[i];
            }
        } else {
            int16_t *q = s_mid + r * MS + c0;
            if (MS % 8 == 0) {
                *(uint4 *)q = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            } else {
                *(uint2 *)q = make_uint2(pk[0], pk[1]);
                *(uint2 *)(q + 4) = make_uint2(pk[2], pk[3]);
            }
        }
    }
}

// Vertical pass: 8 output rows per lane.  SRC is the staged pixel window (column offset 3,
// stride McSrcGeo::STRIDE) or the int16 mid tile.  CPL = columns per lane: with 2 the lane
// reads 15 words (two adjacent int16 columns), splits them into vertical pairs with PRMT and
// writes its outputs two at a time; with 1 it reads 15 scalars.  Either way an output is
// 4 dp2a on (row, row + 1) pairs.
template <typename pixel, bool PREP, typename SRC, int CPL, int G>
DEV void mc_vpass_dp(const SRC *src, const int sstride, const uint2 t, const int tw, const int th,
                     const int rnd, const int sh, const int bdmax,
                     typename McOut<pixel, PREP>::type *out, const int ostride, const int lane)
{
    constexpr bool SG = std::is_same<SRC, int16_t>::value;    // signed source
    typedef McOut<pixel, PREP> O;
    const unsigned ta = t.x, tb = t.y;
    const int ncol = tw / CPL;                     // power of two
    const int lc = 31 - __clz(ncol);
    const int total = ((th + 7) >> 3) << lc;
    for (int s = lane; s < total; s += G) {
        const int xc = s & (ncol - 1), y0 = (s >> lc) * 8;
        unsigned pa[14], pb[CPL == 2 ? 14 : 1];
        if (CPL == 2) {
            const unsigned *p = (const unsigned *)(src + y0 * sstride) + xc;
            unsigned w[15];
#pragma unroll
            for (int k = 0; k < 15; k++) w[k] = p[k * (sstride >> 1)];
#pragma unroll
            for (int k = 0; k < 14; k++) {
                pa[k] = __byte_perm(w[k], w[k + 1], 0x5410);
                pb[k] = __byte_perm(w[k], w[k + 1], 0x7632);
            }
        } else {
            const SRC *p = src + y0 * sstride + xc;
            unsigned w[15];
#pragma unroll
            for (int k = 0; k < 15; k++) w[k] = (unsigned)(int)p[k * sstride];
#pragma unroll
            for (int k = 0; k < 14; k++) pa[k] = __byte_perm(w[k], w[k + 1], 0x5410);
        }
        typename O::type *q = out + y0 * ostride + xc * CPL;
#pragma unroll
        for (int o = 0; o < 8; o++) {
            if (y0 + o < th) {
                int a = rnd;
                a = dp2a_lo<SG>(pa[o], ta, a);
                a = dp2a_hi<SG>(pa[o + 2], ta, a);
                a = dp2a_lo<SG>(pa[o + 4], tb, a);
                a = dp2a_hi<SG>(pa[o + 6], tb, a);
                a = O::fin2(a, sh, bdmax);
                if (CPL == 2) {
                    int b = rnd;
                    b = dp2a_lo<SG>(pb[o], ta, b);
                    b = dp2a_hi<SG>(pb[o + 2], ta, b);
                    b = dp2a_lo<SG>(pb[o + 4], tb, b);
                    b = dp2a_hi<SG>(pb[o + 6], tb, b);
                    b = O::fin2(b, sh, bdmax);
                    mc_store2(q + o * ostride, a, b);
                } else {
                    q[o * ostride] = (typename O::type)a;
                }
            }
        }
    }
}

// One tile of put (PREP = false) / prep (PREP = true), in two steps: mc_stage() brings the reference
// window into shared memory, mc_filter() runs the separable filter over it.
//   ref       reference plane (clamped reads)
//   sx, sy    integer sample position of the tile's top-left in the reference
//   tw, th    tile size (<= 32); bw, bh: full block size (filter selection)
//   out       tile's top-left in the destination (pixels, or int16 for prep)
//   lane      lane index inside the group of G lanes that owns this tile (G = 32: the warp;
//             G = 8: four tiles per warp, `gmask` = the group's lanes for the barriers)
// mc_stage returns `off`: window column c of row r lives at ssrc[r * STRIDE + off + c].
template <typename pixel, int TMAX = MC_T, int G = 32>
DEV int mc_stage(const PlaneView &ref, const int sx, const int sy, const int tw, const int th,
                 const int mx, const int my, pixel *ssrc, const int lane, const unsigned gmask = 0xffffffffu)
{
    // ---- stage the window: rows/cols -3..+4 only where a filter needs them.
    // Fast path (window columns inside the plane): 16-byte cp.async copies of
    // the aligned superset of each row, global -> shared without a register
    // round trip; `off` = position of window column 0 inside the staged row.
    // A lane owns one vector column and walks down the rows (LPR lanes per row).
    // Row indices are clamped in both paths (top/bottom edge emulation); tiles
    // that cross the left/right picture edge take the per-pixel clamped path.
    constexpr int SS = McSrcGeo<pixel, TMAX>::STRIDE, VPX = McSrcGeo<pixel, TMAX>::VPX, MS = McSrcGeo<pixel, TMAX>::MID;
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
            constexpr int LPR = TMAX == 32 ? 8 : 4;           // lanes per row >= vectors per row
            const int lr = lane / LPR, v = v_lo + lane % LPR;
            if (lane % LPR < nv) {
                const pixel *gcol = rp + a0 + v * VPX;
                const unsigned scol = (unsigned)__cvta_generic_to_shared(ssrc + v * VPX);
                for (int r = r_lo + lr; r < r_hi; r += G / LPR) {
                    const int yy = iclip(sy - 3 + r, 0, ref.h - 1);
                    const pixel *g = gcol + yy * rstride;
                    const unsigned sa = scol + r * (SS * (int)sizeof(pixel));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(g) : "memory");
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        } else {
            // per-pixel clamped reads; a lane owns at most two columns and fetches four rows
            // per step so that eight loads are in flight (a load -> shared store -> load chain
            // would serialise on the global-memory latency and leave a long tail in the grid)
            const int ncols = c_hi - c_lo;
            const int lpr = ncols <= 8 || G == 8 ? 8 : ncols <= 16 ? 16 : 32;   // lanes per row, 2 * lpr >= ncols
            const int rpi = G / lpr;                                   // rows per iteration
            const int lr = lane / lpr, lc = lane % lpr;
            const int c1 = c_lo + lc, c2 = c1 + lpr;
            const bool h1 = c1 < c_hi, h2 = c2 < c_hi;
            const int x1 = iclip(sx - 3 + c1, 0, ref.w - 1), x2 = iclip(sx - 3 + c2, 0, ref.w - 1);
            for (int r = r_lo + lr; r < r_hi; r += 4 * rpi) {
                pixel v1[4], v2[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int rr = r + u * rpi;
                    const pixel *row = rp + iclip(sy - 3 + rr, 0, ref.h - 1) * rstride;
                    v1[u] = v2[u] = 0;
                    if (rr < r_hi && h1) v1[u] = row[x1];
                    if (rr < r_hi && h2) v2[u] = row[x2];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int rr = r + u * rpi;
                    if (rr < r_hi && h1) ssrc[rr * SS + c1] = v1[u];
                    if (rr < r_hi && h2) ssrc[rr * SS + c2] = v2[u];
                }
            }
        }
    }
    __syncwarp(gmask);
    return off;
}

template <typename pixel, bool PREP, int TMAX = MC_T, int G = 32>
DEV void mc_filter(const pixel *ssrc, const int off, int16_t *smid, const int tw, const int th,
                   const int bw, const int bh, const int mx, const int my, const int filter_2d,
                   const int bdmax, typename McOut<pixel, PREP>::type *out, const int ostride, const int lane,
                   const unsigned gmask = 0xffffffffu)
{
    typedef McOut<pixel, PREP> O;
    const int ib = PxTraits<pixel>::inter_bits(bdmax);
    const int bs = filter_2d == 9 ? 4 : 6;
    uint2 ph = make_uint2(0, 0), pv = make_uint2(0, 0);
    if (mx) ph = mc_load_taps(filter_2d, false, mx, bw);
    if (my) pv = mc_load_taps(filter_2d, true, my, bh);
    constexpr int SS = McSrcGeo<pixel, TMAX>::STRIDE, MS = McSrcGeo<pixel, TMAX>::MID;
    const pixel *wsrc = ssrc + off;      // window column c lives at wsrc[r * SS + c]
    // paired / vector stores need an even column, an even stride and an aligned base
    const bool vec_ok = (((uintptr_t)out & 7) | (ostride & 3)) == 0;
    const bool pair_ok = (((uintptr_t)out & 3) | (ostride & 1)) == 0 && tw * ((th + 7) >> 3) >= 2 * G;

    if (mx && my) {
        const int sh1 = bs - ib, rnd1 = (1 << sh1) >> 1;
        if (sizeof(pixel) == 2 && tw >= 8) {
            if constexpr (sizeof(pixel) == 2)
                mc_hpass_dp<pixel, PREP, false, TMAX, G>(ssrc, off, ph, tw, 0, th + 7, rnd1, sh1, bdmax, smid, nullptr, 0, false, lane);
        } else {
            int fh[8];
            mc_unpack_taps(ph, fh);
            if (tw >= 8)      mc_hpass<pixel, PREP, 8, false, TMAX, G>(wsrc, fh, tw, 0, th + 7, rnd1, sh1, bdmax, smid, nullptr, 0, lane);
            else if (tw == 4) mc_hpass<pixel, PREP, 4, false, TMAX, G>(wsrc, fh, tw, 0, th + 7, rnd1, sh1, bdmax, smid, nullptr, 0, lane);
            else              mc_hpass<pixel, PREP, 2, false, TMAX, G>(wsrc, fh, tw, 0, th + 7, rnd1, sh1, bdmax, smid, nullptr, 0, lane);
        }
        __syncwarp(gmask);
        const int sh2 = PREP ? bs : bs + ib, rnd2 = (1 << sh2) >> 1;
        if (pair_ok) mc_vpass_dp<pixel, PREP, int16_t, 2, G>(smid, MS, pv, tw, th, rnd2, sh2, bdmax, out, ostride, lane);
        else         mc_vpass_dp<pixel, PREP, int16_t, 1, G>(smid, MS, pv, tw, th, rnd2, sh2, bdmax, out, ostride, lane);
    } else if (mx) {
        const int sh = PREP ? bs - ib : bs;
        const int rnd = PREP ? (1 << sh) >> 1 : (1 << (bs - 1)) + ((1 << (bs - ib)) >> 1);
        if (sizeof(pixel) == 2 && tw >= 8) {
            if constexpr (sizeof(pixel) == 2)
                mc_hpass_dp<pixel, PREP, true, TMAX, G>(ssrc, off, ph, tw, 3, th + 3, rnd, sh, bdmax, nullptr, out, ostride, vec_ok, lane);
        } else {
            int fh[8];
            mc_unpack_taps(ph, fh);
            if (tw >= 8)      mc_hpass<pixel, PREP, 8, true, TMAX, G>(wsrc, fh, tw, 3, th + 3, rnd, sh, bdmax, nullptr, out, ostride, lane);
            else if (tw == 4) mc_hpass<pixel, PREP, 4, true, TMAX, G>(wsrc, fh, tw, 3, th + 3, rnd, sh, bdmax, nullptr, out, ostride, lane);
            else              mc_hpass<pixel, PREP, 2, true, TMAX, G>(wsrc, fh, tw, 3, th + 3, rnd, sh, bdmax, nullptr, out, ostride, lane);
        }
    } else if (my) {
        const int sh = PREP ? bs - ib : bs, rnd = (1 << sh) >> 1;
        mc_vpass_dp<pixel, PREP, pixel, 1, G>(wsrc + 3, SS, pv, tw, th, rnd, sh, bdmax, out, ostride, lane);
    } else {
        const int lw = 31 - __clz(tw);
        for (int i = lane; i < tw * th; i += G) {
            const int y = i >> lw, x = i & (tw - 1);
            const int px = wsrc[(y + 3) * SS + x + 3];
            if (PREP) out[y * ostride + x] = (typename O::type)((px << ib) - PxTraits<pixel>::prep_bias);
            else out[y * ostride + x] = (typename O::type)px;
        }
    }
    __syncwarp(gmask);
}


template <typename pixel, bool PREP, int TMAX = MC_T, int G = 32>
DEV void mc_tile(const PlaneView &ref, const int sx, const int sy, const int tw, const int th,
                 const int bw, const int bh, const int mx, const int my, const int filter_2d,
                 const int bdmax, McSmem<pixel, TMAX> *sm,
                 typename McOut<pixel, PREP>::type *out, const int ostride, const int lane,
                 const unsigned gmask = 0xffffffffu)
{
    const int off = mc_stage<pixel, TMAX, G>(ref, sx, sy, tw, th, mx, my, sm->src, lane, gmask);
    mc_filter<pixel, PREP, TMAX, G>(sm->src, off, sm->mid, tw, th, bw, bh, mx, my, filter_2d, bdmax, out, ostride,
                                    lane, gmask);
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
    // w is a power of two; two pixels per step (one 32-bit read per intermediate, one store)
    // when the rows are even-aligned
    const int lw = 31 - __clz(w);
    const bool pair = w >= 2 && ((((uintptr_t)t1 | (uintptr_t)t2) & 3) | (ts & 1) | (dstride & 1) |
                                 (int)((uintptr_t)dst & (2 * sizeof(pixel) - 1))) == 0;
    if (kind == DAV1D_CUDA_MC_W_MASK) {
        // below
    } else if (pair) {
        const int sh = kind == DAV1D_CUDA_MC_AVG ? ib + 1 : kind == DAV1D_CUDA_MC_W_AVG ? ib + 4 : ib + 6;
        const int rnd = kind == DAV1D_CUDA_MC_AVG ? (1 << ib) + bias * 2
                      : kind == DAV1D_CUDA_MC_W_AVG ? (8 << ib) + bias * 16 : (32 << ib) + bias * 64;
        const int wsum = kind == DAV1D_CUDA_MC_AVG ? 2 : kind == DAV1D_CUDA_MC_W_AVG ? 16 : 64;
        for (int i = tid; i < (w * h) >> 1; i += nthr) {
            const int y = i >> (lw - 1), x = (i & ((w >> 1) - 1)) << 1;
            const unsigned a = *(const unsigned *)(t1 + y * ts + x), b = *(const unsigned *)(t2 + y * ts + x);
            const int a0 = (int)(int16_t)a, a1 = (int)a >> 16, b0 = (int)(int16_t)b, b1 = (int)b >> 16;
            int m0 = 1, m1 = 1;
            if (kind == DAV1D_CUDA_MC_W_AVG) m0 = m1 = weight_or_sign;
            else if (kind == DAV1D_CUDA_MC_MASK) { m0 = mask[y * ms + x]; m1 = mask[y * ms + x + 1]; }
            const int r0 = clip_px<pixel>((a0 * m0 + b0 * (wsum - m0) + rnd) >> sh, bdmax);
            const int r1 = clip_px<pixel>((a1 * m1 + b1 * (wsum - m1) + rnd) >> sh, bdmax);
            mc_store2(dst + y * dstride + x, r0, r1);
        }
    } else if (kind == DAV1D_CUDA_MC_AVG) {
        const int sh = ib + 1, rnd = (1 << ib) + bias * 2;
        for (int i = tid; i < w * h; i += nthr) {
            const int y = i >> lw, x = i & (w - 1);
            dst[y * dstride + x] = (pixel)clip_px<pixel>((t1[y * ts + x] + t2[y * ts + x] + rnd) >> sh, bdmax);
        }
    } else if (kind == DAV1D_CUDA_MC_W_AVG) {
        const int sh = ib + 4, rnd = (8 << ib) + bias * 16, wt = weight_or_sign;
        for (int i = tid; i < w * h; i += nthr) {
            const int y = i >> lw, x = i & (w - 1);
            dst[y * dstride + x] = (pixel)clip_px<pixel>(
                (t1[y * ts + x] * wt + t2[y * ts + x] * (16 - wt) + rnd) >> sh, bdmax);
        }
    } else if (kind == DAV1D_CUDA_MC_MASK) {
        const int sh = ib + 6, rnd = (32 << ib) + bias * 64;
        for (int i = tid; i < w * h; i += nthr) {
            const int y = i >> lw, x = i & (w - 1);
            const int m = mask[y * ms + x];
            dst[y * dstride + x] = (pixel)clip_px<pixel>(
                (t1[y * ts + x] * m + t2[y * ts + x] * (64 - m) + rnd) >> sh, bdmax);
        }
    }
    if (kind == DAV1D_CUDA_MC_W_MASK) {

        const int sh = ib + 6, rnd = (32 << ib) + bias * 64;
        const int bitdepth = PxTraits<pixel>::bitdepth(bdmax);
        const int mask_sh = bitdepth + ib - 4, mask_rnd = 1 << (mask_sh - 5);
        const int sign = weight_or_sign;
        const int ssh = mask_ss >= 1, ssv = mask_ss == 2;
        const int qw = w >> ssh, qh = h >> ssv;
        const int lq = lw - ssh;
        for (int i = tid; i < qw * qh; i += nthr) {
            const int qy = i >> lq, qx = i & (qw - 1);
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
        // all four rows of a lane are summed before the first shared store, so that their
        // global loads overlap (a store in between would order them)
        int sums[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = lane + 32 * u;
            const int y = imin(i >> 3, 14), x = i & 7;
            const int tmx = mx0 + y * a1 + x * a0;
            const int8_t *f = g_warp_filter + (64 + ((tmx + 512) >> 10)) * 8;
            const int yy = iclip(sy - 3 + y, 0, ref.h - 1);
            const pixel *row = rp + yy * rstride;
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += f[k] * row[iclip(sx + x - 3 + k, 0, ref.w - 1)];
            sums[u] = sum;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = lane + 32 * u;
            if (i < 15 * 8) mid[i] = (int16_t)((sums[u] + rnd) >> sh);
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
