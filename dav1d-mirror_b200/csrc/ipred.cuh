// Intra prediction device functions, executed by one warp per (transform)
// block.  `edge` is the reference's `topleft` pointer convention
// (src/ipred_prepare.h:66-72): edge[0] corner, edge[1..] top (+top-right),
// edge[-1..] left going down (+bottom-left).  `dst` may be global or shared.
//
// Reference being matched bit for bit: src/ipred_tmpl.c
//   DC family :86-218   V/H :220-242   Paeth :244-265   smooth* :267-325
//   edge filter/upsample :327-406   Z1 :408-460  Z2 :462-540  Z3 :542-599
//   filter-intra :618-655   cfl_ac :657-703   cfl_pred :71-84   pal_pred :717-730
// and src/ipred_prepare_tmpl.c:77-204 for the on-device edge preparation.
#pragma once
#include "common.cuh"
#include "tables.cuh"
#include "itx.cuh"     // load_px / store_px

namespace d1 {

enum {
    M_DC = 0, M_VERT, M_HOR, M_LEFT_DC, M_TOP_DC, M_DC_128, M_Z1, M_Z2, M_Z3,
    M_SMOOTH, M_SMOOTH_V, M_SMOOTH_H, M_PAETH, M_FILTER
};

constexpr int IPRED_SCRATCH = 2 * 64 + 2 * 64 + 16;   // pixels of edge scratch for Z modes

DEV int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// dc value for the DC / TOP_DC / LEFT_DC / DC_128 variants (ipred_tmpl.c:86-218)
template <typename pixel>
DEV int ipred_dc_value(const int mode, const pixel *edge, const int w, const int h, const int bdmax, const int lane) {
    if (mode == M_DC_128) return (bdmax + 1) >> 1;
    int part = 0;
    if (mode == M_DC || mode == M_TOP_DC)
        for (int i = lane; i < w; i += 32) part += edge[1 + i];
    if (mode == M_DC || mode == M_LEFT_DC)
        for (int i = lane; i < h; i += 32) part += edge[-(1 + i)];
    const unsigned sum = (unsigned)warp_sum(part);
    if (mode == M_TOP_DC) return (int)((sum + (w >> 1)) >> (31 - __clz(w)));
    if (mode == M_LEFT_DC) return (int)((sum + (h >> 1)) >> (31 - __clz(h)));
    unsigned dc = sum + ((w + h) >> 1);
    dc >>= __ffs(w + h) - 1;
    if (w != h) {
        const bool x4 = w > h * 2 || h > w * 2;
        if (PxTraits<pixel>::hbd) dc = (dc * (x4 ? 0x6667u : 0xAAABu)) >> 17;
        else dc = (dc * (x4 ? 0x3334u : 0x5556u)) >> 16;
    }
    return (int)dc;
}

DEV int filter_strength(const int wh, const int angle, const int is_sm) {   // ipred_tmpl.c:327-360
    if (is_sm) {
        if (wh <= 8) { if (angle >= 64) return 2; if (angle >= 40) return 1; }
        else if (wh <= 16) { if (angle >= 48) return 2; if (angle >= 20) return 1; }
        else if (wh <= 24) { if (angle >= 4) return 3; }
        else return 3;
    } else {
        if (wh <= 8) { if (angle >= 56) return 1; }
        else if (wh <= 16) { if (angle >= 40) return 1; }
        else if (wh <= 24) { if (angle >= 32) return 3; if (angle >= 16) return 2; if (angle >= 8) return 1; }
        else if (wh <= 32) { if (angle >= 32) return 3; if (angle >= 4) return 2; return 1; }
        else return 3;
    }
    return 0;
}
DEV int use_upsample(const int wh, const int angle, const int is_sm) { return angle < 40 && wh <= (16 >> is_sm); }

// filter_edge (ipred_tmpl.c:362-385), out[0..sz)
template <typename pixel>
DEV void edge_filter(pixel *out, const int sz, const int lim_from, const int lim_to, const pixel *in,
                     const int from, const int to, const int strength, const int lane)
{
    const int k0 = strength == 3 ? 2 : 0;
    const int k1 = strength == 1 ? 4 : strength == 2 ? 5 : 4;
    const int k2 = strength == 1 ? 8 : strength == 2 ? 6 : 4;
    for (int i = lane; i < sz; i += 32) {
        int v;
        if (i < imin(sz, lim_from) || i >= imin(lim_to, sz)) {
            v = in[iclip(i, from, to - 1)];
        } else {
            const int s = in[iclip(i - 2, from, to - 1)] * k0 + in[iclip(i - 1, from, to - 1)] * k1 +
                          in[iclip(i, from, to - 1)] * k2 + in[iclip(i + 1, from, to - 1)] * k1 +
                          in[iclip(i + 2, from, to - 1)] * k0;
            v = (s + 8) >> 4;
        }
        out[i] = (pixel)v;
    }
}

// upsample_edge (ipred_tmpl.c:391-406), out[0 .. 2*hsz-2]
template <typename pixel>
DEV void edge_upsample(pixel *out, const int hsz, const pixel *in, const int from, const int to,
                       const int bdmax, const int lane)
{
    for (int i = lane; i < hsz; i += 32) {
        out[i * 2] = in[iclip(i, from, to - 1)];
        if (i < hsz - 1) {
            const int s = -in[iclip(i - 1, from, to - 1)] + 9 * in[iclip(i, from, to - 1)] +
                          9 * in[iclip(i + 1, from, to - 1)] - in[iclip(i + 2, from, to - 1)];
            out[i * 2 + 1] = (pixel)clip_px<pixel>((s + 8) >> 4, bdmax);
        }
    }
}


// ---- vectorised block fill: each lane produces VW = 8 (w >= 8) or 4 (w == 4)
// consecutive pixels of a row and writes them with one 64/128-bit store.
// Blocks are aligned to their own width (AV1 partitioning), so the vector
// stores are naturally aligned; misaligned callers fall back to scalar stores.
// f(x, y) -> pixel value
template <typename pixel, typename F>
DEV void fill_block(pixel *dst, const int dstride, const int w, const int h, const int lane, F f) {
    if (w >= 8) {
        const int segs = w >> 3, sh = 31 - __clz(segs);       // segs = 1, 2, 4, 8
        const int x0 = (lane & (segs - 1)) << 3;
        for (int y = lane >> sh; y < h; y += 32 >> sh) {
            int v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = f(x0 + k, y);
            store_px<pixel, 8>(dst + y * dstride + x0, v);
        }
    } else {
        for (int y = lane; y < h; y += 32) {
            int v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = f(k, y);
            store_px<pixel, 4>(dst + y * dstride, v);
        }
    }
}

// All 14 predictors. `scratch`: IPRED_SCRATCH pixels of shared memory (Z modes).
template <typename pixel>
DEV void ipred_block(const int mode, pixel *dst, const int dstride, const pixel *edge, const int w, const int h,
                     int angle, const int max_w, const int max_h, const int bdmax, pixel *scratch, const int lane)
{
    switch (mode) {
    case M_DC: case M_TOP_DC: case M_LEFT_DC: case M_DC_128: {
        const int dc = ipred_dc_value<pixel>(mode, edge, w, h, bdmax, lane);
        fill_block<pixel>(dst, dstride, w, h, lane, [=](int, int) { return dc; });
        break;
    }
    case M_VERT:
        fill_block<pixel>(dst, dstride, w, h, lane, [=](int x, int) { return (int)edge[1 + x]; });
        break;
    case M_HOR:
        fill_block<pixel>(dst, dstride, w, h, lane, [=](int, int y) { return (int)edge[-(1 + y)]; });
        break;
    case M_PAETH: {
        const int tl = edge[0];
        fill_block<pixel>(dst, dstride, w, h, lane, [=](int x, int y) {
            const int left = edge[-(y + 1)], top = edge[1 + x];
            const int base = left + top - tl;
            const int ld = iabs(left - base), td = iabs(top - base), tld = iabs(tl - base);
            return ld <= td && ld <= tld ? left : td <= tld ? top : tl;
        });
        break;
    }
    case M_SMOOTH: {
        const uint8_t *wh = g_sm_weights + w, *wv = g_sm_weights + h;
        const int right = edge[w], bottom = edge[-h];
        fill_block<pixel>(dst, dstride, w, h, lane, [=](int x, int y) {
            const int p = wv[y] * edge[1 + x] + (256 - wv[y]) * bottom +
                          wh[x] * edge[-(1 + y)] + (256 - wh[x]) * right;
            return (p + 256) >> 9;
        });
        break;
    }
    case M_SMOOTH_V: {
        const uint8_t *wv = g_sm_weights + h;
        const int bottom = edge[-h];
        fill_block<pixel>(dst, dstride, w, h, lane, [=](int x, int y) {
            return (wv[y] * edge[1 + x] + (256 - wv[y]) * bottom + 128) >> 8;
        });
        break;
    }
    case M_SMOOTH_H: {
        const uint8_t *wh = g_sm_weights + w;
        const int right = edge[w];
        fill_block<pixel>(dst, dstride, w, h, lane, [=](int x, int y) {
            return (wh[x] * edge[-(y + 1)] + (256 - wh[x]) * right + 128) >> 8;
        });
        break;
    }
    case M_Z1: {
        const int is_sm = (angle >> 9) & 1, ef = angle >> 10;
        angle &= 511;
        int dx = g_dr_intra_derivative[angle >> 1];
        const int ups = ef ? use_upsample(w + h, 90 - angle, is_sm) : 0;
        const pixel *top;
        int max_base_x;
        if (ups) {
            edge_upsample<pixel>(scratch, w + h, edge + 1, -1, w + imin(w, h), bdmax, lane);
            top = scratch;
            max_base_x = 2 * (w + h) - 2;
            dx <<= 1;
        } else {
            const int fs = ef ? filter_strength(w + h, 90 - angle, is_sm) : 0;
            if (fs) {
                edge_filter<pixel>(scratch, w + h, 0, w + h, edge + 1, -1, w + imin(w, h), fs, lane);
                top = scratch;
                max_base_x = w + h - 1;
            } else {
                top = edge + 1;
                max_base_x = w + imin(w, h) - 1;
            }
        }
        __syncwarp();
        const int inc = 1 + ups;
        fill_block<pixel>(dst, dstride, w, h, lane, [=](int x, int y) {
            const int xpos = (y + 1) * dx, frac = xpos & 0x3E;
            const int base = (xpos >> 6) + x * inc;
            if (base < max_base_x) return (top[base] * (64 - frac) + top[base + 1] * frac + 32) >> 6;
            return (int)top[max_base_x];
        });
        break;
    }
    case M_Z3: {
        const int is_sm = (angle >> 9) & 1, ef = angle >> 10;
        angle &= 511;
        int dy = g_dr_intra_derivative[(270 - angle) >> 1];
        const int ups = ef ? use_upsample(w + h, angle - 180, is_sm) : 0;
        const pixel *left;
        int max_base_y;
        if (ups) {
            edge_upsample<pixel>(scratch, w + h, edge - (w + h), imax(w - h, 0), w + h + 1, bdmax, lane);
            left = scratch + 2 * (w + h) - 2;
            max_base_y = 2 * (w + h) - 2;
            dy <<= 1;
        } else {
            const int fs = ef ? filter_strength(w + h, angle - 180, is_sm) : 0;
            if (fs) {
                edge_filter<pixel>(scratch, w + h, 0, w + h, edge - (w + h), imax(w - h, 0), w + h + 1, fs, lane);
                left = scratch + w + h - 1;
                max_base_y = w + h - 1;
            } else {
                left = edge - 1;
                max_base_y = h + imin(w, h) - 1;
            }
        }
        __syncwarp();
        const int inc = 1 + ups;
        fill_block<pixel>(dst, dstride, w, h, lane, [=](int x, int y) {
            const int ypos = (x + 1) * dy, frac = ypos & 0x3E;
            const int base = (ypos >> 6) + y * inc;
            if (base < max_base_y) return (left[-base] * (64 - frac) + left[-(base + 1)] * frac + 32) >> 6;
            return (int)left[-max_base_y];
        });
        break;
    }
    case M_Z2: {
        const int is_sm = (angle >> 9) & 1, ef = angle >> 10;
        angle &= 511;
        int dy = g_dr_intra_derivative[(angle - 90) >> 1];
        int dx = g_dr_intra_derivative[(180 - angle) >> 1];
        const int ups_l = ef ? use_upsample(w + h, 180 - angle, is_sm) : 0;
        const int ups_a = ef ? use_upsample(w + h, angle - 90, is_sm) : 0;
        pixel *tl = scratch + 128 + 8;
        if (ups_a) {
            edge_upsample<pixel>(tl, w + 1, edge, 0, w + 1, bdmax, lane);
            dx <<= 1;
        } else {
            const int fs = ef ? filter_strength(w + h, angle - 90, is_sm) : 0;
            if (fs) edge_filter<pixel>(tl + 1, w, 0, max_w, edge + 1, -1, w, fs, lane);
            else for (int i = lane; i < w; i += 32) tl[1 + i] = edge[1 + i];
        }
        if (ups_l) {
            edge_upsample<pixel>(tl - h * 2, h + 1, edge - h, 0, h + 1, bdmax, lane);
            dy <<= 1;
        } else {
            const int fs = ef ? filter_strength(w + h, 180 - angle, is_sm) : 0;
            if (fs) edge_filter<pixel>(tl - h, h, h - max_h, h, edge - h, 0, h + 1, fs, lane);
            else for (int i = lane; i < h; i += 32) tl[-h + i] = edge[-h + i];
        }
        __syncwarp();
        if (lane == 0) tl[0] = edge[0];
        __syncwarp();
        const int inc_x = 1 + ups_a;
        const pixel *left = tl - (1 + ups_l);
        const pixel *tlc = tl;
        fill_block<pixel>(dst, dstride, w, h, lane, [=](int x, int y) {
            const int xpos = ((1 + ups_a) << 6) - dx * (y + 1);
            const int base_x = (xpos >> 6) + x * inc_x;
            int v;
            if (base_x >= 0) {
                const int fx = xpos & 0x3E;
                v = tlc[base_x] * (64 - fx) + tlc[base_x + 1] * fx;
            } else {
                const int ypos = (y << (6 + ups_l)) - dy * (x + 1);
                const int base_y = ypos >> 6, fy = ypos & 0x3E;
                v = left[-base_y] * (64 - fy) + left[-(base_y + 1)] * fy;
            }
            return (v + 32) >> 6;
        });
        break;
    }
    default: {   // M_FILTER: 4x2 sub-blocks on anti-diagonals
        const int8_t *taps = g_filter_intra_taps + (angle & 511) * 64;
        const int nbx = w >> 2, nby = h >> 1;
        for (int d = 0; d < nbx + nby - 1; d++) {
            // sub-blocks with bx + by == d; 8 outputs each
            const int by_lo = imax(0, d - (nbx - 1)), by_hi = imin(d, nby - 1);
            const int cnt = (by_hi - by_lo + 1) * 8;
            for (int i = lane; i < cnt; i += 32) {
                const int by = by_lo + (i >> 3), bx = d - by, o = i & 7;
                const int x = bx * 4, y = by * 2;
                int p[7];
                // p0 = (x-1, y-1), p1..p4 = (x..x+3, y-1), p5 = (x-1, y), p6 = (x-1, y+1)
                if (y == 0) {
                    p[0] = edge[x];           // edge[0] when x == 0, else top[x-1]
#pragma unroll
                    for (int k = 0; k < 4; k++) p[1 + k] = edge[1 + x + k];
                } else {
                    p[0] = x == 0 ? edge[-y] : dst[(y - 1) * dstride + x - 1];
#pragma unroll
                    for (int k = 0; k < 4; k++) p[1 + k] = dst[(y - 1) * dstride + x + k];
                }
                if (x == 0) {
                    p[5] = edge[-(1 + y)];
                    p[6] = edge[-(2 + y)];
                } else {
                    p[5] = dst[y * dstride + x - 1];
                    p[6] = dst[(y + 1) * dstride + x - 1];
                }
                const int8_t *f = taps + o * 8;
                int acc = 0;
#pragma unroll
                for (int k = 0; k < 7; k++) acc += f[k] * p[k];
                dst[(y + (o >> 2)) * dstride + x + (o & 3)] = (pixel)clip_px<pixel>((acc + 8) >> 4, bdmax);
            }
            __syncwarp();
        }
        break;
    }
    }
    __syncwarp();
}

// cfl_pred (ipred_tmpl.c:71-84) with the dc of `dc_mode` (DC / LEFT_DC / TOP_DC / DC_128)
template <typename pixel>
DEV void cfl_pred_block(const int dc_mode, pixel *dst, const int dstride, const pixel *edge, const int w, const int h,
                        const int16_t *ac, const int alpha, const int bdmax, const int lane)
{
    const int dc = ipred_dc_value<pixel>(dc_mode, edge, w, h, bdmax, lane);
    fill_block<pixel>(dst, dstride, w, h, lane, [=](int x, int y) {
        const int diff = alpha * ac[y * w + x];
        const int m = (iabs(diff) + 32) >> 6;
        return clip_px<pixel>(dc + (diff < 0 ? -m : m), bdmax);
    });
    __syncwarp();
}

// cfl_ac (ipred_tmpl.c:657-703): ac[w*h] dense from the reconstructed luma
template <typename pixel>
DEV void cfl_ac_block(int16_t *ac, const pixel *ypx, const int ystride, const int w_pad, const int h_pad,
                      const int w, const int h, const int ss_hor, const int ss_ver, const int lane)
{
    const int vw = w - 4 * w_pad, vh = h - 4 * h_pad;
    int part = 0;
    for (int i = lane; i < w * h; i += 32) {
        const int y = imin(i / w, vh - 1), x = imin(i % w, vw - 1);
        const pixel *p = ypx + (y << ss_ver) * ystride + (x << ss_hor);
        int s = __ldcg(p);
        if (ss_hor) s += __ldcg(p + 1);
        if (ss_ver) {
            s += __ldcg(p + ystride);
            if (ss_hor) s += __ldcg(p + ystride + 1);
        }
        const int v = s << (1 + !ss_ver + !ss_hor);
        ac[i] = (int16_t)v;
        part += v;
    }
    const int log2sz = (__ffs(w) - 1) + (__ffs(h) - 1);
    const int sum = (warp_sum(part) + ((1 << log2sz) >> 1)) >> log2sz;
    __syncwarp();
    for (int i = lane; i < w * h; i += 32) ac[i] = (int16_t)(ac[i] - sum);
    __syncwarp();
}

// pal_pred (ipred_tmpl.c:717-730): two pixels per index byte
template <typename pixel>
DEV void pal_pred_block(pixel *dst, const int dstride, const pixel *pal, const uint8_t *idx, const int w, const int h,
                        const int tid, const int nthr)
{
    const int hw = w >> 1;
    for (int i = tid; i < hw * h; i += nthr) {
        const int y = i / hw, x2 = i % hw;
        const int v = idx[i];
        dst[y * dstride + 2 * x2] = pal[v & 7];
        dst[y * dstride + 2 * x2 + 1] = pal[v >> 4];
    }
}

// ------------------------------------------------------------------ edge preparation
// dav1d_prepare_intra_edges (ipred_prepare_tmpl.c:77-204).  x, y, w, h, tw, th
// in 4-pixel units; `dst` = the block's top-left in the frame; `top_sb_edge`
// = optional pre-filter row backup (may be null).  Writes edge[-2*th*4 ..
// 2*tw*4] as needed and returns the DSP mode index; *angle: in = angle_delta,
// out = absolute angle.
template <typename pixel>
DEV int prepare_edges(const int x, const int have_left, const int y, const int have_top, const int w, const int h,
                      const int edge_flags, const pixel *dst, const int stride, const pixel *top_sb_edge,
                      int mode, int *angle, const int tw, const int th, const int filter_edge_flag,
                      pixel *edge, const int bdmax, const int lane)
{
    const int bitdepth = PxTraits<pixel>::bitdepth(bdmax);
    if (mode >= 1 && mode <= 8) {            // VERT .. VERT_LEFT: directional
        int a;
        switch (mode) {
        case 1: a = 90; break;  case 2: a = 180; break; case 3: a = 45; break;  case 4: a = 135; break;
        case 5: a = 113; break; case 6: a = 157; break; case 7: a = 203; break; default: a = 67; break;
        }
        a += 3 * *angle;
        *angle = a;
        if (a <= 90) mode = a < 90 && have_top ? M_Z1 : M_VERT;
        else if (a < 180) mode = M_Z2;
        else mode = a > 180 && have_left ? M_Z3 : M_HOR;
    } else if (mode == 0) {                  // DC_PRED
        mode = have_left ? (have_top ? M_DC : M_LEFT_DC) : (have_top ? M_TOP_DC : M_DC_128);
    } else if (mode == 12) {                 // PAETH_PRED
        mode = have_left ? (have_top ? M_PAETH : M_HOR) : (have_top ? M_VERT : M_DC_128);
    }
    // needs: bit0 left, bit1 top, bit2 topleft, bit3 topright, bit4 bottomleft
    int needs;
    switch (mode) {
    case M_DC: needs = 3; break;
    case M_VERT: needs = 2; break;
    case M_HOR: needs = 1; break;
    case M_LEFT_DC: needs = 1; break;
    case M_TOP_DC: needs = 2; break;
    case M_DC_128: needs = 0; break;
    case M_Z1: needs = 2 | 8 | 4; break;
    case M_Z2: needs = 1 | 2 | 4; break;
    case M_Z3: needs = 1 | 16 | 4; break;
    case M_SMOOTH: case M_SMOOTH_V: case M_SMOOTH_H: needs = 3; break;
    default: needs = 1 | 2 | 4; break;       // PAETH, FILTER
    }
    const pixel *dst_top = nullptr;
    if (have_top && ((needs & 2) || (needs & 4) || ((needs & 1) && !have_left)))
        dst_top = top_sb_edge ? top_sb_edge + x * 4 : dst - stride;

    // Every edge entry is one pixel of the frame (or a constant).  All loads are
    // issued first - they are independent, so their latencies overlap - and the
    // shared-memory stores follow; one warp barrier at the end.
    const int mid = (1 << bitdepth) >> 1;
    const int szl = th << 2, szt = tw << 2;
    const int have_l = have_left && (needs & 1), pxl = imin(szl, (h - y) << 2);
    const int have_bl = (needs & 16) && have_left && y + th < h && (edge_flags & 8);
    const int pxbl = imin(szl, (h - y - th) << 2);
    const int pxt = imin(szt, (w - x) << 2);
    const int have_tr = (needs & 8) && have_top && x + tw < w && (edge_flags & 1);
    const int pxtr = imin(szt, (w - x - tw) << 2);
    auto ld = [](const pixel *p, const int dflt) { return p ? (int)__ldcg(p) : dflt; };
    // fallbacks when a side is unavailable (ipred_prepare_tmpl.c:139-197)
    const pixel *no_left = have_top ? dst_top : nullptr;                      // else mid + 1
    const pixel *no_top = have_left ? dst - 1 : nullptr;                      // else mid - 1
    int vl[2] = { 0, 0 }, vbl[2] = { 0, 0 }, vt[2] = { 0, 0 }, vtr[2] = { 0, 0 }, vc = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int i = lane + 32 * k;
        if ((needs & 1) && i < szl)
            vl[k] = have_left ? (int)__ldcg(dst + stride * imin(i, pxl - 1) - 1) : ld(no_left, mid + 1);
        if ((needs & 16) && i < szl)
            vbl[k] = have_bl ? (int)__ldcg(dst + (szl + imin(i, pxbl - 1)) * stride - 1)
                   : have_left ? (int)__ldcg(dst + stride * (pxl - 1) - 1) : ld(no_left, mid + 1);
        if ((needs & 2) && i < szt)
            vt[k] = have_top ? (int)__ldcg(dst_top + imin(i, pxt - 1)) : ld(no_top, mid - 1);
        if ((needs & 8) && i < szt)
            vtr[k] = have_tr ? (int)__ldcg(dst_top + szt + imin(i, pxtr - 1))
                   : have_top ? (int)__ldcg(dst_top + pxt - 1) : ld(no_top, mid - 1);
    }
    if ((needs & 4) && lane == 0) {
        if (have_left) vc = have_top ? __ldcg(dst_top - 1) : __ldcg(dst - 1);
        else vc = have_top ? (int)__ldcg(dst_top) : mid;
    }
    (void)have_l;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int i = lane + 32 * k;
        if ((needs & 1) && i < szl) edge[-1 - i] = (pixel)vl[k];             // left[szl-1-i] = edge[-szl + szl-1-i]
        if ((needs & 16) && i < szl) edge[-szl - 1 - i] = (pixel)vbl[k];
        if ((needs & 2) && i < szt) edge[1 + i] = (pixel)vt[k];
        if ((needs & 8) && i < szt) edge[1 + szt + i] = (pixel)vtr[k];
    }
    if ((needs & 4) && lane == 0) {
        // Z2 corner smoothing uses edge[-1] and edge[1]: both live in lane 0's registers
        if (mode == M_Z2 && tw + th >= 6 && filter_edge_flag) vc = ((vl[0] + vt[0]) * 5 + vc * 6 + 8) >> 4;
        edge[0] = (pixel)vc;
    }
    __syncwarp();
    return mode;
}

}  // namespace d1
