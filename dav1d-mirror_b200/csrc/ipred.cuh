// Intra prediction device functions, executed by a GROUP of G lanes of a warp per (transform)
// block: G = 32 (one block per warp) or 8 (four small blocks per warp, one per octet; the
// octets of a warp run independently - every barrier / shuffle names the group's lanes only).
// `edge` is the reference's `topleft` pointer convention (src/ipred_prepare.h:66-72): edge[0]
// corner, edge[1..] top (+top-right), edge[-1..] left going down (+bottom-left).
//
// Structure: prepare_edges() gathers the edge from the frame, ipred_setup() turns (mode, angle)
// into a small parameter block (filtering / upsampling the edge where the Z modes ask for it),
// and ipred_pixel() evaluates ONE pixel from that block.  Callers run a single loop over the
// block's pixels with ipred_pixel() inside: the code stays small and loop-structured, which is
// what keeps the warps of an SM - all busy with different blocks - inside the instruction caches.
//
// Reference being matched bit for bit: src/ipred_tmpl.c
//   DC family :86-218   V/H :220-242   Paeth :244-265   smooth* :267-325
//   edge filter/upsample :327-406   Z1 :408-460  Z2 :462-540  Z3 :542-599
//   filter-intra :618-655   cfl_ac :657-703   cfl_pred :71-84   pal_pred :717-730
// and src/ipred_prepare_tmpl.c:77-204 for the on-device edge preparation.
#pragma once
#include "common.cuh"
#include "tables.cuh"
#include "itx_geom.cuh"     // load_px / store_px

namespace d1 {

enum {
    M_DC = 0, M_VERT, M_HOR, M_LEFT_DC, M_TOP_DC, M_DC_128, M_Z1, M_Z2, M_Z3,
    M_SMOOTH, M_SMOOTH_V, M_SMOOTH_H, M_PAETH, M_FILTER
};

constexpr int IPRED_SCRATCH = 2 * 64 + 2 * 64 + 16;   // pixels of edge scratch for Z modes

// lane `gl` of a group of `G` lanes (power of two, aligned) whose lanes are `mask`
struct Grp {
    int gl, G;
    unsigned mask;
};
DEV Grp grp_warp(const int lane) { return Grp{ lane, 32, 0xffffffffu }; }
DEV Grp grp_octet(const int lane) { return Grp{ lane & 7, 8, 0xffu << (lane & 24) }; }
DEV Grp grp_quad(const int lane) { return Grp{ lane & 3, 4, 0xfu << (lane & 28) }; }
DEV void grp_sync(const Grp &g) { __syncwarp(g.mask); }
DEV int grp_sum(const Grp &g, int v) {
    for (int o = g.G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(g.mask, v, o);
    return v;
}

// bitstream mode (0..12 enum IntraPredMode, 13 = filter) + angle_delta -> DSP-table mode; *angle:
// in = angle_delta, out = absolute angle for the directional modes (ipred_prepare_tmpl.c:94-117)
HD int ipred_resolve_mode(int mode, int *angle, const int have_left, const int have_top) {
    if (mode >= 1 && mode <= 8) {            // VERT .. VERT_LEFT: directional
        const int base = mode == 1 ? 90 : mode == 2 ? 180 : mode == 3 ? 45 : mode == 4 ? 135 : mode == 5 ? 113
                       : mode == 6 ? 157 : mode == 7 ? 203 : 67;
        const int a = base + 3 * *angle;
        *angle = a;
        if (a <= 90) return a < 90 && have_top ? M_Z1 : M_VERT;
        if (a < 180) return M_Z2;
        return a > 180 && have_left ? M_Z3 : M_HOR;
    }
    if (mode == 0) return have_left ? (have_top ? M_DC : M_LEFT_DC) : (have_top ? M_TOP_DC : M_DC_128);
    if (mode == 12) return have_left ? (have_top ? M_PAETH : M_HOR) : (have_top ? M_VERT : M_DC_128);
    return mode;
}
// edges a DSP-table mode reads: bit0 left, bit1 top, bit2 topleft, bit3 topright, bit4 bottomleft
// (ipred_prepare_tmpl.c:50-74)
HD int ipred_mode_needs(const int m) {
    switch (m) {
    case M_DC: return 3;
    case M_VERT: return 2;
    case M_HOR: return 1;
    case M_LEFT_DC: return 1;
    case M_TOP_DC: return 2;
    case M_DC_128: return 0;
    case M_Z1: return 2 | 8 | 4;
    case M_Z2: return 1 | 2 | 4;
    case M_Z3: return 1 | 16 | 4;
    case M_SMOOTH: case M_SMOOTH_V: case M_SMOOTH_H: return 3;
    default: return 1 | 2 | 4;               // PAETH, FILTER
    }
}

static __device__ __noinline__ int filter_strength(const int wh, const int angle, const int is_sm) {   // ipred_tmpl.c:327-360
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
__device__ __noinline__ void edge_filter(const Grp &g, pixel *out, const int sz, const int lim_from, const int lim_to, const pixel *in,
                     const int from, const int to, const int strength)
{
    const int k0 = strength == 3 ? 2 : 0;
    const int k1 = strength == 1 ? 4 : strength == 2 ? 5 : 4;
    const int k2 = strength == 1 ? 8 : strength == 2 ? 6 : 4;
    for (int i = g.gl; i < sz; i += g.G) {
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
__device__ __noinline__ void edge_upsample(const Grp &g, pixel *out, const int hsz, const pixel *in, const int from, const int to,
                       const int bdmax)
{
    for (int i = g.gl; i < hsz; i += g.G) {
        out[i * 2] = in[iclip(i, from, to - 1)];
        if (i < hsz - 1) {
            const int s = -in[iclip(i - 1, from, to - 1)] + 9 * in[iclip(i, from, to - 1)] +
                          9 * in[iclip(i + 1, from, to - 1)] - in[iclip(i + 2, from, to - 1)];
            out[i * 2 + 1] = (pixel)clip_px<pixel>((s + 8) >> 4, bdmax);
        }
    }
}

// dc value for the DC / TOP_DC / LEFT_DC / DC_128 variants (ipred_tmpl.c:86-218)
template <typename pixel>
__device__ __noinline__ int ipred_dc_value(const Grp &g, const int mode, const pixel *edge, const int w, const int h, const int bdmax) {
    if (mode == M_DC_128) return (bdmax + 1) >> 1;
    int part = 0;
    if (mode == M_DC || mode == M_TOP_DC)
        for (int i = g.gl; i < w; i += g.G) part += edge[1 + i];
    if (mode == M_DC || mode == M_LEFT_DC)
        for (int i = g.gl; i < h; i += g.G) part += edge[-(1 + i)];
    const unsigned sum = (unsigned)grp_sum(g, part);
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

// ------------------------------------------------------------------ one pixel
// pixel modes of the loop: the 14 predictors collapse to ten cases, plus the tile (filter-intra)
// and the CfL cases
enum PixMode {
    PM_CONST, PM_V, PM_H, PM_PAETH, PM_SMOOTH, PM_SMOOTH_V, PM_SMOOTH_H, PM_Z1, PM_Z2, PM_Z3,
    PM_TILE, PM_CFL, PM_PAL
};

template <typename pixel> struct PixParams {
    int pm;
    int p0, p1, p2, p3;      // meaning depends on pm (see ipred_setup)
    const pixel *edge;       // prepared edge, centre
    const pixel *e0, *e1;    // (filtered / upsampled) edge arrays of the Z modes
    const void *tile;        // PM_TILE: w*h pixels; PM_CFL: w*h int16 ac values
    int w, h;
};

// (mode m, angle with the flag bits 9 / 10 of src/ipred_prepare.h:87-93) -> parameters.  scratch:
// IPRED_SCRATCH pixels for the Z modes (blocks of up to 16 + 4: 80 pixels with z2_centre = 40);
// tile: w*h pixels for filter-intra (w, h <= 32).
// tile_t: sample type of the filter-intra tile (pixel for the per-call kernel, uint16_t for the executor)
template <typename pixel, typename tile_t = pixel>
DEV PixParams<pixel> ipred_setup(const Grp &g, const int m, const int angle_in, const int w, const int h,
                                 const int max_w, const int max_h, const pixel *edge, pixel *scratch, tile_t *tile,
                                 const int bdmax, const int z2_centre = 128 + 8)
{
    PixParams<pixel> P;
    P.pm = PM_CONST; P.p0 = P.p1 = P.p2 = P.p3 = 0;
    P.edge = edge; P.e0 = edge; P.e1 = edge; P.tile = tile; P.w = w; P.h = h;
    const int is_sm = (angle_in >> 9) & 1, ef = (angle_in >> 10) & 1;
    const int ang = angle_in & 511;
    switch (m) {
    case M_DC: case M_TOP_DC: case M_LEFT_DC: case M_DC_128:
        P.p0 = ipred_dc_value<pixel>(g, m, edge, w, h, bdmax);
        break;
    case M_VERT: P.pm = PM_V; break;
    case M_HOR: P.pm = PM_H; break;
    case M_PAETH: P.pm = PM_PAETH; P.p0 = edge[0]; break;
    case M_SMOOTH: P.pm = PM_SMOOTH; P.p0 = edge[w]; P.p1 = edge[-h]; break;         // right, bottom
    case M_SMOOTH_V: P.pm = PM_SMOOTH_V; P.p1 = edge[-h]; break;
    case M_SMOOTH_H: P.pm = PM_SMOOTH_H; P.p0 = edge[w]; break;
    case M_Z1: {
        // ipred_z1_c (ipred_tmpl.c:408-460): p0 = dx, p1 = max_base_x, p2 = 1 + upsample, e0 = top
        int dx = g_dr_intra_derivative[ang >> 1];
        const int ups = ef ? use_upsample(w + h, 90 - ang, is_sm) : 0;
        if (ups) {
            edge_upsample<pixel>(g, scratch, w + h, edge + 1, -1, w + imin(w, h), bdmax);
            P.e0 = scratch; P.p1 = 2 * (w + h) - 2; dx <<= 1;
        } else {
            const int fs = ef ? filter_strength(w + h, 90 - ang, is_sm) : 0;
            if (fs) {
                edge_filter<pixel>(g, scratch, w + h, 0, w + h, edge + 1, -1, w + imin(w, h), fs);
                P.e0 = scratch; P.p1 = w + h - 1;
            } else {
                P.e0 = edge + 1; P.p1 = w + imin(w, h) - 1;
            }
        }
        P.p0 = dx; P.p2 = 1 + ups;
        P.pm = PM_Z1;
        break;
    }
    case M_Z3: {
        // ipred_z3_c (:542-599): p0 = dy, p1 = max_base_y, p2 = 1 + upsample, e0 = left
        int dy = g_dr_intra_derivative[(270 - ang) >> 1];
        const int ups = ef ? use_upsample(w + h, ang - 180, is_sm) : 0;
        if (ups) {
            edge_upsample<pixel>(g, scratch, w + h, edge - (w + h), imax(w - h, 0), w + h + 1, bdmax);
            P.e0 = scratch + 2 * (w + h) - 2; P.p1 = 2 * (w + h) - 2; dy <<= 1;
        } else {
            const int fs = ef ? filter_strength(w + h, ang - 180, is_sm) : 0;
            if (fs) {
                edge_filter<pixel>(g, scratch, w + h, 0, w + h, edge - (w + h), imax(w - h, 0), w + h + 1, fs);
                P.e0 = scratch + w + h - 1; P.p1 = w + h - 1;
            } else {
                P.e0 = edge - 1; P.p1 = h + imin(w, h) - 1;
            }
        }
        P.p0 = dy; P.p2 = 1 + ups;
        P.pm = PM_Z3;
        break;
    }
    case M_Z2: {
        // ipred_z2_c (:462-540): p0 = dx, p1 = dy, p2 = upsample above, p3 = upsample left,
        // e0 = top-left of the prepared edge copy, e1 = its left part
        int dy = g_dr_intra_derivative[(ang - 90) >> 1];
        int dx = g_dr_intra_derivative[(180 - ang) >> 1];
        const int ups_l = ef ? use_upsample(w + h, 180 - ang, is_sm) : 0;
        const int ups_a = ef ? use_upsample(w + h, ang - 90, is_sm) : 0;
        pixel *tl = scratch + z2_centre;     // 2h entries below, 2w + 1 above
        if (ups_a) {
            edge_upsample<pixel>(g, tl, w + 1, edge, 0, w + 1, bdmax);
            dx <<= 1;
        } else {
            const int fs = ef ? filter_strength(w + h, ang - 90, is_sm) : 0;
            if (fs) edge_filter<pixel>(g, tl + 1, w, 0, max_w, edge + 1, -1, w, fs);
            else for (int i = g.gl; i < w; i += g.G) tl[1 + i] = edge[1 + i];
        }
        if (ups_l) {
            edge_upsample<pixel>(g, tl - h * 2, h + 1, edge - h, 0, h + 1, bdmax);
            dy <<= 1;
        } else {
            const int fs = ef ? filter_strength(w + h, 180 - ang, is_sm) : 0;
            if (fs) edge_filter<pixel>(g, tl - h, h, h - max_h, h, edge - h, 0, h + 1, fs);
            else for (int i = g.gl; i < h; i += g.G) tl[-h + i] = edge[-h + i];
        }
        grp_sync(g);
        if (g.gl == 0) tl[0] = edge[0];
        P.e0 = tl; P.e1 = tl - (1 + ups_l);
        P.p0 = dx; P.p1 = dy; P.p2 = ups_a; P.p3 = ups_l;
        P.pm = PM_Z2;
        break;
    }
    default: {
        // filter-intra (ipred_filter_c :618-655): 4x2 sub-blocks on anti-diagonals, each from its
        // left / top neighbours' OUTPUT; predicted into the tile
        const int8_t *taps = g_filter_intra_taps + (ang & 511) * 64;
        const int nbx = w >> 2, nby = h >> 1;
        for (int dg = 0; dg < nbx + nby - 1; dg++) {
            const int by_lo = imax(0, dg - (nbx - 1)), by_hi = imin(dg, nby - 1);
            const int cnt = (by_hi - by_lo + 1) * 8;
            for (int i = g.gl; i < cnt; i += g.G) {
                const int by = by_lo + (i >> 3), bx = dg - by, o = i & 7;
                const int x = bx * 4, y = by * 2;
                int p[7];
                // p0 = (x-1, y-1), p1..p4 = (x..x+3, y-1), p5 = (x-1, y), p6 = (x-1, y+1)
                if (y == 0) {
                    p[0] = edge[x];
#pragma unroll
                    for (int k = 0; k < 4; k++) p[1 + k] = edge[1 + x + k];
                } else {
                    p[0] = x == 0 ? edge[-y] : tile[(y - 1) * w + x - 1];
#pragma unroll
                    for (int k = 0; k < 4; k++) p[1 + k] = tile[(y - 1) * w + x + k];
                }
                if (x == 0) { p[5] = edge[-(1 + y)]; p[6] = edge[-(2 + y)]; }
                else { p[5] = tile[y * w + x - 1]; p[6] = tile[(y + 1) * w + x - 1]; }
                const int8_t *f = taps + o * 8;
                int acc = 0;
#pragma unroll
                for (int k = 0; k < 7; k++) acc += f[k] * p[k];
                tile[(y + (o >> 2)) * w + x + (o & 3)] = (tile_t)clip_px<pixel>((acc + 8) >> 4, bdmax);
            }
            grp_sync(g);
        }
        P.pm = PM_TILE;
        break;
    }
    }
    grp_sync(g);
    return P;
}

// cfl_pred (ipred_tmpl.c:71-84) as a pixel mode: dc of `dc_mode`, p1 = alpha, tile = ac
template <typename pixel>
DEV PixParams<pixel> cfl_setup(const Grp &g, const int dc_mode, const pixel *edge, const int w, const int h,
                               const int16_t *ac, const int alpha, const int bdmax)
{
    PixParams<pixel> P;
    P.pm = PM_CFL; P.p0 = ipred_dc_value<pixel>(g, dc_mode, edge, w, h, bdmax); P.p1 = alpha; P.p2 = P.p3 = 0;
    P.edge = edge; P.e0 = edge; P.e1 = edge; P.tile = ac; P.w = w; P.h = h;
    return P;
}

// pixel (x, y) = index i = y * w + x of the block
template <typename pixel>
DEV int ipred_pixel(const PixParams<pixel> &P, const int x, const int y, const int i, const int bdmax) {
    const pixel *edge = P.edge;
    int pm = P.pm;
    asm volatile("" : "+r"(pm));               // keep ONE loop with the switch inside, not one loop per mode
    switch (pm) {
    case PM_CONST: return P.p0;
    case PM_V: return edge[1 + x];
    case PM_H: return edge[-(1 + y)];
    case PM_PAETH: {
        const int left = edge[-(y + 1)], top = edge[1 + x], tl = P.p0;
        const int bs = left + top - tl;
        const int ld = iabs(left - bs), td = iabs(top - bs), tld = iabs(tl - bs);
        return ld <= td && ld <= tld ? left : td <= tld ? top : tl;
    }
    case PM_SMOOTH: {
        const int wv = g_sm_weights[P.h + y], wh = g_sm_weights[P.w + x];
        return (wv * edge[1 + x] + (256 - wv) * P.p1 + wh * edge[-(1 + y)] + (256 - wh) * P.p0 + 256) >> 9;
    }
    case PM_SMOOTH_V: {
        const int wv = g_sm_weights[P.h + y];
        return (wv * edge[1 + x] + (256 - wv) * P.p1 + 128) >> 8;
    }
    case PM_SMOOTH_H: {
        const int wh = g_sm_weights[P.w + x];
        return (wh * edge[-(y + 1)] + (256 - wh) * P.p0 + 128) >> 8;
    }
    case PM_Z1: {
        const int xpos = (y + 1) * P.p0, frac = xpos & 0x3E;
        const int bx = (xpos >> 6) + x * P.p2;
        return bx < P.p1 ? (P.e0[bx] * (64 - frac) + P.e0[bx + 1] * frac + 32) >> 6 : (int)P.e0[P.p1];
    }
    case PM_Z3: {
        const int ypos = (x + 1) * P.p0, frac = ypos & 0x3E;
        const int by = (ypos >> 6) + y * P.p2;
        return by < P.p1 ? (P.e0[-by] * (64 - frac) + P.e0[-(by + 1)] * frac + 32) >> 6 : (int)P.e0[-P.p1];
    }
    case PM_Z2: {
        const int xpos = ((1 + P.p2) << 6) - P.p0 * (y + 1);
        const int base_x = (xpos >> 6) + x * (1 + P.p2);
        int v;
        if (base_x >= 0) {
            const int fx = xpos & 0x3E;
            v = P.e0[base_x] * (64 - fx) + P.e0[base_x + 1] * fx;
        } else {
            const int ypos = (y << (6 + P.p3)) - P.p1 * (x + 1);
            const int base_y = ypos >> 6, fy = ypos & 0x3E;
            v = P.e1[-base_y] * (64 - fy) + P.e1[-(base_y + 1)] * fy;
        }
        return (v + 32) >> 6;
    }
    case PM_TILE: return ((const pixel *)P.tile)[i];
    case PM_PAL: {                              // pal_pred (ipred_tmpl.c:717-730): tile = packed indices, e0 = palette
        const int q = ((const uint8_t *)P.tile)[i >> 1];
        return P.e0[(x & 1) ? q >> 4 : q & 7];
    }
    default: {                                  // PM_CFL
        const int diff = P.p1 * ((const int16_t *)P.tile)[i];
        const int mg = (iabs(diff) + 32) >> 6;
        return clip_px<pixel>(P.p0 + (diff < 0 ? -mg : mg), bdmax);
    }
    }
}

// PW pixels of row y starting at x (index i = y * w + x): the predictor switch once per segment, a
// straight-line body per pixel.  tile_t: sample type of a PM_TILE tile.
template <typename pixel, int PW, typename tile_t>
DEV void ipred_seg(const PixParams<pixel> &P, const int x, const int y, const int i, const int bdmax, int *v) {
    const pixel *edge = P.edge;
    switch (P.pm) {
    case PM_CONST:
#pragma unroll
        for (int k = 0; k < PW; k++) v[k] = P.p0;
        break;
    case PM_V:
#pragma unroll
        for (int k = 0; k < PW; k++) v[k] = edge[1 + x + k];
        break;
    case PM_H: {
        const int l = edge[-(1 + y)];
#pragma unroll
        for (int k = 0; k < PW; k++) v[k] = l;
        break;
    }
    case PM_PAETH: {
        const int left = edge[-(y + 1)], tl = P.p0;
#pragma unroll
        for (int k = 0; k < PW; k++) {
            const int top = edge[1 + x + k];
            const int bs = left + top - tl;
            const int ld = iabs(left - bs), td = iabs(top - bs), tld = iabs(tl - bs);
            v[k] = ld <= td && ld <= tld ? left : td <= tld ? top : tl;
        }
        break;
    }
    case PM_SMOOTH: {
        const int wv = g_sm_weights[P.h + y], left = edge[-(1 + y)];
        const int vert = (256 - wv) * P.p1 + 256;
#pragma unroll
        for (int k = 0; k < PW; k++) {
            const int wh = g_sm_weights[P.w + x + k];
            v[k] = (wv * edge[1 + x + k] + vert + wh * left + (256 - wh) * P.p0) >> 9;
        }
        break;
    }
    case PM_SMOOTH_V: {
        const int wv = g_sm_weights[P.h + y];
#pragma unroll
        for (int k = 0; k < PW; k++) v[k] = (wv * edge[1 + x + k] + (256 - wv) * P.p1 + 128) >> 8;
        break;
    }
    case PM_SMOOTH_H: {
        const int left = edge[-(y + 1)];
#pragma unroll
        for (int k = 0; k < PW; k++) {
            const int wh = g_sm_weights[P.w + x + k];
            v[k] = (wh * left + (256 - wh) * P.p0 + 128) >> 8;
        }
        break;
    }
    case PM_Z1: {
        const int xpos = (y + 1) * P.p0, frac = xpos & 0x3E;
        const int b0 = (xpos >> 6) + x * P.p2;
#pragma unroll
        for (int k = 0; k < PW; k++) {
            const int bx = b0 + k * P.p2;
            v[k] = bx < P.p1 ? (P.e0[bx] * (64 - frac) + P.e0[bx + 1] * frac + 32) >> 6 : (int)P.e0[P.p1];
        }
        break;
    }
    case PM_Z3: {
#pragma unroll
        for (int k = 0; k < PW; k++) {
            const int ypos = (x + k + 1) * P.p0, frac = ypos & 0x3E;
            const int by = (ypos >> 6) + y * P.p2;
            v[k] = by < P.p1 ? (P.e0[-by] * (64 - frac) + P.e0[-(by + 1)] * frac + 32) >> 6 : (int)P.e0[-P.p1];
        }
        break;
    }
    case PM_Z2: {
        const int xpos = ((1 + P.p2) << 6) - P.p0 * (y + 1), fx = xpos & 0x3E;
#pragma unroll
        for (int k = 0; k < PW; k++) {
            const int base_x = (xpos >> 6) + (x + k) * (1 + P.p2);
            int t;
            if (base_x >= 0) {
                t = P.e0[base_x] * (64 - fx) + P.e0[base_x + 1] * fx;
            } else {
                const int ypos = (y << (6 + P.p3)) - P.p1 * (x + k + 1);
                const int base_y = ypos >> 6, fy = ypos & 0x3E;
                t = P.e1[-base_y] * (64 - fy) + P.e1[-(base_y + 1)] * fy;
            }
            v[k] = (t + 32) >> 6;
        }
        break;
    }
    case PM_TILE:
#pragma unroll
        for (int k = 0; k < PW; k++) v[k] = ((const tile_t *)P.tile)[i + k];
        break;
    case PM_PAL: {                              // pal_pred (ipred_tmpl.c:717-730): tile = packed indices, e0 = palette
#pragma unroll
        for (int k = 0; k < PW; k += 2) {
            const int q = ((const uint8_t *)P.tile)[(i + k) >> 1];
            v[k] = P.e0[q & 7];
            v[k + 1] = P.e0[q >> 4];
        }
        break;
    }
    default:                                    // PM_CFL
#pragma unroll
        for (int k = 0; k < PW; k++) {
            const int diff = P.p1 * ((const int16_t *)P.tile)[i + k];
            const int mg = (iabs(diff) + 32) >> 6;
            v[k] = clip_px<pixel>(P.p0 + (diff < 0 ? -mg : mg), bdmax);
        }
        break;
    }
}

// cfl_ac (ipred_tmpl.c:657-703): ac[w*h] dense from the reconstructed luma
template <typename pixel>
DEV void cfl_ac_block(const Grp &g, int16_t *ac, const pixel *ypx, const int ystride, const int w_pad,
                      const int h_pad, const int w, const int h, const int ss_hor, const int ss_ver)
{
    const int vw = w - 4 * w_pad, vh = h - 4 * h_pad;
    const int lw = 31 - __clz(w);
    int part = 0;
    for (int i = g.gl; i < w * h; i += g.G) {
        const int y = imin(i >> lw, vh - 1), x = imin(i & (w - 1), vw - 1);
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
    const int sum = (grp_sum(g, part) + ((1 << log2sz) >> 1)) >> log2sz;
    grp_sync(g);
    for (int i = g.gl; i < w * h; i += g.G) ac[i] = (int16_t)(ac[i] - sum);
    grp_sync(g);
}

// ------------------------------------------------------------------ edge preparation
// dav1d_prepare_intra_edges (ipred_prepare_tmpl.c:77-204).  x, y, w, h, tw, th in 4-pixel units;
// `dst` = the block's top-left in the frame; `top_sb_edge` = optional pre-filter row backup (may
// be null).  Writes edge[-2*th*4 .. 2*tw*4] as needed and returns the DSP mode index; *angle:
// in = angle_delta, out = absolute angle.  One loop over the edge entries the mode needs
// (left, bottom-left, top, top-right: each entry is one pixel of the frame or a constant).
template <typename pixel>
DEV int prepare_edges(const Grp &g, const int x, const int have_left, const int y, const int have_top, const int w,
                      const int h, const int edge_flags, const pixel *dst, const int stride,
                      const pixel *top_sb_edge, int mode, int *angle, const int tw, const int th,
                      const int filter_edge_flag, pixel *edge, const int bdmax)
{
    const int bitdepth = PxTraits<pixel>::bitdepth(bdmax);
    mode = ipred_resolve_mode(mode, angle, have_left, have_top);
    const int needs = ipred_mode_needs(mode);
    const pixel *dst_top = nullptr;
    if (have_top && ((needs & 2) || (needs & 4) || ((needs & 1) && !have_left)))
        dst_top = top_sb_edge ? top_sb_edge + x * 4 : dst - stride;
    const int mid = (1 << bitdepth) >> 1;
    const int szl = th << 2, szt = tw << 2;
    const int pxl = imin(szl, (h - y) << 2);
    const int have_bl = (needs & 16) && have_left && y + th < h && (edge_flags & 8);
    const int pxbl = imin(szl, (h - y - th) << 2);
    const int pxt = imin(szt, (w - x) << 2);
    const int have_tr = (needs & 8) && have_top && x + tw < w && (edge_flags & 1);
    const int pxtr = imin(szt, (w - x - tw) << 2);
    // fallbacks when a side is unavailable (ipred_prepare_tmpl.c:139-197): the first pixel of the
    // other side, or mid +- 1
    // the corner first: its load overlaps the others
    int vc = mid;
    if ((needs & 4) && g.gl == 0) {
        if (have_left) vc = have_top ? __ldcg(dst_top - 1) : __ldcg(dst - 1);
        else vc = have_top ? (int)__ldcg(dst_top) : mid;
    }
    // entries: [0, nl) left + bottom-left going down, [nl, nl + nt) top + top-right; four entries
    // per lane and step, loaded before any of them is stored (independent loads in flight together)
    const int nl = (needs & 16) ? 2 * szl : (needs & 1) ? szl : 0;
    const int nt = (needs & 8) ? 2 * szt : (needs & 2) ? szt : 0;
    for (int e0 = g.gl; e0 < nl + nt; e0 += 4 * g.G) {
        int v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + u * g.G;
            v[u] = 0;
            if (e < nl) {
                if (!have_left) v[u] = have_top ? (int)__ldcg(dst_top) : mid + 1;
                else if (e < szl) v[u] = __ldcg(dst + stride * imin(e, pxl - 1) - 1);
                else if (have_bl) v[u] = __ldcg(dst + (szl + imin(e - szl, pxbl - 1)) * stride - 1);
                else v[u] = __ldcg(dst + stride * (pxl - 1) - 1);
            } else if (e < nl + nt) {
                const int i = e - nl;
                if (!have_top) v[u] = have_left ? (int)__ldcg(dst - 1) : mid - 1;
                else if (i < szt) v[u] = __ldcg(dst_top + imin(i, pxt - 1));
                else if (have_tr) v[u] = __ldcg(dst_top + szt + imin(i - szt, pxtr - 1));
                else v[u] = __ldcg(dst_top + pxt - 1);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + u * g.G;
            if (e < nl) edge[-1 - e] = (pixel)v[u];
            else if (e < nl + nt) edge[1 + (e - nl)] = (pixel)v[u];
        }
    }
    grp_sync(g);
    if ((needs & 4) && g.gl == 0) {
        // Z2 corner smoothing (ipred_prepare_tmpl.c:198-200)
        if (mode == M_Z2 && tw + th >= 6 && filter_edge_flag) vc = ((edge[-1] + edge[1]) * 5 + vc * 6 + 8) >> 4;
        edge[0] = (pixel)vc;
    }
    grp_sync(g);
    return mode;
}

}  // namespace d1
