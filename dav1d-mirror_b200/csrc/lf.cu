// Deblocking loop filter of a whole frame on the device (first of the in-loop post-filters).
//
// Reference: dav1d_loopfilter_sbrow_cols / _rows (src/lf_apply_tmpl.c:306-466) walking the frame's
// Av1Filter masks (src/lf_mask.h:52-58) and level cache, calling loop_filter_sb[plane][dir]
// (src/loopfilter_tmpl.c:141-246), whose core is loop_filter() (:36-139).
//
// The reference filters superblock row by superblock row: column edges of the row, then its row edges.
// Every edge's filter reaches at most as far as the smaller of the two transforms it separates
// (lf_mask.c: the mask index is min(tx of this side, tx of the neighbour)), so the supports of two edges
// of one direction never overlap, and a superblock row's row-edge pass never touches a pixel a later
// column-edge pass reads.  The frame is therefore TWO data-parallel passes: all column edges (one thread
// per edge and pixel line), then all row edges (one thread per edge and pixel column) - no wavefront.
// Inputs are dav1d's own structures (device copies): f->lf.mask, f->lf.level, f->lf.lim_lut.
#include "ctx.h"
#include "common.cuh"

namespace d1 {

constexpr int AV1FILTER_BYTES = 1348;     // sizeof(Av1Filter): filter_y 768, filter_uv 512, cdef_idx 4, noskip_mask 64
constexpr int AV1FILTER_UV_OFF = 768;

struct LfArgs {
    PlaneView p[3];
    int bdmax, ss_hor, ss_ver;
    int w4, h4, b4_stride, sb128w;
    const uint8_t *masks;
    const uint8_t *level;          // [b4_stride * rows][4]
    uint8_t e[64], i[64];
};

// one pixel line across an edge: loop_filter() for i = one of its four lines.  `s` = distance between the
// filter's taps (1 across a column edge, the stride across a row edge)
template <typename pixel>
DEV void lf_line(pixel *const dst, const int64_t s, const int wd, int E, int I, int H, const int bdmax) {
    const int sh = PxTraits<pixel>::bitdepth(bdmax) - 8;
    const int F = 1 << sh;
    E <<= sh; I <<= sh; H <<= sh;
    const int p1 = dst[-2 * s], p0 = dst[-s], q0 = dst[0], q1 = dst[s];
    int p2 = 0, p3 = 0, q2 = 0, q3 = 0;
    bool fm = iabs(p1 - p0) <= I && iabs(q1 - q0) <= I && iabs(p0 - q0) * 2 + (iabs(p1 - q1) >> 1) <= E;
    if (wd > 4) {
        p2 = dst[-3 * s]; q2 = dst[2 * s];
        fm = fm && iabs(p2 - p1) <= I && iabs(q2 - q1) <= I;
        if (wd > 6) {
            p3 = dst[-4 * s]; q3 = dst[3 * s];
            fm = fm && iabs(p3 - p2) <= I && iabs(q3 - q2) <= I;
        }
    }
    if (!fm) return;
    bool flat_in = false;
    if (wd >= 6) flat_in = iabs(p2 - p0) <= F && iabs(p1 - p0) <= F && iabs(q1 - q0) <= F && iabs(q2 - q0) <= F;
    if (wd >= 8) flat_in = flat_in && iabs(p3 - p0) <= F && iabs(q3 - q0) <= F;
    if (wd >= 16 && flat_in) {
        const int p6 = dst[-7 * s], p5 = dst[-6 * s], p4 = dst[-5 * s];
        const int q4 = dst[4 * s], q5 = dst[5 * s], q6 = dst[6 * s];
        const bool flat_out = iabs(p6 - p0) <= F && iabs(p5 - p0) <= F && iabs(p4 - p0) <= F &&
                              iabs(q4 - q0) <= F && iabs(q5 - q0) <= F && iabs(q6 - q0) <= F;
        if (flat_out) {
            // the 16-weight smoothing kernel [1 1 1 1 1 2 2 2 1 1 1 1 1] over p6 .. q6 with the ends repeated
            // (loopfilter_tmpl.c:89-112), as a running sum: output c (v[c] is its centre) follows from output
            // c - 1 by moving the 13-wide window and the three doubled taps one to the right
            const int v[14] = { p6, p5, p4, p3, p2, p1, p0, q0, q1, q2, q3, q4, q5, q6 };
            int sum = p6 * 7 + p5 * 2 + p4 * 2 + p3 + p2 + p1 + p0 + q0;
            dst[-6 * s] = (pixel)((sum + 8) >> 4);
#pragma unroll
            for (int c = 2; c <= 12; c++) {
                sum += v[c + 6 > 13 ? 13 : c + 6] - v[c - 7 < 0 ? 0 : c - 7] + v[c + 1] - v[c - 2];
                dst[(c - 7) * s] = (pixel)((sum + 8) >> 4);
            }
            return;
        }
    }
    if (wd >= 8 && flat_in) {
        dst[-3 * s] = (pixel)((p3 + p3 + p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3);
        dst[-2 * s] = (pixel)((p3 + p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3);
        dst[-1 * s] = (pixel)((p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3);
        dst[0]      = (pixel)((p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3);
        dst[1 * s]  = (pixel)((p1 + p0 + q0 + 2 * q1 + q2 + q3 + q3 + 4) >> 3);
        dst[2 * s]  = (pixel)((p0 + q0 + q1 + 2 * q2 + q3 + q3 + q3 + 4) >> 3);
        return;
    }
    if (wd == 6 && flat_in) {
        dst[-2 * s] = (pixel)((p2 + 2 * p2 + 2 * p1 + 2 * p0 + q0 + 4) >> 3);
        dst[-1 * s] = (pixel)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
        dst[0]      = (pixel)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
        dst[1 * s]  = (pixel)((p0 + 2 * q0 + 2 * q1 + 2 * q2 + q2 + 4) >> 3);
        return;
    }
    // narrow filter, with or without high edge variance
    const int lo = -128 * (1 << sh), hi = 128 * (1 << sh) - 1;
    const bool hev = iabs(p1 - p0) > H || iabs(q1 - q0) > H;
    int f = hev ? iclip(p1 - q1, lo, hi) : 0;
    f = iclip(3 * (q0 - p0) + f, lo, hi);
    const int f1 = imin(f + 4, hi) >> 3, f2 = imin(f + 3, hi) >> 3;
    dst[-s] = (pixel)clip_px<pixel>(p0 + f2, bdmax);
    dst[0] = (pixel)clip_px<pixel>(q0 - f1, bdmax);
    if (!hev) {
        const int g = (f1 + 1) >> 1;
        dst[-2 * s] = (pixel)clip_px<pixel>(p1 + g, bdmax);
        dst[s] = (pixel)clip_px<pixel>(q1 - g, bdmax);
    }
}

// DIR 0: column edges (filter_*[0], taps along x); DIR 1: row edges (filter_*[1], taps along y).
// blockIdx.z = plane (luma launch: 0; chroma launch: 1, 2).
template <typename pixel, int DIR>
__global__ void __launch_bounds__(256) lf_pass_kernel(const __grid_constant__ LfArgs a, const int first_plane) {
    // DIR 0: a thread per (edge column x4, pixel row); DIR 1: a thread per (pixel column, edge row y4) - no
    // divisions, consecutive lanes on consecutive edges / pixels of one row
    const int pl = first_plane + blockIdx.z;
    const int sh = pl ? a.ss_hor : 0, sv = pl ? a.ss_ver : 0;
    const int pw4 = (a.w4 + sh) >> sh, ph4 = (a.h4 + sv) >> sv;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y;
    int x4, y4, line;
    if (DIR == 0) { x4 = gx; y4 = gy >> 2; line = gy & 3; }
    else { x4 = gx >> 2; line = gx & 3; y4 = gy; }
    if (x4 >= pw4 || y4 >= ph4) return;
    if ((DIR == 0 ? x4 : y4) == 0) return;                      // no edge at the frame's left / top border
    // the 128x128 area and the position inside it, in this plane's 4-px units
    const int lx = 5 - sh, ly = 5 - sv;
    const uint8_t *F = a.masks + (size_t)((y4 >> ly) * a.sb128w + (x4 >> lx)) * AV1FILTER_BYTES;
    const int ax = x4 & ((1 << lx) - 1), ay = y4 & ((1 << ly) - 1);
    const int along = DIR == 0 ? ax : ay;                       // index of the edge line
    const int across = DIR == 0 ? ay : ax;                      // bit inside the line's mask
    const int half = 16 >> (DIR == 0 ? sv : sh);                // bits per uint16_t half
    const int sidx = across >= half, bit = across - sidx * half;
    int idx = -1;
    if (pl == 0) {
        const uint16_t *m = (const uint16_t *)F + ((DIR * 32 + along) * 3) * 2 + sidx;
        idx = ((m[4] >> bit) & 1) ? 2 : ((m[2] >> bit) & 1) ? 1 : ((m[0] >> bit) & 1) ? 0 : -1;
    } else {
        const uint16_t *m = (const uint16_t *)(F + AV1FILTER_UV_OFF) + ((DIR * 32 + along) * 2) * 2 + sidx;
        idx = ((m[2] >> bit) & 1) ? 1 : ((m[0] >> bit) & 1) ? 0 : -1;
    }
    if (idx < 0) return;
    // level of this side, else of the other side (loopfilter_tmpl.c:153, 175)
    const int comp = pl == 0 ? DIR : 1 + pl;
    const uint8_t *l = a.level + ((size_t)y4 * a.b4_stride + x4) * 4 + comp;
    int L = l[0];
    if (!L) L = DIR == 0 ? l[-4] : l[-(int64_t)a.b4_stride * 4];
    if (!L) return;
    const int wd = pl == 0 ? 4 << idx : 4 + 2 * idx;
    const PlaneView &pv = a.p[pl];
    const int64_t stride = pv.stride / (int64_t)sizeof(pixel);
    pixel *dst = (pixel *)pv.data + (DIR == 0 ? (int64_t)(y4 * 4 + line) * stride + x4 * 4
                                              : (int64_t)(y4 * 4) * stride + x4 * 4 + line);
    lf_line<pixel>(dst, DIR == 0 ? 1 : stride, wd, a.e[L], a.i[L], L >> 4, a.bdmax);
}

}  // namespace d1

using namespace d1;

extern "C" int dav1d_cuda_loopfilter_frame(Dav1dCudaContext *c, const Dav1dCudaPicture *pic, const Dav1dCudaLfFrame *lf) {
    if (!c || !pic || !lf || !pic->p[0].data || !lf->masks || !lf->level || lf->w4 <= 0 || lf->h4 <= 0 ||
        lf->b4_stride < lf->w4 || lf->sb128w < (lf->w4 + 31) / 32) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    LfArgs a;
    const PicView v = pic_view(pic);
    for (int i = 0; i < 3; i++) a.p[i] = v.p[i];
    a.bdmax = v.bdmax; a.ss_hor = v.ss_hor; a.ss_ver = v.ss_ver;
    a.w4 = lf->w4; a.h4 = lf->h4; a.b4_stride = lf->b4_stride; a.sb128w = lf->sb128w;
    a.masks = (const uint8_t *)lf->masks; a.level = lf->level;
    memcpy(a.e, lf->lut_e, 64); memcpy(a.i, lf->lut_i, 64);
    const bool chroma = lf->filter_uv && pic->p[1].data && pic->p[2].data;
    const bool hbd = v.bdmax > 0xff;
    const int cw4 = (lf->w4 + v.ss_hor) >> v.ss_hor, ch4 = (lf->h4 + v.ss_ver) >> v.ss_ver;
    for (int dir = 0; dir < 2; dir++) {
        for (int ch = 0; ch < (chroma ? 2 : 1); ch++) {
            const int pw4 = ch ? cw4 : lf->w4, ph4 = ch ? ch4 : lf->h4;
            const dim3 grid(dir ? (unsigned)((pw4 * 4 + 127) / 128) : (unsigned)((pw4 + 127) / 128),
                            dir ? (unsigned)ph4 : (unsigned)(ph4 * 4), ch ? 2 : 1);
            if (hbd) {
                if (dir) lf_pass_kernel<uint16_t, 1><<<grid, 128, 0, c->stream>>>(a, ch);
                else lf_pass_kernel<uint16_t, 0><<<grid, 128, 0, c->stream>>>(a, ch);
            } else {
                if (dir) lf_pass_kernel<uint8_t, 1><<<grid, 128, 0, c->stream>>>(a, ch);
                else lf_pass_kernel<uint8_t, 0><<<grid, 128, 0, c->stream>>>(a, ch);
            }
            count_launch();
        }
    }
    return cuda_ok(cudaGetLastError(), "lf_pass_kernel") ? 0 : -5;
}
