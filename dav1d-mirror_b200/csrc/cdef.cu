// CDEF (constrained directional enhancement filter) of a whole frame on the device, second of the in-loop
// post-filters.
//
// Reference: dav1d_filter_sbrow_cdef (src/recon_tmpl.c:2073-2100) -> dav1d_cdef_brow
// (src/cdef_apply_tmpl.c:98-309) -> dsp->cdef.dir = cdef_find_dir_c and dsp->cdef.fb[] =
// cdef_filter_block_c (src/cdef_tmpl.c:36-296).
//
// The reference filters in place and therefore keeps copies of every pre-filter pixel a later block still
// reads (two lines above: f->lf.cdef_line; two columns to the left: lr_bak; blocks to the right and below
// are not filtered yet): every tap of every block reads the DEBLOCKED, not yet CDEF-filtered picture.  On
// the device that is simply an out-of-place filter: src = deblocked picture, dst = another picture, one
// warp per 8x8 luma block (+ its chroma blocks), all blocks independent.  Skipped blocks are copied.
#include "ctx.h"
#include "common.cuh"

namespace d1 {

constexpr int AV1FILTER_BYTES = 1348;          // sizeof(Av1Filter), src/lf_mask.h:52-58
constexpr int AV1FILTER_CDEF_IDX = 1280;       // int8_t cdef_idx[4]
constexpr int AV1FILTER_NOSKIP = 1284;         // uint16_t noskip_mask[16][2]
constexpr int CDEF_WARPS = 8;
constexpr int TS = 12;                         // tile stride, as tmp_stride in cdef_filter_block_c
constexpr int UNAVAILABLE = -32768;            // INT16_MIN: huge as unsigned, very negative as signed (cdef_tmpl.c:47-48)

struct CdefArgs {
    PlaneView src[3], dst[3];
    int bdmax, ss_hor, ss_ver, n_planes;
    int bw, bh, sb128w, damping;
    uint8_t y_strength[8], uv_strength[8];
    const uint8_t *masks;
};

// dav1d_cdef_directions (src/tables.c:400-413): tap offsets in a stride-12 tile; entry d + 2 is direction d
__device__ const int8_t g_cdef_directions[12][2] = {
    { 1 * 12 + 0, 2 * 12 + 0 }, { 1 * 12 + 0, 2 * 12 - 1 }, { -1 * 12 + 1, -2 * 12 + 2 }, { 0 * 12 + 1, -1 * 12 + 2 },
    { 0 * 12 + 1, 0 * 12 + 2 }, { 0 * 12 + 1, 1 * 12 + 2 }, { 1 * 12 + 1, 2 * 12 + 2 }, { 1 * 12 + 0, 2 * 12 + 1 },
    { 1 * 12 + 0, 2 * 12 + 0 }, { 1 * 12 + 0, 2 * 12 - 1 }, { -1 * 12 + 1, -2 * 12 + 2 }, { 0 * 12 + 1, -1 * 12 + 2 },
};

DEV int ulog2(const unsigned v) { return 31 - __clz(v); }
DEV int constrain(const int diff, const int threshold, const int shift) {
    const int adiff = iabs(diff);
    const int v = imin(adiff, imax(0, threshold - (adiff >> shift)));
    return diff < 0 ? -v : v;
}
DEV int umin_i(const int a, const int b) { return (unsigned)a < (unsigned)b ? a : b; }

// the (w + 4) x (h + 4) neighbourhood of a block into the warp's tile (padding(), cdef_tmpl.c:57-102)
template <typename pixel, int w, int h>
DEV void cdef_load_tile(int16_t *tile, const PlaneView &pv, const int x0, const int y0, const int edges, const int lane)
{
    const pixel *src = (const pixel *)pv.data;
    const int64_t stride = pv.stride / (int64_t)sizeof(pixel);
    for (int i = lane; i < (w + 4) * (h + 4); i += 32) {
        const int ty = i / (w + 4) - 2, tx = i % (w + 4) - 2;
        const bool ok = (ty >= 0 || (edges & 4)) && (ty < h || (edges & 8)) && (tx >= 0 || (edges & 1)) && (tx < w || (edges & 2));
        tile[(ty + 2) * TS + tx + 2] = ok ? (int16_t)src[(int64_t)(y0 + ty) * stride + x0 + tx] : (int16_t)UNAVAILABLE;
    }
    __syncwarp();
}

// cdef_filter_block_c (cdef_tmpl.c:104-210) over the tile; pri / sec: strengths (either may be 0, not both)
template <typename pixel, int w, int h>
DEV void cdef_filter(const int16_t *tile, const PlaneView &pv, const int x0, const int y0,
                     const int pri, const int sec, const int dir, const int damping, const int bdmax, const int lane)
{
    pixel *dst = (pixel *)pv.data;
    const int64_t stride = pv.stride / (int64_t)sizeof(pixel);
    const int bdm8 = PxTraits<pixel>::bitdepth(bdmax) - 8;
    const int pri_tap = 4 - ((pri >> bdm8) & 1);
    const int pri_shift = pri ? imax(0, damping - ulog2(pri)) : 0;
    const int sec_shift = sec ? damping - ulog2(sec) : 0;
    for (int i = lane; i < w * h; i += 32) {
        const int y = i / w, x = i % w;
        const int16_t *t = tile + (y + 2) * TS + x + 2;
        const int px = t[0];
        int sum = 0, mx = px, mn = px;
        int pri_tap_k = pri_tap;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (pri) {
                const int off = g_cdef_directions[dir + 2][k];
                const int p0 = t[off], p1 = t[-off];
                sum += pri_tap_k * constrain(p0 - px, pri, pri_shift);
                sum += pri_tap_k * constrain(p1 - px, pri, pri_shift);
                pri_tap_k = (pri_tap_k & 3) | 2;            // 4 -> 2, 3 stays 3
                mn = umin_i(p0, mn); mx = imax(p0, mx);
                mn = umin_i(p1, mn); mx = imax(p1, mx);
            }
            if (sec) {
                const int off2 = g_cdef_directions[dir + 4][k], off3 = g_cdef_directions[dir][k];
                const int s0 = t[off2], s1 = t[-off2], s2 = t[off3], s3 = t[-off3];
                const int sec_tap = 2 - k;
                sum += sec_tap * constrain(s0 - px, sec, sec_shift);
                sum += sec_tap * constrain(s1 - px, sec, sec_shift);
                sum += sec_tap * constrain(s2 - px, sec, sec_shift);
                sum += sec_tap * constrain(s3 - px, sec, sec_shift);
                mn = umin_i(s0, mn); mx = imax(s0, mx);
                mn = umin_i(s1, mn); mx = imax(s1, mx);
                mn = umin_i(s2, mn); mx = imax(s2, mx);
                mn = umin_i(s3, mn); mx = imax(s3, mx);
            }
        }
        int v = px + ((sum - (sum < 0) + 8) >> 4);
        if (pri && sec) v = iclip(v, mn, mx);              // only the two-strength form clips (:155)
        dst[(int64_t)(y0 + y) * stride + x0 + x] = (pixel)v;
    }
}

template <typename pixel, int w, int h>
DEV void cdef_copy(const PlaneView &s, const PlaneView &d, const int x0, const int y0, const int lane)
{
    const int64_t ss = s.stride / (int64_t)sizeof(pixel), ds = d.stride / (int64_t)sizeof(pixel);
    for (int i = lane; i < w * h; i += 32) {
        const int y = i / w, x = i % w;
        ((pixel *)d.data)[(int64_t)(y0 + y) * ds + x0 + x] = ((const pixel *)s.data)[(int64_t)(y0 + y) * ss + x0 + x];
    }
}

template <typename pixel, int cw, int ch>
__global__ void __launch_bounds__(CDEF_WARPS * 32) cdef_kernel(const __grid_constant__ CdefArgs a) {
    __shared__ int16_t tiles[CDEF_WARPS][TS * TS];
    __shared__ int partial[CDEF_WARPS][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w8 = a.bw >> 1, h8 = a.bh >> 1;
    const int blk = blockIdx.x * CDEF_WARPS + warp;
    if (blk >= w8 * h8) return;
    const int bx = (blk % w8) * 2, by = (blk / w8) * 2;                     // 4-px units, as in dav1d_cdef_brow
    int16_t *tile = tiles[warp];
    const int bdm8 = PxTraits<pixel>::bitdepth(a.bdmax) - 8;
    const int cx0 = (bx * 4) >> a.ss_hor, cy0 = (by * 4) >> a.ss_ver;

    // strength index of the 64x64 area and the block's skip bit (cdef_apply_tmpl.c:147-183)
    const uint8_t *F = a.masks + (size_t)((by >> 5) * a.sb128w + (bx >> 5)) * AV1FILTER_BYTES;
    const int cdef_idx = (int8_t)F[AV1FILTER_CDEF_IDX + ((by & 16) >> 3) + ((bx & 16) >> 4)];
    const uint16_t *ns = (const uint16_t *)(F + AV1FILTER_NOSKIP) + ((by & 30) >> 1) * 2;
    const unsigned noskip = (unsigned)ns[1] << 16 | ns[0];
    const int y_lvl = cdef_idx >= 0 ? a.y_strength[cdef_idx] : 0, uv_lvl = cdef_idx >= 0 && a.n_planes > 1 ? a.uv_strength[cdef_idx] : 0;
    const bool run = cdef_idx >= 0 && (y_lvl || uv_lvl) && (noskip & (3u << (bx & 30)));
    if (!run) {
        cdef_copy<pixel, 8, 8>(a.src[0], a.dst[0], bx * 4, by * 4, lane);
        for (int pl = 1; pl < a.n_planes; pl++) cdef_copy<pixel, cw, ch>(a.src[pl], a.dst[pl], cx0, cy0, lane);
        return;
    }
    const int edges = (bx > 0 ? 1 : 0) | (bx + 2 < a.bw ? 2 : 0) | (by > 0 ? 4 : 0) | (by + 2 < a.bh ? 8 : 0);
    const int y_pri = (y_lvl >> 2) << bdm8, uv_pri = (uv_lvl >> 2) << bdm8;
    int y_sec = y_lvl & 3, uv_sec = uv_lvl & 3;
    y_sec += y_sec == 3; uv_sec += uv_sec == 3;
    y_sec <<= bdm8; uv_sec <<= bdm8;
    const int damping = a.damping + bdm8;

    cdef_load_tile<pixel, 8, 8>(tile, a.src[0], bx * 4, by * 4, edges, lane);
    int dir = 0;
    unsigned var = 0;
    if (y_pri || uv_pri) {
        // cdef_find_dir_c (cdef_tmpl.c:233-296): per direction the sums along its lines; the cost of a
        // direction is the sum over its lines of sum^2 * 840 / (pixels on the line)
        // 90 lines in all (15 + 11 + 8 + 11 + 15 + 11 + 8 + 11): a lane sums the pixels of a line, squares,
        // weighs; the costs of a direction's lines meet in shared memory
        unsigned *costs = (unsigned *)partial[warp];
        if (lane < 8) costs[lane] = 0;
        __syncwarp();
        for (int g = lane; g < 90; g += 32) {
            const int d = g < 15 ? 0 : g < 26 ? 1 : g < 34 ? 2 : g < 45 ? 3 : g < 60 ? 4 : g < 71 ? 5 : g < 79 ? 6 : 7;
            const int l = g - (d == 0 ? 0 : d == 1 ? 15 : d == 2 ? 26 : d == 3 ? 34 : d == 4 ? 45 : d == 5 ? 60 : d == 6 ? 71 : 79);
            // pixel t of the line: x = ax * t + bx * (t >> 1) + cx, y likewise (the eight line equations of
            // cdef_find_dir_c with t along x for d = 0..4 and along y for d = 5..7)
            const int ax = d < 5, ay = d == 0 ? -1 : d == 4 || d >= 5;
            const int bx = d == 5 ? 1 : d == 7 ? -1 : 0, by = d == 1 ? -1 : d == 3 ? 1 : 0;
            const int cx = d < 5 ? 0 : d == 5 ? l - 3 : l;
            const int cy = d >= 5 ? 0 : d == 3 ? l - 3 : d == 4 ? l - 7 : l;
            int sum = 0, len = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const int x = ax * t + bx * (t >> 1) + cx, y = ay * t + by * (t >> 1) + cy;
                if ((unsigned)x < 8u && (unsigned)y < 8u) {
                    sum += (tile[(y + 2) * TS + x + 2] >> bdm8) - 128;
                    len++;
                }
            }
            atomicAdd(&costs[d], (unsigned)(sum * sum) * (unsigned)(840 / len));
        }
        __syncwarp();
        const unsigned cost = lane < 8 ? costs[lane] : 0;
        unsigned best_cost = __shfl_sync(0xffffffffu, cost, 0);
        for (int n = 1; n < 8; n++) {
            const unsigned c = __shfl_sync(0xffffffffu, cost, n);
            if (c > best_cost) { best_cost = c; dir = n; }
        }
        var = (best_cost - __shfl_sync(0xffffffffu, cost, dir ^ 4)) >> 10;
    }
    // luma (:236-246)
    int pri = 0, sec = 0, ydir = 0;
    if (y_pri) {
        int adj = 0;
        if (var) {
            const int i = (var >> 6) ? imin(ulog2(var >> 6), 12) : 0;
            adj = (y_pri * (4 + i) + 8) >> 4;
        }
        if (adj || y_sec) { pri = adj; sec = y_sec; ydir = dir; }
    } else if (y_sec) {
        sec = y_sec;
    }
    if (pri || sec) cdef_filter<pixel, 8, 8>(tile, a.dst[0], bx * 4, by * 4, pri, sec, ydir, damping, a.bdmax, lane);
    else cdef_copy<pixel, 8, 8>(a.src[0], a.dst[0], bx * 4, by * 4, lane);
    // chroma (:248-285)
    if (a.n_planes > 1) {
        const int uvdir = uv_pri ? (a.ss_hor && !a.ss_ver ? (0x66654207u >> (4 * dir)) & 7 : dir) : 0;   // uv_dirs[4:2:2] = {7,0,2,4,5,6,6,6}
        for (int pl = 1; pl <= 2; pl++) {
            if (!uv_lvl) { cdef_copy<pixel, cw, ch>(a.src[pl], a.dst[pl], cx0, cy0, lane); continue; }
            __syncwarp();
            cdef_load_tile<pixel, cw, ch>(tile, a.src[pl], cx0, cy0, edges, lane);
            cdef_filter<pixel, cw, ch>(tile, a.dst[pl], cx0, cy0, uv_pri, uv_sec, uvdir, damping - 1, a.bdmax, lane);
        }
    }
}

}  // namespace d1

using namespace d1;

extern "C" int dav1d_cuda_cdef_frame(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, const Dav1dCudaPicture *src,
                                     const Dav1dCudaCdefFrame *p)
{
    if (!c || !dst || !src || !p || !p->masks || !dst->p[0].data || !src->p[0].data || dst->p[0].data == src->p[0].data ||
        p->bw <= 0 || p->bh <= 0 || (p->bw & 1) || (p->bh & 1) || p->sb128w < (p->bw + 31) / 32 ||
        p->damping < 3 || p->damping > 6 || dst->bitdepth_max != src->bitdepth_max || dst->ss_hor != src->ss_hor ||
        dst->ss_ver != src->ss_ver) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    CdefArgs a;
    const PicView s = pic_view(src), d = pic_view(dst);
    for (int i = 0; i < 3; i++) { a.src[i] = s.p[i]; a.dst[i] = d.p[i]; }
    a.bdmax = s.bdmax; a.ss_hor = s.ss_hor; a.ss_ver = s.ss_ver;
    a.n_planes = (src->p[1].data && src->p[2].data && dst->p[1].data && dst->p[2].data) ? 3 : 1;
    a.bw = p->bw; a.bh = p->bh; a.sb128w = p->sb128w; a.damping = p->damping;
    memcpy(a.y_strength, p->y_strength, 8); memcpy(a.uv_strength, p->uv_strength, 8);
    a.masks = (const uint8_t *)p->masks;
    const int n = (p->bw >> 1) * (p->bh >> 1);
    const int grid = (n + CDEF_WARPS - 1) / CDEF_WARPS;
    const int lay = a.n_planes == 1 ? 0 : s.ss_hor ? (s.ss_ver ? 2 : 1) : 0;      // chroma block: 8x8, 4x8, 4x4
    const bool hbd = s.bdmax > 0xff;
#define LAUNCH_CDEF(pixel, cw, ch) cdef_kernel<pixel, cw, ch><<<grid, CDEF_WARPS * 32, 0, c->stream>>>(a)
    if (lay == 0) { if (hbd) LAUNCH_CDEF(uint16_t, 8, 8); else LAUNCH_CDEF(uint8_t, 8, 8); }
    else if (lay == 1) { if (hbd) LAUNCH_CDEF(uint16_t, 4, 8); else LAUNCH_CDEF(uint8_t, 4, 8); }
    else { if (hbd) LAUNCH_CDEF(uint16_t, 4, 4); else LAUNCH_CDEF(uint8_t, 4, 4); }
#undef LAUNCH_CDEF
    count_launch();
    return cuda_ok(cudaGetLastError(), "cdef_kernel") ? 0 : -5;
}
