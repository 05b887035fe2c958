// Frame-level batched reconstruction: the intra-class executor and the per-frame / per-group
// submit entry points.
//
// Reference call sites replaced: dav1d_recon_b_intra (src/recon_tmpl.c:1195-1596) and, through the
// MC / ITX launches, dav1d_recon_b_inter (:1598-2036).
//
// Intra executor.  Intra prediction of a transform block reads final pixels of its neighbours
// (recon_tmpl.c:1259-1347), so the intra-class operations of a frame form a dependency DAG.  The
// reference resolves it by decoding superblocks in order (and, across threads, by superblock-row
// progress counters: src/decode.c:2001-2090, src/thread_task.c:409-430).  Here ONE persistent
// cooperative launch per group of frames runs the DAG level by level, and finds the levels itself:
//   * the recorder hands over the operations in decode order - nothing is scheduled, sorted or
//     levelled on the host;
//   * a byte per 4x4 cell and plane counts the operations that still have to write the cell (set
//     up by a small marking launch, back at zero when the frame is done: the map needs no clearing
//     between frames);
//   * round r: (1) every pending operation is looked at by one THREAD: it is ready when the cells
//     of the pixels it reads are at zero (dav1d_prepare_intra_edges' own rules); ready operations
//     go to the round's list with a key (predictor class, size class), the others stay pending;
//     (2) the list is sorted by key (counting sort); (3) warps execute it - four small
//     operations per warp, one per octet of lanes, neighbouring warps on the same predictor - and
//     count the cells down.  Grid barriers separate the steps.  The operations of a round are
//     independent of each other by construction, so nothing ever waits: round r executes exactly
//     dependency level r.
//   * the residuals do not depend on neighbours: a transform pre-pass (itx2.cu, next to the motion
//     compensation) leaves them in an int16 plane and the executor adds them to its predictions.
// A frame whose descriptors are inconsistent (operations that wait for each other) ends with
// pending operations and no ready one: the executor raises the context's status word and stops;
// dav1d_cuda_synchronize() reports it.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "ctx.h"
#include "itx2.cuh"
#include "ipred.cuh"
#include "mc.cuh"

namespace d1 {

// defined in itx2.cu / mc.cu
int itx_batch_launch(const PicView &pic, const PicView *res, void *cf, const Dav1dCudaItxDesc *descs,
                     const int32_t *class_count, int zero_coefs, cudaStream_t st);
int itx_task_launch(const PicView &pic, const PicView *res, void *cf, const Dav1dCudaItxDesc *descs,
                    const uint32_t *tasks, int n_small, int n_big, int zero_coefs, cudaStream_t st_small,
                    cudaStream_t st_big);
int mc_obmc_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcDesc *descs,
                       const uint32_t *tiles, int n_tiles, cudaStream_t st);
void itx_init_attrs();
int mc_put_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcDesc *descs,
                      const uint32_t *tiles, int n_tiles, int n_small, uint8_t *masks, int16_t *tmp,
                      bool compound, cudaStream_t st);

constexpr int EDGE_BUF = 288;
constexpr int EDGE_C = 144;
constexpr int I2_MAXF = DAV1D_CUDA_MAX_GROUP;

// one frame of the group
struct Intra2Frame {
    PicView pic;
    PicView res;                         // int16 residual planes written by the transform pre-pass
    int bw4, bh4;
    void *cf;
    const Dav1dCudaIntraDesc *descs;     // decode order
    const void *pal;
    const uint8_t *pal_idx;
    int n_ops;
    uint8_t *map;                        // cell map: plane 0, 1, 2 one after the other
};
struct Intra2Args {
    Intra2Frame f[I2_MAXF];
    int nf;
    unsigned *status;                    // context status word: bit0 = operations that wait for each other
};

HD int map_w(const Intra2Frame &f, const int pl) { return pl ? (f.bw4 + f.pic.ss_hor) >> f.pic.ss_hor : f.bw4; }
HD int map_h(const Intra2Frame &f, const int pl) { return pl ? (f.bh4 + f.pic.ss_ver) >> f.pic.ss_ver : f.bh4; }
HD int map_off(const Intra2Frame &f, const int pl) {
    const int s0 = f.bw4 * f.bh4, s1 = map_w(f, 1) * map_h(f, 1);
    return pl == 0 ? 0 : pl == 1 ? s0 : s0 + s1;
}

// Edges the resolved predictor needs: bit0 left, bit1 top, bit2 topleft, bit3 topright,
// bit4 bottomleft (mirror of the table in ipred.cuh prepare_edges(); ipred_prepare_tmpl.c:50-74,94-117)
HD int intra_needs(const int mode, const int angle_delta, const int have_left, const int have_top) {
    if (mode >= 1 && mode <= 8) {
        const int base = mode == 1 ? 90 : mode == 2 ? 180 : mode == 3 ? 45 : mode == 4 ? 135 : mode == 5 ? 113
                       : mode == 6 ? 157 : mode == 7 ? 203 : 67;
        const int a = base + 3 * angle_delta;
        if (a <= 90) return (a < 90 && have_top) ? (2 | 8 | 4) : 2;             // Z1 : VERT
        if (a < 180) return 1 | 2 | 4;                                          // Z2
        return (a > 180 && have_left) ? (1 | 16 | 4) : 1;                       // Z3 : HOR
    }
    if (mode == 0) return have_left ? (have_top ? 3 : 1) : (have_top ? 2 : 0);  // DC family
    if (mode == 12) return have_left ? (have_top ? 7 : 1) : (have_top ? 2 : 0); // PAETH -> HOR / VERT / DC_128
    if (mode >= 9 && mode <= 11) return 3;                                      // SMOOTH*
    return 1 | 2 | 4;                                                           // FILTER
}

DEV unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Grid barrier of the cooperative launch: a monotonic arrival counter (zeroed before the launch).
DEV void grid_barrier(unsigned *bar, unsigned &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acquire_u32(bar) < target) __nanosleep(64);
        __threadfence();
    }
    __syncthreads();
}

// ---- step 1 of a round: is operation d ready?  One thread, L2 loads of the map bytes.
// all cells of [x0, x1) x [y0, y1) at most `thr`?  else *blocker = map offset of one that is not.
// Eight loads are in flight at a time (a thread that checks the 48 neighbour cells of a 64x64
// block one after the other would hold its whole block up at the end of the step).
DEV bool cells_le(const uint8_t *map, const unsigned mo, const int W, const int x0, const int x1, const int y0,
                  const int y1, const unsigned thr, unsigned *blocker) {
    const int nx = x1 - x0, n = nx * (y1 - y0);
    if (n <= 0) return true;
    for (int j0 = 0; j0 < n; j0 += 8) {
        unsigned v[8], off[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int j = imin(j0 + u, n - 1);
            const int y = y1 - y0 == 1 ? y0 : nx == 1 ? y0 + j : y0 + j / nx;
            const int x = y1 - y0 == 1 ? x0 + j : nx == 1 ? x0 : x0 + j % nx;
            off[u] = mo + (unsigned)(y * W + x);
            v[u] = __ldcg(map + off[u]);
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (v[u] > thr) { *blocker = off[u] | (thr << 31); return false; }
    }
    return true;
}

// On false *blocker names one cell the operation waits for (bit 31: the cell may stay at 1 - the
// operation's own count): while that cell is above its threshold there is no need to look again.
DEV bool op_ready(const Intra2Frame &f, const Dav1dCudaIntraDesc &d, int *cls_out, unsigned *blocker) {
    const int mode = d.mode;
    const int pl = d.plane;
    const int sh = pl ? f.pic.ss_hor : 0, sv = pl ? f.pic.ss_ver : 0;
    const int W = map_w(f, pl), H = map_h(f, pl);
    const unsigned mo = (unsigned)map_off(f, pl);
    const int x0 = d.x4, y0 = d.y4;
    if (mode == DAV1D_CUDA_INTRA_PAL) { *cls_out = 15; return true; }
    if (mode == DAV1D_CUDA_INTRA_NONE) {
        // residual on top of an earlier operation's pixels (palette, inter-intra, intrabc): that one is done
        *cls_out = 18;
        return cells_le(f.map, mo, W, x0, imin(x0 + d.tw4, W), y0, imin(y0 + d.th4, H), 1, blocker);
    }
    if (mode == DAV1D_CUDA_INTRA_IBC) {
        *cls_out = 16;
        const int sx = (int16_t)(d.aux & 0xffff), sy = (int16_t)(d.aux >> 16);
        const int pw = 4 * W, ph = 4 * H;
        const int xa = iclip(sx, 0, pw - 1) >> 2, xb = iclip(sx + 4 * d.tw4 + (d.angle_delta ? 1 : 0) - 1, 0, pw - 1) >> 2;
        const int ya = iclip(sy, 0, ph - 1) >> 2, yb = iclip(sy + 4 * d.th4 + (d.flags ? 1 : 0) - 1, 0, ph - 1) >> 2;
        return cells_le(f.map, mo, W, xa, xb + 1, ya, yb + 1, 0, blocker);
    }
    const int have_left = x0 > d.tile_x4_start, have_top = y0 > d.tile_y4_start;
    int ang = mode == DAV1D_CUDA_INTRA_II ? 0 : d.angle_delta;
    const int rm = ipred_resolve_mode(mode == DAV1D_CUDA_INTRA_II ? d.angle_delta : mode == DAV1D_CUDA_INTRA_CFL ? 0 : mode,
                                      &ang, have_left, have_top);
    *cls_out = mode == DAV1D_CUDA_INTRA_II ? 17 : mode == DAV1D_CUDA_INTRA_CFL ? 14 : rm;
    const int needs = ipred_mode_needs(rm);
    // exactly the pixels dav1d_prepare_intra_edges reads (ipred_prepare_tmpl.c:119-200)
    if (have_top && ((needs & 2) || (needs & 4) || ((needs & 1) && !have_left))) {
        const bool tr = (needs & 8) && (d.edge_flags & 1);
        const int xs = ((needs & 4) && have_left) ? x0 - 1 : x0;
        int xe = (needs & 2) ? imin(x0 + d.tw4 + (tr ? d.tw4 : 0), d.tile_x4_end) : x0 + 1;
        xe = imin(xe, W);
        if (!cells_le(f.map, mo, W, xs, xe, y0 - 1, y0, 0, blocker)) return false;
    }
    if (have_left && ((needs & 1) || ((needs & 2) && !have_top) || ((needs & 4) && !have_top))) {
        const bool bl = (needs & 16) && (d.edge_flags & 8);
        int ye = (needs & 1) ? imin(y0 + d.th4 + (bl ? d.th4 : 0), d.tile_y4_end) : y0 + 1;
        ye = imin(ye, H);
        if (!cells_le(f.map, mo, W, x0 - 1, x0, y0, ye, 0, blocker)) return false;
    }
    if (mode == DAV1D_CUDA_INTRA_CFL)      // the co-located luma
        return cells_le(f.map, 0u, f.bw4, x0 << sh, imin((x0 + d.tw4) << sh, f.bw4), y0 << sv,
                        imin((y0 + d.th4) << sv, f.bh4), 0, blocker);
    return true;
}

// ---- step 3: one operation by a group of lanes (a warp, or an octet for operations of up to 64
// pixels): edge preparation, then ONE loop over the block's pixels (a pixel per lane and step)
// that evaluates the predictor, adds the residual of the transform pre-pass and stores the final
// pixel; then the operation's cells are counted down.
#ifdef D1_EXPERIMENT
__device__ int g_skip;
#define D1_SKIP(bit) (g_skip & (bit))
#else
#define D1_SKIP(bit) 0
#endif
// part / n_parts: the pixel loop of a large operation is shared by several warps (each prepares the
// edge for itself and takes every n_parts-th chunk of 256 pixels; part 0 counts the cells down).
template <typename pixel>
__device__ __noinline__ void intra_exec(const Grp g, const Intra2Frame &a, const Dav1dCudaIntraDesc &d,
                                        pixel *edge, pixel *scratch, const int z2_centre, int16_t *tile,
                                        const int part, const int n_parts) {
    const int pl = d.plane;
    const int ss_hor = pl ? a.pic.ss_hor : 0, ss_ver = pl ? a.pic.ss_ver : 0;
    const PlaneView &pv = a.pic.p[pl];
    const int stride = (int)(pv.stride / (int)sizeof(pixel));
    pixel *dst = (pixel *)pv.data + (int64_t)d.y4 * 4 * stride + d.x4 * 4;
    const int w = d.tw4 * 4, h = d.th4 * 4;
    const int lw = 31 - __clz(w);
    const int bdmax = a.pic.bdmax;
    const int have_left = d.x4 > d.tile_x4_start, have_top = d.y4 > d.tile_y4_start;
    const int mode = d.mode;
    const bool has_res = d.eob >= 0 && mode != DAV1D_CUDA_INTRA_PAL;
    const int16_t *res = nullptr;
    int rstride = 0;
    if (has_res) {
        const PlaneView &rv = a.res.p[pl];
        rstride = (int)(rv.stride / 2);
        res = (const int16_t *)rv.data + (int64_t)d.y4 * 4 * rstride + d.x4 * 4;
    }
    // first residual / current pixel of this lane: loaded now, used after the edge preparation
    // Four pixels of a row per lane and step.  Residuals / current pixels of a step are loaded one
    // step ahead (the first ones right here, so that they overlap the edge preparation).
    const int n = w * h;
    const int i0 = (part << 8) + 4 * g.gl;
    const bool rd_dst = mode == DAV1D_CUDA_INTRA_NONE || mode == DAV1D_CUDA_INTRA_II;
    uint2 r_first = make_uint2(0u, 0u);
    int c_first[4] = { 0, 0, 0, 0 };
    if (i0 < n) {
        if (res) r_first = *(const uint2 *)(res + (i0 >> lw) * rstride + (i0 & (w - 1)));
        if (rd_dst) load_px<pixel, 4>(dst + (i0 >> lw) * stride + (i0 & (w - 1)), c_first);
    }
    // what the pixel loop does: 0 predictor, 1 palette, 2 intrabc, 3 keep the current pixel
    int kind = 3;
    PixParams<pixel> P;
    P.pm = PM_CONST; P.p0 = P.p1 = P.p2 = P.p3 = 0; P.edge = P.e0 = P.e1 = edge; P.tile = tile; P.w = w; P.h = h;
    const uint8_t *bmask = nullptr;                // inter-intra blend mask
    if (D1_SKIP(8)) return;
    if (D1_SKIP(1)) {
        kind = 0;
    } else if (mode == DAV1D_CUDA_INTRA_PAL) {
        kind = 1;
    } else if (mode == DAV1D_CUDA_INTRA_IBC) {
        kind = 2;
    } else if (mode == DAV1D_CUDA_INTRA_CFL) {
        const PlaneView &lv = a.pic.p[0];
        const int lstride = (int)(lv.stride / (int)sizeof(pixel));
        const pixel *luma = (const pixel *)lv.data + (int64_t)((d.y4 * 4) << ss_ver) * lstride + ((d.x4 * 4) << ss_hor);
        cfl_ac_block<pixel>(g, tile, luma, lstride, d.aux & 0xff, (d.aux >> 8) & 0xff, w, h, ss_hor, ss_ver);
        int angle = 0;
        const int m = prepare_edges<pixel>(g, d.x4, have_left, d.y4, have_top, d.tile_x4_end, d.tile_y4_end, 0, dst,
                                           stride, nullptr, 0, &angle, d.tw4, d.th4, 0, edge, bdmax);
        P = cfl_setup<pixel>(g, m, edge, w, h, tile, d.angle_delta, bdmax);
        kind = 0;
    } else if (mode != DAV1D_CUDA_INTRA_NONE) {
        // inter-intra (recon_tmpl.c:1658-1681): the whole-block intra prediction (edge_flags 0, no
        // edge filter) is blended onto the inter prediction that is in dst
        const bool ii = mode == DAV1D_CUDA_INTRA_II;
        int angle = ii ? 0 : d.angle_delta;
        const int m = prepare_edges<pixel>(g, d.x4, have_left, d.y4, have_top, d.tile_x4_end, d.tile_y4_end,
                                           ii ? 0 : d.edge_flags, dst, stride, nullptr, ii ? d.angle_delta : mode,
                                           &angle, d.tw4, d.th4, ii ? 0 : (d.flags >> 10) & 1, edge, bdmax);
        if (ii) { bmask = a.pal_idx + d.coef_off; angle = 0; }
        else angle |= d.flags;
        const int max_w = ((4 * a.bw4 + ss_hor) >> ss_hor) - 4 * d.x4;
        const int max_h = ((4 * a.bh4 + ss_ver) >> ss_ver) - 4 * d.y4;
        P = ipred_setup<pixel>(g, m, angle, w, h, max_w, max_h, edge, scratch, (pixel *)tile, bdmax, z2_centre);
        kind = 0;
    }
    grp_sync(g);

    // ---- the pixel loop, one instance per predictor (the warps of a round are sorted by predictor
    // class, so neighbouring warps run the same instance)
    const int step = n_parts > 1 ? (n_parts << 8) - 256 + 4 * g.G : 4 * g.G;   // the next chunk of this part after 256 pixels
    auto run = [&](auto pix) {
        int i = i0;
        uint2 r_nx = r_first;
        int c_nx[4] = { c_first[0], c_first[1], c_first[2], c_first[3] };
#pragma unroll 1
        while (i < n) {
            const int y = i >> lw, x = i & (w - 1);
            const uint2 r_cur = r_nx;
            int c_cur[4] = { c_nx[0], c_nx[1], c_nx[2], c_nx[3] };
            const int inx = (n_parts > 1 && ((i + 4 * g.G) & 255) < 4 * g.G) ? i + step : i + 4 * g.G;
            if (inx < n) {
                if (res) r_nx = *(const uint2 *)(res + (inx >> lw) * rstride + (inx & (w - 1)));
                if (rd_dst) load_px<pixel, 4>(dst + (inx >> lw) * stride + (inx & (w - 1)), c_nx);
            }
            int v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                v[k] = pix(x + k, y, i + k, c_cur[k]);
                if (bmask) v[k] = (c_cur[k] * (64 - bmask[i + k]) + v[k] * bmask[i + k] + 32) >> 6;    // mc.blend (mc_tmpl.c:642-653)
            }
            if (res) {
                v[0] = clip_px<pixel>(v[0] + (int)(int16_t)(r_cur.x & 0xffff), bdmax);
                v[1] = clip_px<pixel>(v[1] + ((int)r_cur.x >> 16), bdmax);
                v[2] = clip_px<pixel>(v[2] + (int)(int16_t)(r_cur.y & 0xffff), bdmax);
                v[3] = clip_px<pixel>(v[3] + ((int)r_cur.y >> 16), bdmax);
            }
            store_px<pixel, 4>(dst + y * stride + x, v);
            i = inx;
        }
    };
    if (D1_SKIP(2)) {
    } else if (kind == 0) {
        const pixel *e = P.edge;
        switch (P.pm) {
        case PM_CONST: run([&](int, int, int, int) { return P.p0; }); break;
        case PM_V: run([&](int x, int, int, int) { return (int)e[1 + x]; }); break;
        case PM_H: run([&](int, int y, int, int) { return (int)e[-(1 + y)]; }); break;
        default: run([&](int x, int y, int i, int) { return ipred_pixel<pixel>(P, x, y, i, bdmax); }); break;
        }
    } else if (kind == 1) {
        // pal_pred (ipred_tmpl.c:717-730): two pixels per index byte
        const uint8_t *idx = a.pal_idx + d.coef_off;
        const pixel *pal = (const pixel *)a.pal + d.aux;
        run([&](int x, int, int i, int) { const int q = idx[i >> 1]; return (int)pal[(x & 1) ? q >> 4 : q & 7]; });
    } else if (kind == 2) {
        // intrabc: put_bilin (mc_tmpl.c:395-450) from the current picture; coordinates clamped to
        // the 4*bw4 x 4*bh4 area (= emu_edge, recon_tmpl.c:974-995)
        const pixel *base = (const pixel *)pv.data;
        const int ib = PxTraits<pixel>::inter_bits(bdmax);
        const int sx = (int16_t)(d.aux & 0xffff), sy = (int16_t)(d.aux >> 16);
        const int mx = d.angle_delta, my = d.flags;
        const int pw = (4 * a.bw4) >> ss_hor, ph = (4 * a.bh4) >> ss_ver;
        run([&](int x, int y, int, int) {
            const int xa = iclip(sx + x, 0, pw - 1), xb = iclip(sx + x + 1, 0, pw - 1);
            const int ya = iclip(sy + y, 0, ph - 1), yb = iclip(sy + y + 1, 0, ph - 1);
            const int q00 = __ldcg(base + (int64_t)ya * stride + xa), q01 = __ldcg(base + (int64_t)ya * stride + xb);
            const int q10 = __ldcg(base + (int64_t)yb * stride + xa), q11 = __ldcg(base + (int64_t)yb * stride + xb);
            if (mx && my) {
                const int sh1 = 4 - ib, r1 = (1 << sh1) >> 1;
                const int m0 = (16 * q00 + mx * (q01 - q00) + r1) >> sh1;
                const int m1 = (16 * q10 + mx * (q11 - q10) + r1) >> sh1;
                const int sh2 = 4 + ib;
                return clip_px<pixel>((16 * m0 + my * (m1 - m0) + ((1 << sh2) >> 1)) >> sh2, bdmax);
            } else if (mx) {
                const int sh1 = 4 - ib;
                const int px = (16 * q00 + mx * (q01 - q00) + ((1 << sh1) >> 1)) >> sh1;
                return clip_px<pixel>((px + ((1 << ib) >> 1)) >> ib, bdmax);
            } else if (my) {
                return clip_px<pixel>((16 * q00 + my * (q10 - q00) + 8) >> 4, bdmax);
            }
            return q00;
        });
    } else {
        run([&](int, int, int, int c) { return c; });      // residual on top of what an earlier round left there
    }
    // the operation's pixels are stored: one count less on each of its cells (the grid barrier at
    // the end of the round makes pixels and counts visible together)
    if (part == 0 && !D1_SKIP(4)) {
        const int W = map_w(a, pl), H = map_h(a, pl);
        uint8_t *m = a.map + map_off(a, pl);
        const int ltw = 31 - __clz((int)d.tw4);
        const int nc = d.th4 << ltw;
        for (int j = g.gl; j < nc; j += g.G) {
            const int cx = d.x4 + (j & (d.tw4 - 1)), cy = d.y4 + (j >> ltw);
            if (cx < W && cy < H) {
                // minus one on the cell's byte: a reduction on the containing word (fire and forget)
                const size_t off = (size_t)(m - a.map) + (size_t)cy * W + cx;
                atomicAdd((unsigned *)(a.map + (off & ~(size_t)3)), 0u - (1u << (8 * (off & 3))));
            }
        }
    }
    grp_sync(g);
}

// per-warp shared memory of step 3: edge + Z-mode scratch of four octets (operations of up to 64
// pixels: w, h <= 16, w + h <= 20, i.e. 2 * 16 + 1 + 2 * 16 edge pixels: 80 + 80 pixels per octet) or
// of one warp-wide operation (EDGE_BUF + IPRED_SCRATCH pixels) in the same bytes, and the CfL ac /
// filter-intra tile (1024 values for a warp-wide operation, 256 per octet)
constexpr int OCT_PX = 80, OCT_CENTRE = 36, OCT_Z2 = 40;
template <typename pixel> struct __align__(16) ExecSmem {
    pixel es[4 * 2 * OCT_PX > EDGE_BUF + IPRED_SCRATCH ? 4 * 2 * OCT_PX : EDGE_BUF + IPRED_SCRATCH];
    int16_t tile[32 * 32];
};
// 256-pixel parts of an operation's pixel loop (tile-based predictors - CfL, filter-intra - stay whole)
DEV int op_parts(const Dav1dCudaIntraDesc &d) {
    const int px = d.tw4 * d.th4 * 16;
    if (px <= 256 || d.mode == DAV1D_CUDA_INTRA_CFL || d.mode == DAV1D_CUDA_INTRA_FILTER) return 1;
    return px >> 8;
}

constexpr int R_WARPS = 8;
// operation id: part (4 bits) | frame (6 bits) | index inside the frame (20 bits)
constexpr int OP_FRAME_SHIFT = 20, OP_PART_SHIFT = 28;
DEV int op_frame(const unsigned id) { return (int)((id >> OP_FRAME_SHIFT) & 63u); }
DEV int op_index(const unsigned id) { return (int)(id & ((1u << OP_FRAME_SHIFT) - 1u)); }
constexpr unsigned R_NOCELL = 0x7fffffffu;
constexpr int R_BINS = 128;             // sort key: size class * 32 + predictor class
// size classes: 0 = up to 16 pixels, 1 = up to 64 (both: one operation per octet, four per warp),
// 2 = up to 256, 3 = larger (one operation or 256-pixel part per warp)
DEV int size_class(const Dav1dCudaIntraDesc &d) {
    const int c4 = d.tw4 * d.th4;
    return c4 <= 1 ? 0 : c4 <= 4 ? 1 : c4 <= 16 ? 2 : 3;
}
// counters of a round (two sets, used alternately)
struct RoundCtr {
    unsigned n_next, n_ready;          // operations that stay pending / are ready
    unsigned hist[R_BINS];
    unsigned cursor[R_BINS];
};
struct RoundsArgs {
    Intra2Args g;
    int op_base[I2_MAXF + 1];           // first global operation number of every frame
    unsigned *pend[2];                  // pending operation ids
    unsigned *pend_blk[2];              // ... and one cell each of them waits for (R_NOCELL: unknown)
    unsigned *ready;                    // ready operations of the round, unsorted
    uint8_t *ready_key;
    unsigned *sorted;                   // ... sorted by key
    unsigned *bar;                      // grid barrier counter (zeroed before the launch)
    RoundCtr *ctr;                      // [2] (zeroed before the launch)
#ifdef D1_EXPERIMENT
    unsigned long long *trace;          // per round: 4 time stamps (ns), pending, ready entries
    int skip;                           // timing experiments (wrong results): 1 edges + set-up, 2 pixel loop, 4 count-down, 8 whole exec
#endif
};
#ifdef D1_EXPERIMENT
DEV unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define D1_TRACE(slot, val) do { if (gtid == 0 && a.trace && round < 256) a.trace[round * 6 + (slot)] = (val); } while (0)
#else
#define D1_TRACE(slot, val) do { } while (0)
#endif

template <typename pixel>
__global__ void __launch_bounds__(R_WARPS * 32, 4) intra_rounds_kernel(const __grid_constant__ RoundsArgs a) {
    extern __shared__ __align__(16) uint8_t rounds_smem_raw[];
    __shared__ unsigned s_hist[R_BINS], s_base[R_BINS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned gtid = blockIdx.x * blockDim.x + tid, gthreads = gridDim.x * blockDim.x;
    ExecSmem<pixel> *sm = (ExecSmem<pixel> *)rounds_smem_raw + warp;
    unsigned target = 0;
    unsigned n_pend = (unsigned)a.op_base[a.g.nf];
#ifdef D1_EXPERIMENT
    if (gtid == 0) g_skip = a.skip;
#endif
    for (int round = 0; n_pend > 0; round++) {
        RoundCtr *ctr = a.ctr + (round & 1), *nxt = a.ctr + ((round & 1) ^ 1);
        const unsigned *pend = a.pend[round & 1], *pblk = a.pend_blk[round & 1];
        unsigned *pend_next = a.pend[(round & 1) ^ 1], *pblk_next = a.pend_blk[(round & 1) ^ 1];
        // ---- step 1: every pending operation is looked at by one thread
        D1_TRACE(0, gtimer()); D1_TRACE(4, n_pend);
        if (tid < R_BINS) s_hist[tid] = 0;
        __syncthreads();
        for (unsigned k0 = blockIdx.x * blockDim.x; k0 < n_pend; k0 += gthreads) {
            const unsigned k = k0 + tid;
            bool ready = false, live = k < n_pend;
            unsigned id = 0, blk = R_NOCELL;
            int key = 0, parts = 0;
            if (live) {
                if (round == 0) {
                    int fi = 0;
                    while (fi + 1 < a.g.nf && (unsigned)a.op_base[fi + 1] <= k) fi++;
                    id = ((unsigned)fi << OP_FRAME_SHIFT) | (k - (unsigned)a.op_base[fi]);
                } else {
                    id = __ldcg(pend + k);
                    blk = __ldcg(pblk + k);
                }
                const Intra2Frame &f = a.g.f[op_frame(id)];
                // the cell this operation was seen waiting for: still above its threshold?
                const bool still = blk != R_NOCELL && __ldcg(f.map + (blk & 0x7fffffffu)) > (blk >> 31);
                if (!still) {
                    const Dav1dCudaIntraDesc d = f.descs[op_index(id)];
                    int cls = 0;
                    ready = op_ready(f, d, &cls, &blk);
                    key = size_class(d) * 32 + cls;
                    if (ready) parts = op_parts(d);
                }
            }
            // warp-aggregated appends: a ready operation adds one entry per part to the round's list
            const unsigned mp = __ballot_sync(0xffffffffu, live && !ready);
            int incl = parts;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            unsigned br = 0, bp = 0;
            if (lane == 0) {
                if (tot) br = atomicAdd(&ctr->n_ready, (unsigned)tot);
                if (mp) bp = atomicAdd(&ctr->n_next, __popc(mp));
            }
            br = __shfl_sync(0xffffffffu, br, 0);
            bp = __shfl_sync(0xffffffffu, bp, 0);
            if (ready) {
                const unsigned pos = br + (unsigned)(incl - parts);
                for (int q = 0; q < parts; q++) {
                    a.ready[pos + q] = id | ((unsigned)q << OP_PART_SHIFT);
                    a.ready_key[pos + q] = (uint8_t)key;
                }
                atomicAdd(&s_hist[key], (unsigned)parts);
            } else if (live) {
                const unsigned pos = bp + __popc(mp & ((1u << lane) - 1u));
                pend_next[pos] = id;
                pblk_next[pos] = blk;
            }
        }
        __syncthreads();
        if (tid < R_BINS && s_hist[tid]) atomicAdd(&ctr->hist[tid], s_hist[tid]);
        grid_barrier(a.bar, target);
        const unsigned n_ready = ld_acquire_u32(&ctr->n_ready), n_next = ld_acquire_u32(&ctr->n_next);
        D1_TRACE(1, gtimer()); D1_TRACE(5, n_ready);
        if (n_ready == 0) {
            // pending operations, none ready: they wait for each other - inconsistent descriptors
            if (gtid == 0) atomicOr(a.g.status, 1u);
            return;
        }
        // ---- step 2: counting sort of the ready list by key.  Start of every bin (all blocks
        // compute the same prefix), then every block scatters a contiguous chunk of the list.
        if (tid < R_BINS) {                    // warps 0..3: exclusive prefix over the bins
            const unsigned h = ld_acquire_u32(&ctr->hist[tid]);
            unsigned incl = h;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            s_base[tid] = incl - h;
            if (lane == 31) s_hist[warp] = incl;      // total of this warp's 32 bins
        }
        __syncthreads();
        if (tid < R_BINS) {
            unsigned add = 0;
            for (int q = 0; q < warp; q++) add += s_hist[q];
            s_base[tid] += add;
        }
        __syncthreads();
        const unsigned n_tot_check = s_base[R_BINS - 1];
        (void)n_tot_check;
        if (tid < R_BINS) s_hist[tid] = 0;
        __syncthreads();
        // list layout: [octet-mode entries | warp-mode entries]
        const unsigned n_oct = s_base[64];
        {
            const unsigned chunk = (n_ready + gridDim.x - 1) / gridDim.x;
            const unsigned c0 = blockIdx.x * chunk, c1 = min(n_ready, c0 + chunk);
            for (unsigned k = c0 + tid; k < c1; k += blockDim.x) atomicAdd(&s_hist[__ldcg(a.ready_key + k)], 1u);
            __syncthreads();
            if (tid < R_BINS) {
                const unsigned c = s_hist[tid];
                s_hist[tid] = s_base[tid] + (c ? atomicAdd(&ctr->cursor[tid], c) : 0u);   // this block's range of the bin
            }
            __syncthreads();
            for (unsigned k = c0 + tid; k < c1; k += blockDim.x) {
                const unsigned pos = atomicAdd(&s_hist[__ldcg(a.ready_key + k)], 1u);
                a.sorted[pos] = __ldcg(a.ready + k);
            }
        }
        // the other counter set is free (last read before the previous round's final barrier)
        if (gtid < sizeof(RoundCtr) / 4) ((unsigned *)nxt)[gtid] = 0;
        grid_barrier(a.bar, target);
        D1_TRACE(2, gtimer());
        // ---- step 3: execute.  Items: four small operations per warp (one per octet), then the
        // others one per warp
        {
            const unsigned items_oct = (n_oct + 3) / 4, items = items_oct + (n_ready - n_oct);
            const unsigned gw = blockIdx.x * R_WARPS + warp, nw = gridDim.x * R_WARPS;
            const int o = lane >> 3;
            for (unsigned it = gw; it < items; it += nw) {
                if (it < items_oct) {
                    const unsigned k = it * 4 + o;
                    if (k < n_oct) {
                        const unsigned id = __ldcg(a.sorted + k);
                        const Intra2Frame &f = a.g.f[op_frame(id)];
                        const Dav1dCudaIntraDesc d = f.descs[op_index(id)];
                        pixel *es = sm->es + o * 2 * OCT_PX;
                        intra_exec<pixel>(grp_octet(lane), f, d, es + OCT_CENTRE, es + OCT_PX, OCT_Z2, sm->tile + 256 * o, 0, 1);
                    }
                } else {
                    const unsigned id = __ldcg(a.sorted + n_oct + (it - items_oct));
                    const Intra2Frame &f = a.g.f[op_frame(id)];
                    const Dav1dCudaIntraDesc d = f.descs[op_index(id)];
                    intra_exec<pixel>(grp_warp(lane), f, d, sm->es + EDGE_C, sm->es + EDGE_BUF, 128 + 8, sm->tile,
                                      (int)(id >> OP_PART_SHIFT), op_parts(d));
                }
                __syncwarp();
            }
        }
        grid_barrier(a.bar, target);
        D1_TRACE(3, gtimer());
        n_pend = n_next;
        if (round > (1 << 20)) { if (gtid == 0) atomicOr(a.g.status, 1u); return; }
    }
}

// cell map set-up: every operation adds one to each of its cells (four cells per word)
__global__ void intra2_mark_kernel(const __grid_constant__ Intra2Args a) {
    const int fi = blockIdx.y;
    const Intra2Frame &f = a.f[fi];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < f.n_ops; i += gridDim.x * blockDim.x) {
        const Dav1dCudaIntraDesc &d = f.descs[i];
        const int pl = d.plane, W = map_w(f, pl), H = map_h(f, pl);
        uint8_t *m = f.map + map_off(f, pl);
        for (int y = d.y4; y < imin(d.y4 + d.th4, H); y++)
            for (int x = d.x4; x < imin(d.x4 + d.tw4, W); x++) {
                const size_t off = (size_t)(m - f.map) + (size_t)y * W + x;
                atomicAdd((unsigned *)(f.map + (off & ~(size_t)3)), 1u << (8 * (off & 3)));
            }
    }
}

static int g_rounds_blocks[2] = { 0, 0 };   // co-resident blocks of the executor per pixel type
template <typename pixel> static size_t rounds_smem_bytes() { return R_WARPS * sizeof(ExecSmem<pixel>); }

void recon_init_attrs() {
    itx_init_attrs();
    cudaFuncSetAttribute(intra_rounds_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)rounds_smem_bytes<uint8_t>());
    cudaFuncSetAttribute(intra_rounds_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)rounds_smem_bytes<uint16_t>());
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, intra_rounds_kernel<uint8_t>, R_WARPS * 32, rounds_smem_bytes<uint8_t>());
    g_rounds_blocks[0] = std::max(1, occ) * std::max(1, sms);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, intra_rounds_kernel<uint16_t>, R_WARPS * 32, rounds_smem_bytes<uint16_t>());
    g_rounds_blocks[1] = std::max(1, occ) * std::max(1, sms);
}

// ---- warp batch: one warp per 8x8
struct WarpBatchArgs {
    PicView dst;
    PicView refs[7];
    const Dav1dCudaWarpDesc *descs;
    int n;
};
template <typename pixel>
__global__ void __launch_bounds__(128) warp_batch_kernel(const __grid_constant__ WarpBatchArgs a) {
    __shared__ int16_t mid[4][15 * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 4 + warp;
    if (i >= a.n) return;
    const Dav1dCudaWarpDesc d = a.descs[i];
    const PlaneView &dp = a.dst.p[d.plane];
    const int dstride = (int)(dp.stride / (int)sizeof(pixel));
    pixel *out = (pixel *)dp.data + (int64_t)d.y * dstride + d.x;
    mc_warp8x8<pixel, false>(a.refs[d.ref].p[d.plane], d.sx, d.sy, d.abcd, d.mx, d.my, a.dst.bdmax, mid[warp],
                             out, dstride, lane);
}

static int warp_batch_launch(const PicView &dst, const PicView *refs, const Dav1dCudaWarpDesc *descs, int n,
                             cudaStream_t st)
{
    if (n <= 0) return 0;
    WarpBatchArgs a;
    a.dst = dst;
    for (int i = 0; i < 7; i++) a.refs[i] = refs[i];
    a.descs = descs;
    a.n = n;
    if (dst.bdmax > 0xff) warp_batch_kernel<uint16_t><<<(n + 3) / 4, 128, 0, st>>>(a);
    else warp_batch_kernel<uint8_t><<<(n + 3) / 4, 128, 0, st>>>(a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "warp_batch_kernel") ? 0 : -5;
}

static void refs_view(PicView *out, const Dav1dCudaPicture *const refs[7]) {
    memset(out, 0, 7 * sizeof(PicView));
    for (int i = 0; i < 7; i++)
        if (refs[i]) out[i] = pic_view(refs[i]);
}

static bool ensure_aux(Dav1dCudaContext *c) {
    if (c->aux_ready) return true;
    for (int i = 0; i < Dav1dCudaContext::N_AUX; i++) {
        if (!cuda_ok(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking), "aux stream")) return false;
        if (!cuda_ok(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming), "aux event")) return false;
    }
    if (!cuda_ok(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming), "fork event")) return false;
    c->aux_ready = true;
    return true;
}
// fork: aux streams wait for everything submitted to `st` so far
static bool fork_aux(Dav1dCudaContext *c, cudaStream_t st) {
    if (!cuda_ok(cudaEventRecord(c->ev_fork, st), "fork record")) return false;
    for (int i = 0; i < Dav1dCudaContext::N_AUX; i++)
        if (!cuda_ok(cudaStreamWaitEvent(c->aux[i], c->ev_fork, 0), "fork wait")) return false;
    return true;
}
// join: `st` waits for the aux streams
static bool join_aux(Dav1dCudaContext *c, cudaStream_t st) {
    for (int i = 0; i < Dav1dCudaContext::N_AUX; i++) {
        if (!cuda_ok(cudaEventRecord(c->ev_join[i], c->aux[i]), "join record")) return false;
        if (!cuda_ok(cudaStreamWaitEvent(st, c->ev_join[i], 0), "join wait")) return false;
    }
    return true;
}

// workspace of the executor's rounds: barrier + counters, then pending x 2, ready, sorted (ids) and
// the ready keys.  One per context, grown on demand outside any stream capture; submissions of a
// context are ordered on its stream.
static size_t rounds_ws_hdr() { return 256 + ((2 * sizeof(RoundCtr) + 255) & ~(size_t)255); }
#ifdef D1_EXPERIMENT
static unsigned long long *d1_last_trace = nullptr;
#endif
static size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }
// total = operations of the group (pending lists); cap = entries a round's list can hold: an
// operation per entry, large ones one entry per 256 pixels
static size_t rounds_ws_need(size_t total, size_t cap) {
    return rounds_ws_hdr() + 4 * al256(total * 4) + 2 * al256(cap * 4) + al256(cap);
}
static void rounds_counts(const Dav1dCudaReconBatch *const *bs, int n, size_t *total, size_t *cap) {
    *total = 0; *cap = 0;
    for (int f = 0; f < n; f++) {
        if (!bs[f] || bs[f]->n_intra <= 0) continue;
        *total += (size_t)bs[f]->n_intra;
        *cap += (size_t)bs[f]->n_intra + 3 * ((size_t)bs[f]->bw4 * bs[f]->bh4 * 16 / 256 + 1);
    }
}
static int ensure_rounds_ws(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n) {
    size_t total, cap;
    rounds_counts(bs, n, &total, &cap);
    const size_t need = rounds_ws_need(total, cap);
    if (need <= c->rounds_ws_bytes) return 0;
    if (c->rounds_ws) {
        D1_CHECK(cudaStreamSynchronize(c->stream));
        cudaFree(c->rounds_ws);
        c->rounds_ws = nullptr; c->rounds_ws_bytes = 0;
    }
    const size_t want = need + need / 4;
    D1_CHECK(cudaMalloc(&c->rounds_ws, want));
    c->rounds_ws_bytes = want;
    return 0;
}

static int check_group(const Dav1dCudaReconBatch *const *bs, int n) {
    if (!bs || n < 1 || n > I2_MAXF) return -22;
    for (int f = 0; f < n; f++) {
        const Dav1dCudaReconBatch *b = bs[f];
        if (!b || !b->dst || !b->dst->p[0].data) return -22;
        // one pixel type per group; a frame may not predict from another member's output
        if ((b->dst->bitdepth_max > 0xff) != (bs[0]->dst->bitdepth_max > 0xff)) return -22;
        for (int g = 0; g < n; g++)
            for (int r = 0; r < 7; r++)
                if (g != f && bs[g] && bs[g]->refs[r] && bs[g]->refs[r]->p[0].data == b->dst->p[0].data) return -22;
        if (b->n_intra > 0 && (!b->intra || !b->intra_cellmap)) return -22;
        if (b->n_intra >= (1 << 20)) return -22;
        if (b->n_intra > 0 && b->intra_itx && !b->intra_res) return -22;
    }
    return 0;
}

// The frames of a group (independent streams, one pixel type).  Phases A (prediction from
// reference frames) and B (inter residuals) are launched per frame, spread over the fork / join
// streams; phase C is the marking launch + ONE executor launch for the whole group.
// phase_mask: bit0 put (+OBMC), bit1 compound, bit2 warp, bit3 inter residual, bit4 intra.
static int group_submit_on(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n, cudaStream_t st,
                           const int mask)
{
    int r;
    if ((r = check_group(bs, n))) return r;
    if (!ensure_aux(c)) return -5;
    cudaStream_t ss[1 + Dav1dCudaContext::N_AUX] = { st, c->aux[0], c->aux[1], c->aux[2] };
    constexpr int NS = 1 + Dav1dCudaContext::N_AUX;
    const bool hbd = bs[0]->dst->bitdepth_max > 0xff;
    // intra: cell map set-up first, on its own branch (it touches nothing the other phases use)
    Intra2Args ia;
    memset(&ia, 0, sizeof(ia));
    int n_ops = 0;
    if (mask & 16) {
        ia.nf = n;
        for (int f = 0; f < n; f++) {
            const Dav1dCudaReconBatch *b = bs[f];
            Intra2Frame &p = ia.f[f];
            p.pic = pic_view(b->dst); p.bw4 = b->bw4; p.bh4 = b->bh4; p.cf = b->cf;
            if (b->intra_res) p.res = pic_view(b->intra_res);
            p.descs = b->intra; p.pal = b->pal; p.pal_idx = b->pal_idx;
            p.n_ops = b->n_intra > 0 ? b->n_intra : 0;
            p.map = b->intra_cellmap;
            n_ops = std::max(n_ops, p.n_ops);
        }
        ia.status = c->status;
    }
    if (!fork_aux(c, st)) return -5;
    if ((mask & 16) && n_ops > 0) {
        intra2_mark_kernel<<<dim3((unsigned)std::min((n_ops + 255) / 256, 64), (unsigned)n), 256, 0, ss[NS - 1]>>>(ia);
        count_launch();
    }
    for (int f = 0; f < n; f++) {
        const Dav1dCudaReconBatch *b = bs[f];
        cudaStream_t s = ss[f % NS];
        const PicView dst = pic_view(b->dst);
        PicView refs[7];
        refs_view(refs, b->refs);
        if ((mask & 1) && (r = mc_put_launch_raw(dst, refs, b->mc_put, b->mc_put_tiles, b->n_mc_put_tiles, b->n_mc_put_small, nullptr,
                                                 nullptr, false, s))) return r;
        if ((mask & 2) && (r = mc_put_launch_raw(dst, refs, b->mc_comp, b->mc_comp_tiles, b->n_mc_comp_tiles[0],
                                                 b->n_mc_comp_small[0], b->masks, nullptr, true, s))) return r;
        if ((mask & 2) && (r = mc_put_launch_raw(dst, refs, b->mc_comp, b->mc_comp_tiles + b->n_mc_comp_tiles[0],
                                                 b->n_mc_comp_tiles[1], b->n_mc_comp_small[1], b->masks, nullptr, true, s))) return r;
        if ((mask & 4) && (r = warp_batch_launch(dst, refs, b->warp, b->n_warp, s))) return r;
        if ((mask & 1) && b->mc_obmc) {
            // OBMC blends onto the finished predictions: top neighbours, then left neighbours
            if ((r = mc_obmc_launch_raw(dst, refs, b->mc_obmc, b->mc_obmc_tiles, b->n_mc_obmc_tiles[0], s))) return r;
            if ((r = mc_obmc_launch_raw(dst, refs, b->mc_obmc, b->mc_obmc_tiles + b->n_mc_obmc_tiles[0],
                                        b->n_mc_obmc_tiles[1], s))) return r;
        }
        if ((mask & 8) && b->itx && b->itx_tasks) {
            if ((r = itx_task_launch(dst, nullptr, b->cf, b->itx, b->itx_tasks, b->n_itx_tasks[0], b->n_itx_tasks[1], 0, s, s)))
                return r;
        } else if ((mask & 8) && b->itx && (r = itx_batch_launch(dst, nullptr, b->cf, b->itx, b->itx_class_count, 0, s))) return r;
        // intra residual pre-pass -> int16 residual planes (independent of everything above)
        if ((mask & 16) && b->n_intra > 0 && b->intra_itx && b->intra_res) {
            const PicView rv = pic_view(b->intra_res);
            cudaStream_t s2 = ss[(f + 2) % NS];
            if (b->intra_itx_tasks) {
                if ((r = itx_task_launch(dst, &rv, b->cf, b->intra_itx, b->intra_itx_tasks, b->n_intra_itx_tasks[0],
                                         b->n_intra_itx_tasks[1], 0, s2, s2))) return r;
            } else if ((r = itx_batch_launch(dst, &rv, b->cf, b->intra_itx, b->intra_itx_class_count, 0, s2))) return r;
        }
    }
    if (!join_aux(c, st)) return -5;
    if (!(mask & 16) || n_ops <= 0) return 0;
    // the executor: one cooperative launch (its blocks wait for each other at the grid barriers)
    RoundsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.g = ia;
    size_t total = 0;
    for (int f = 0; f < n; f++) { ra.op_base[f] = (int)total; total += (size_t)ia.f[f].n_ops; }
    ra.op_base[n] = (int)total;
    if (total >= ((size_t)1 << 31)) return -22;
    const size_t hdr = rounds_ws_hdr();
    size_t total2, cap;
    rounds_counts(bs, n, &total2, &cap);
    if (rounds_ws_need(total, cap) > c->rounds_ws_bytes) return -12;      // ensure_rounds_ws() comes first
    uint8_t *ws = (uint8_t *)c->rounds_ws;
    ra.bar = (unsigned *)ws;
    ra.ctr = (RoundCtr *)(ws + 256);
    const size_t lb = al256(total * 4), cb = al256(cap * 4);
    ra.pend[0] = (unsigned *)(ws + hdr);
    ra.pend[1] = (unsigned *)(ws + hdr + lb);
    ra.pend_blk[0] = (unsigned *)(ws + hdr + 2 * lb);
    ra.pend_blk[1] = (unsigned *)(ws + hdr + 3 * lb);
    ra.ready = (unsigned *)(ws + hdr + 4 * lb);
    ra.sorted = (unsigned *)(ws + hdr + 4 * lb + cb);
    ra.ready_key = ws + hdr + 4 * lb + 2 * cb;
    D1_CHECK(cudaMemsetAsync(ws, 0, hdr, st));
#ifdef D1_EXPERIMENT
    static unsigned long long *g_trace = nullptr;
    if (!g_trace) cudaMalloc(&g_trace, 256 * 6 * 8);
    cudaMemsetAsync(g_trace, 0, 256 * 6 * 8, st);
    ra.trace = g_trace;
    ra.skip = getenv("D1_SKIP") ? atoi(getenv("D1_SKIP")) : 0;
    d1_last_trace = g_trace;
#endif
    const size_t smem = hbd ? rounds_smem_bytes<uint16_t>() : rounds_smem_bytes<uint8_t>();
    int resident = g_rounds_blocks[hbd];
#ifdef D1_EXPERIMENT
    if (getenv("D1_ROUNDS_BPSM")) resident = std::min(resident, atoi(getenv("D1_ROUNDS_BPSM")) * c->num_sms);
#endif
    const int grid = (int)std::min<size_t>((size_t)resident, (total + R_WARPS * 32 - 1) / (R_WARPS * 32));
    void *kargs[] = { (void *)&ra };
    const void *fn = hbd ? (const void *)intra_rounds_kernel<uint16_t> : (const void *)intra_rounds_kernel<uint8_t>;
    D1_CHECK(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(R_WARPS * 32), kargs, smem, st));
    count_launch();
    return 0;
}

}  // namespace d1

using namespace d1;

struct Dav1dCudaReconGraph {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    int n_nodes;
};

extern "C" {

int dav1d_cuda_warp_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, const Dav1dCudaPicture *const refs[7],
                          const Dav1dCudaWarpDesc *descs, int n)
{
    if (!c || !dst || !descs) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    PicView rv[7];
    refs_view(rv, refs);
    return warp_batch_launch(pic_view(dst), rv, descs, n, c->stream);
}

size_t dav1d_cuda_intra_cellmap_bytes(int bw4, int bh4, int ss_hor, int ss_ver) {
    const size_t cw = (size_t)((bw4 + ss_hor) >> ss_hor), ch = (size_t)((bh4 + ss_ver) >> ss_ver);
    return (((size_t)bw4 * bh4 + 2 * cw * ch) + 255) & ~(size_t)255;
}

#ifdef D1_EXPERIMENT
// experiment builds only (make EXTRA=-DD1_EXPERIMENT): per-round time stamps of the last executor launch
__attribute__((visibility("default"))) int dav1d_cuda_debug_rounds_trace(unsigned long long *host, int rounds) {
    if (!d1_last_trace) return -1;
    cudaDeviceSynchronize();
    cudaMemcpy(host, d1_last_trace, (size_t)rounds * 6 * 8, cudaMemcpyDeviceToHost);
    return 0;
}
#endif

int dav1d_cuda_recon_submit(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b) {
    if (!c || !b) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    if (ensure_rounds_ws(c, &b, 1)) return -12;
    return group_submit_on(c, &b, 1, c->stream, 31);
}

int dav1d_cuda_recon_submit_phases(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b, int phase_mask) {
    if (!c || !b) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    if (ensure_rounds_ws(c, &b, 1)) return -12;
    return group_submit_on(c, &b, 1, c->stream, phase_mask & 31);
}

int dav1d_cuda_recon_group_submit(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n) {
    if (!c) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    if (!bs || n < 1 || n > I2_MAXF) return -22;
    if (ensure_rounds_ws(c, bs, n)) return -12;
    return group_submit_on(c, bs, n, c->stream, 31);
}

int dav1d_cuda_recon_group_submit_phases(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n,
                                         int phase_mask)
{
    if (!c) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    if (!bs || n < 1 || n > I2_MAXF) return -22;
    if (ensure_rounds_ws(c, bs, n)) return -12;
    return group_submit_on(c, bs, n, c->stream, phase_mask & 31);
}

// The launches of a group captured into a CUDA graph (replayable while addresses and counts stay
// the same).  The capture stream is destroyed on every path.
int dav1d_cuda_recon_graph_build_multi_phases(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n,
                                              int phase_mask, Dav1dCudaReconGraph **out)
{
    if (!c || !bs || n < 1 || !out) return -22;
    *out = nullptr;
    D1_CHECK(cudaSetDevice(c->device));
    if (check_group(bs, n)) return -22;
    if (!ensure_aux(c)) return -5;
    if (ensure_rounds_ws(c, bs, n)) return -12;      // no allocation inside the capture
    cudaStream_t cap;
    D1_CHECK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    if (!cuda_ok(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture")) {
        cudaStreamDestroy(cap);
        return -5;
    }
    const int r = group_submit_on(c, bs, n, cap, phase_mask & 31);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(cap, &graph);
    cudaStreamDestroy(cap);
    if (r) { if (graph) cudaGraphDestroy(graph); return r; }
    if (!cuda_ok(e, "cudaStreamEndCapture")) return -5;
    Dav1dCudaReconGraph *g = new Dav1dCudaReconGraph();
    g->graph = graph;
    size_t nn = 0;
    cudaGraphGetNodes(graph, nullptr, &nn);
    g->n_nodes = (int)nn;
    if (!cuda_ok(cudaGraphInstantiate(&g->exec, graph, 0), "cudaGraphInstantiate")) {
        cudaGraphDestroy(graph);
        delete g;
        return -5;
    }
    *out = g;
    return g->n_nodes;
}

int dav1d_cuda_recon_graph_build_multi(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n,
                                       Dav1dCudaReconGraph **out)
{
    return dav1d_cuda_recon_graph_build_multi_phases(c, bs, n, 31, out);
}

int dav1d_cuda_recon_graph_build(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b, Dav1dCudaReconGraph **out) {
    return dav1d_cuda_recon_graph_build_multi_phases(c, &b, 1, 31, out);
}

int dav1d_cuda_recon_graph_launch(Dav1dCudaContext *c, Dav1dCudaReconGraph *g) {
    if (!c || !g) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    D1_CHECK(cudaGraphLaunch(g->exec, c->stream));
    count_launch(g->n_nodes);
    return 0;
}

void dav1d_cuda_recon_graph_free(Dav1dCudaReconGraph *g) {
    if (!g) return;
    cudaGraphExecDestroy(g->exec);
    cudaGraphDestroy(g->graph);
    delete g;
}

void *dav1d_cuda_malloc(size_t bytes) {
    void *p = nullptr;
    if (!cuda_ok(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc")) return nullptr;
    return p;
}
void dav1d_cuda_free(void *p) { if (p) cudaFree(p); }
int dav1d_cuda_upload(Dav1dCudaContext *c, void *dev, const void *host, size_t bytes) {
    D1_CHECK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}
int dav1d_cuda_download(Dav1dCudaContext *c, void *host, const void *dev, size_t bytes) {
    D1_CHECK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}
int dav1d_cuda_memset(Dav1dCudaContext *c, void *dev, int value, size_t bytes) {
    D1_CHECK(cudaMemsetAsync(dev, value, bytes, c->stream));
    return 0;
}
void *dav1d_cuda_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (!cuda_ok(cudaMallocHost(&p, bytes ? bytes : 1), "cudaMallocHost")) return nullptr;
    return p;
}
void dav1d_cuda_host_free(void *p) { if (p) cudaFreeHost(p); }
void *dav1d_cuda_event_create(void) {
    cudaEvent_t e;
    if (!cuda_ok(cudaEventCreate(&e), "cudaEventCreate")) return nullptr;
    return (void *)e;
}
int dav1d_cuda_event_record(Dav1dCudaContext *c, void *ev) {
    D1_CHECK(cudaEventRecord((cudaEvent_t)ev, c->stream));
    return 0;
}
int dav1d_cuda_stream_wait_event(Dav1dCudaContext *c, void *ev) {
    D1_CHECK(cudaStreamWaitEvent(c->stream, (cudaEvent_t)ev, 0));
    return 0;
}
float dav1d_cuda_event_elapsed_ms(void *start, void *stop) {
    float ms = -1.f;
    if (!cuda_ok(cudaEventSynchronize((cudaEvent_t)stop), "cudaEventSynchronize")) return -1.f;
    if (!cuda_ok(cudaEventElapsedTime(&ms, (cudaEvent_t)start, (cudaEvent_t)stop), "cudaEventElapsedTime")) return -1.f;
    return ms;
}
void dav1d_cuda_event_destroy(void *ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }

}  // extern "C"
