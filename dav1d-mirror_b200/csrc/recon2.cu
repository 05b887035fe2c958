// Frame-level batched reconstruction: the intra-class executor and the per-frame / per-group
// submit entry points.
//
// Reference call sites replaced: dav1d_recon_b_intra (src/recon_tmpl.c:1195-1596) and, through the
// MC / ITX launches, dav1d_recon_b_inter (:1598-2036).
//
// Intra executor.  Intra prediction of a transform block reads final pixels of its neighbours
// (recon_tmpl.c:1259-1347), so the intra-class operations of a frame form a dependency DAG.  The
// reference resolves it by decoding superblocks in order (and, across threads, by superblock-row
// progress counters: src/decode.c:2001-2090, src/thread_task.c:409-430).  Here the recorder hands
// over the operations in decode order and nothing is scheduled, sorted or levelled on the host.
// Per group of frames, all on the device and without a host round trip:
//   1. mark      a byte per 4x4 cell and plane counts the operations that will write the cell;
//   2. levels    one THREAD per operation, claimed in decode order: the operation's dependency
//                level is one more than the highest level among the writers of the cells it reads
//                (dav1d_prepare_intra_edges' own rules decide which cells those are).  Writers
//                publish their level in a second per-cell map, readers poll it: the levels spread
//                through the frame as a dataflow wave, no barrier anywhere;
//   3. sort      counting sort of the operations by (level, predictor class, size class);
//   4. execute   ONE persistent launch: CTAs claim runs of chunks of the sorted list with a ticket, a
//                chunk per warp - four small operations per warp (one per octet of lanes) or one large one - wait
//                until the count of every cell they read is at zero, predict, add the residual of
//                the transform pre-pass, store, and count their own cells down.  Claims follow the
//                level order, so whatever an operation waits for has been claimed before it by a
//                running warp: the launch cannot dead-lock whatever its residency, and because a
//                level's operations are independent the waits only show up in the thin tail of
//                the wavefront, where they cost one operation's latency per level instead of a
//                kernel launch or a grid barrier.
//   * the residuals do not depend on neighbours: a transform pre-pass (itx2.cu, next to the motion
//     compensation) leaves them in an int16 plane and the executor adds them to its predictions.
// A frame whose descriptors are inconsistent (operations that wait for each other or for a later
// one) runs into the bounded waits of steps 2 / 4: the context's status word is raised and
// dav1d_cuda_synchronize() reports -EIO; the launch always terminates.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include <cooperative_groups.h>
#include "ctx.h"
#include "itx2.cuh"
#include "ipred.cuh"
#include "mc.cuh"

namespace cg = cooperative_groups;

namespace d1 {

// defined in itx2.cu / mc.cu
int itx_batch_launch(const PicView &pic, const PicView *res, void *cf, const Dav1dCudaItxDesc *descs,
                     const int32_t *class_count, int zero_coefs, cudaStream_t st, const CoefFmt *fmt = nullptr);
int itx_task_launch(const PicView &pic, const PicView *res, void *cf, const Dav1dCudaItxDesc *descs,
                    const uint32_t *tasks, int n_small, int n_big, int zero_coefs, cudaStream_t st_small,
                    cudaStream_t st_big, const CoefFmt *fmt = nullptr);
int mc_obmc_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcDesc *descs,
                       const uint32_t *tiles, int n_tiles, cudaStream_t st);
void itx_init_attrs();
int mc_scaled_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcScaledDesc *descs, int n,
                         uint8_t *masks, cudaStream_t s);
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
    unsigned op_base;                    // number of the frame's first operation inside the group
    // cell maps (one allocation handed over by the caller): cnt = operations that still have to
    // write the cell, lvl = level + 1 of the cell's writer (bit 15: a residual-only operation will
    // write the cell once more).  Plane pl: mh[pl] rows of ms[pl] cells starting at cell mo[pl].
    uint8_t *cnt;
    uint16_t *lvl;
    int mw[3], mh[3], ms[3];
    unsigned mo[3];
    unsigned map_words;                  // size of the whole allocation in 32-bit words
    int recorded_levels;                 // the descriptors carry their dependency level (reserved = level + 1)
};
struct Intra2Args {
    Intra2Frame f[I2_MAXF];
    int nf;
    unsigned *status;                    // context status word, see ST_*
};
// status bits raised by the executor
constexpr unsigned ST_STUCK = 1u;        // operations that wait for each other (inconsistent descriptors)
constexpr unsigned ST_DEPTH = 2u;        // more dependency levels than the sort supports
constexpr unsigned ST_CELLS = 4u;        // a cell with more than two writers

struct MapGeo { int mw[3], mh[3], ms[3]; unsigned mo[3]; unsigned cells; };
static inline MapGeo map_geo(const int bw4, const int bh4, const int ss_hor, const int ss_ver) {
    MapGeo g;
    unsigned off = 0;
    for (int pl = 0; pl < 3; pl++) {
        g.mw[pl] = pl ? (bw4 + ss_hor) >> ss_hor : bw4;
        g.mh[pl] = pl ? (bh4 + ss_ver) >> ss_ver : bh4;
        g.ms[pl] = (g.mw[pl] + 3) & ~3;                  // rows start on a word
        g.mo[pl] = off;
        off += (unsigned)(g.ms[pl] * g.mh[pl]);
    }
    g.cells = (off + 15u) & ~15u;
    return g;
}

DEV unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
DEV unsigned ld_acquire_u8(const uint8_t *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
DEV unsigned ld_relaxed_u16(const uint16_t *p) {
    unsigned short v;
    asm volatile("ld.relaxed.gpu.global.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
    return v;
}
DEV void st_relaxed_u16(uint16_t *p, const unsigned v) {
    asm volatile("st.relaxed.gpu.global.u16 [%0], %1;" :: "l"(p), "h"((unsigned short)v) : "memory");
}

// cell j of the rectangle [x0, x0 + nx) x [y0, y0 + ny): rows, columns and power-of-two widths
// without a division (the general case is the source area of an intrabc block)
DEV void rect_cell(const int j, const int x0, const int nx, const int y0, const int ny, int *x, int *y) {
    if (ny == 1) { *x = x0 + j; *y = y0; }
    else if ((nx & (nx - 1)) == 0) { const int sh = 31 - __clz(nx); *x = x0 + (j & (nx - 1)); *y = y0 + (j >> sh); }
    else { *x = x0 + j % nx; *y = y0 + j / nx; }
}

// The cells an operation reads, as up to three rectangles of the cell maps: fn(plane, x0, x1, y0,
// y1, own) - own: the operation's own cells (a residual on top of an earlier operation's pixels:
// that one has to be done, the count stays at 1).  Exactly the pixels dav1d_prepare_intra_edges
// reads (ipred_prepare_tmpl.c:119-200), the co-located luma of CfL (recon_tmpl.c:1372-1417) and
// the source area of intrabc (:1624-1637).  *cls_out: predictor class 0..18 of the operation.
template <typename F>
HD void op_deps(const Intra2Frame &f, const Dav1dCudaIntraDesc &d, int *cls_out, F &&fn) {
    const int mode = d.mode;
    const int pl = d.plane;
    const int W = f.mw[pl], H = f.mh[pl];
    const int x0 = d.x4, y0 = d.y4;
    if (mode == DAV1D_CUDA_INTRA_PAL) { *cls_out = 15; return; }
    if (mode == DAV1D_CUDA_INTRA_NONE) {
        *cls_out = 18;
        fn(pl, x0, imin(x0 + d.tw4, W), y0, imin(y0 + d.th4, H), true);
        return;
    }
    if (mode == DAV1D_CUDA_INTRA_IBC) {
        *cls_out = 16;
        const int sx = (int16_t)(d.aux & 0xffff), sy = (int16_t)(d.aux >> 16);
        const int pw = 4 * W, ph = 4 * H;
        const int xa = iclip(sx, 0, pw - 1) >> 2, xb = iclip(sx + 4 * d.tw4 + (d.angle_delta ? 1 : 0) - 1, 0, pw - 1) >> 2;
        const int ya = iclip(sy, 0, ph - 1) >> 2, yb = iclip(sy + 4 * d.th4 + (d.flags ? 1 : 0) - 1, 0, ph - 1) >> 2;
        fn(pl, xa, xb + 1, ya, yb + 1, false);
        return;
    }
    const int have_left = x0 > d.tile_x4_start, have_top = y0 > d.tile_y4_start;
    int ang = mode == DAV1D_CUDA_INTRA_II ? 0 : d.angle_delta;
    const int rm = ipred_resolve_mode(mode == DAV1D_CUDA_INTRA_II ? d.angle_delta : mode == DAV1D_CUDA_INTRA_CFL ? 0 : mode,
                                      &ang, have_left, have_top);
    *cls_out = mode == DAV1D_CUDA_INTRA_II ? 17 : mode == DAV1D_CUDA_INTRA_CFL ? 14 : rm;
    const int needs = ipred_mode_needs(rm);
    if (have_top && ((needs & 2) || (needs & 4) || ((needs & 1) && !have_left))) {
        const bool tr = (needs & 8) && (d.edge_flags & 1);
        const int xs = ((needs & 4) && have_left) ? x0 - 1 : x0;
        int xe = (needs & 2) ? imin(x0 + d.tw4 + (tr ? d.tw4 : 0), d.tile_x4_end) : x0 + 1;
        xe = imin(xe, W);
        fn(pl, xs, xe, y0 - 1, y0, false);
    }
    if (have_left && ((needs & 1) || ((needs & 2) && !have_top) || ((needs & 4) && !have_top))) {
        const bool bl = (needs & 16) && (d.edge_flags & 8);
        int ye = (needs & 1) ? imin(y0 + d.th4 + (bl ? d.th4 : 0), d.tile_y4_end) : y0 + 1;
        ye = imin(ye, H);
        fn(pl, x0 - 1, x0, y0, ye, false);
    }
    if (mode == DAV1D_CUDA_INTRA_CFL) {    // the co-located luma
        const int sh = f.pic.ss_hor, sv = f.pic.ss_ver;
        fn(0, x0 << sh, imin((x0 + d.tw4) << sh, f.bw4), y0 << sv, imin((y0 + d.th4) << sv, f.bh4), false);
    }
}

// ---- one operation by a group of lanes (a warp, or an octet for operations of up to 64 pixels).
// Two compact stages per strip of up to 1024 pixels: (A) the predictor, one pixel per lane and
// step, into a 16-bit tile in shared memory - a single loop with the predictor switch inside, so
// the code every warp of an SM runs stays small; (B) eight pixels of a row per lane and step:
// tile (or current picture) + inter-intra blend + residual of the transform pre-pass -> picture.
template <int N> struct IntC { static constexpr int value = N; };
constexpr int EX_MAX_POLLS = 1 << 15;
// wait until the counts of the cells the operation reads are down (group-parallel polling)
DEV bool exec_wait(const Grp g, const Intra2Frame &f, const Dav1dCudaIntraDesc &d) {
    int cls, polls = 0;
    bool ok = true;
    op_deps(f, d, &cls, [&](const int pl, const int x0, const int x1, const int y0, const int y1, const bool own) {
        const int nx = x1 - x0, n = nx * (y1 - y0);
        const unsigned thr = own ? 1u : 0u;
        const uint8_t *m = f.cnt + f.mo[pl];
        const int S = f.ms[pl];
        for (int j = g.gl; j < n; j += g.G) {
            int x, y;
            rect_cell(j, x0, nx, y0, y1 - y0, &x, &y);
            const uint8_t *p = m + (size_t)y * S + x;
            // back off: many warps poll in the thin tail of the wavefront, the few that work need the issue slots
            unsigned ns = 32;
            while (ld_acquire_u8(p) > thr) {
                if (++polls > EX_MAX_POLLS) { ok = false; break; }
                __nanosleep(ns);
                ns = min(ns * 2u, 2048u);
            }
        }
    });
    return __all_sync(g.mask, ok);
}

// The operation's pixels are stored: one count less on each of its cells.  Every lane makes its own
// stores visible (fence), the group meets, then the counts go down: whoever sees a count at zero
// sees the pixels.
DEV void exec_done(const Grp g, const Intra2Frame &a, const Dav1dCudaIntraDesc &d) {
    const int pl = d.plane;
    __threadfence();
    grp_sync(g);
    {
        const int W = a.mw[pl], H = a.mh[pl], S = a.ms[pl];
        const int ltw = 31 - __clz((int)d.tw4);
        if (d.tw4 >= 4 && !(d.x4 & 3)) {
            // rows of whole words (tx blocks are aligned to their size, rows of the map to a word)
            const int lwpr = ltw - 2, nwd = d.th4 << lwpr;
            for (int j = g.gl; j < nwd; j += g.G) {
                const int cy = d.y4 + (j >> lwpr), cx = d.x4 + 4 * (j & ((1 << lwpr) - 1));
                if (cy < H && cx < W) {
                    const int nb = imin(4, W - cx);
                    atomicAdd((unsigned *)(a.cnt + a.mo[pl] + (size_t)cy * S + cx), 0u - (0x01010101u >> (8 * (4 - nb))));
                }
            }
        } else {
            const int nc = d.th4 << ltw;
            for (int j = g.gl; j < nc; j += g.G) {
                const int cx = d.x4 + (j & (d.tw4 - 1)), cy = d.y4 + (j >> ltw);
                if (cx < W && cy < H) {
                    const size_t off = (size_t)a.mo[pl] + (size_t)cy * S + cx;
                    atomicAdd((unsigned *)(a.cnt + (off & ~(size_t)3)), 0u - (1u << (8 * (off & 3))));
                }
            }
        }
    }
    grp_sync(g);
}

template <typename pixel>
__device__ __noinline__ bool intra_exec(const Grp g, const Intra2Frame &a, const Dav1dCudaIntraDesc *dp,
                                        pixel *edge, pixel *scratch, const int z2_centre, uint16_t *tile,
                                        const bool settled)
{
    // the descriptor in registers: five 64-bit loads, every lane of the group the same address
    Dav1dCudaIntraDesc d;
    {
        const uint2 *q = (const uint2 *)dp;
        uint2 *o = (uint2 *)&d;
#pragma unroll
        for (int k = 0; k < 5; k++) o[k] = __ldg(q + k);
    }
    const int pl = d.plane;
    const int ss_hor = pl ? a.pic.ss_hor : 0, ss_ver = pl ? a.pic.ss_ver : 0;
    const PlaneView &pv = a.pic.p[pl];
    const int stride = (int)(pv.stride / (int)sizeof(pixel));
    pixel *dst = (pixel *)pv.data + (int64_t)d.y4 * 4 * stride + d.x4 * 4;
    const int w = d.tw4 * 4, h = d.th4 * 4;
    const int lw = 31 - __clz(w);
    const int bdmax = a.pic.bdmax;
    const int mode = d.mode;
    const bool has_res = d.eob >= 0 && mode != DAV1D_CUDA_INTRA_PAL;
    const int rps = imin(h, 1024 >> lw);           // rows of a strip
    const int16_t *res = nullptr;
    int rstride = 0;
    // the lane's first residuals: loaded now, used after the wait and the prediction
    uint4 r0 = make_uint4(0u, 0u, 0u, 0u);
    if (has_res) {
        const PlaneView &rv = a.res.p[pl];
        rstride = (int)(rv.stride / 2);
        res = (const int16_t *)rv.data + (int64_t)d.y4 * 4 * rstride + d.x4 * 4;
        const int i = g.gl * 4;
        if (i < w * rps) {
            const int16_t *rp = res + (i >> lw) * rstride + (i & (w - 1));
            const uint2 t = __ldcg((const uint2 *)rp);
            r0.x = t.x; r0.y = t.y;
        }
    }
    // settled: every lower level is complete - nothing this operation reads can still change
    const bool ok = settled || exec_wait(g, a, d);
    if (ok) {
        const int have_left = d.x4 > d.tile_x4_start, have_top = d.y4 > d.tile_y4_start;
        // stage A: 0 predictor (P), 1 intrabc, 2 nothing (the tile is filled / the current picture is the source)
        int kind = 0;
        PixParams<pixel> P;
        P.pm = PM_CONST; P.p0 = P.p1 = P.p2 = P.p3 = 0; P.edge = P.e0 = P.e1 = edge; P.tile = tile; P.w = w; P.h = h;
        const uint8_t *bmask = nullptr;                // inter-intra blend mask
        if (mode == DAV1D_CUDA_INTRA_PAL) {
            P.pm = PM_PAL; P.tile = a.pal_idx + d.coef_off; P.e0 = (const pixel *)a.pal + d.aux;
        } else if (mode == DAV1D_CUDA_INTRA_IBC) {
            kind = 1;
        } else if (mode == DAV1D_CUDA_INTRA_NONE) {
            kind = 2;
        } else {
            // CfL (recon_tmpl.c:1372-1417): ac from the co-located luma, DC edges; inter-intra
            // (:1658-1681): the whole-block intra prediction (edge_flags 0, no edge filter) is blended
            // onto the inter prediction that is in dst
            const bool cfl = mode == DAV1D_CUDA_INTRA_CFL, ii = mode == DAV1D_CUDA_INTRA_II;
            if (cfl) {
                const PlaneView &lv = a.pic.p[0];
                const int lstride = (int)(lv.stride / (int)sizeof(pixel));
                const pixel *luma = (const pixel *)lv.data + (int64_t)((d.y4 * 4) << ss_ver) * lstride + ((d.x4 * 4) << ss_hor);
                cfl_ac_block<pixel>(g, (int16_t *)tile, luma, lstride, d.aux & 0xff, (d.aux >> 8) & 0xff, w, h, ss_hor, ss_ver);
            }
            int angle = (cfl || ii) ? 0 : d.angle_delta;
            const int m = prepare_edges<pixel>(g, d.x4, have_left, d.y4, have_top, d.tile_x4_end, d.tile_y4_end,
                                               (cfl || ii) ? 0 : d.edge_flags, dst, stride, nullptr,
                                               cfl ? 0 : ii ? d.angle_delta : mode, &angle, d.tw4, d.th4,
                                               (cfl || ii) ? 0 : (d.flags >> 10) & 1, edge, bdmax);
            if (cfl) {
                P = cfl_setup<pixel>(g, m, edge, w, h, (const int16_t *)tile, d.angle_delta, bdmax);
            } else {
                if (ii) { bmask = a.pal_idx + d.coef_off; angle = 0; }
                else angle |= d.flags;
                const int max_w = ((4 * a.bw4 + ss_hor) >> ss_hor) - 4 * d.x4;
                const int max_h = ((4 * a.bh4 + ss_ver) >> ss_ver) - 4 * d.y4;
                P = ipred_setup<pixel, uint16_t>(g, m, angle, w, h, max_w, max_h, edge, scratch, tile, bdmax, z2_centre);
            }
        }
        grp_sync(g);
        const bool from_dst = mode == DAV1D_CUDA_INTRA_NONE;
        for (int y0 = 0; y0 < h; y0 += rps) {
            const int n = rps << lw;
            if (kind == 1) {
                // intrabc: put_bilin (mc_tmpl.c:395-450) from the current picture; coordinates clamped to
                // the 4*bw4 x 4*bh4 area (= emu_edge, recon_tmpl.c:974-995)
                const pixel *base = (const pixel *)pv.data;
                const int ib = PxTraits<pixel>::inter_bits(bdmax);
                const int sx = (int16_t)(d.aux & 0xffff), sy = (int16_t)(d.aux >> 16);
                const int mx = d.angle_delta, my = d.flags;
                const int pw = (4 * a.bw4) >> ss_hor, ph = (4 * a.bh4) >> ss_ver;
#pragma unroll 1
                for (int i = g.gl; i < n; i += g.G) {
                    const int y = y0 + (i >> lw), x = i & (w - 1);
                    const int xa = iclip(sx + x, 0, pw - 1), xb = iclip(sx + x + 1, 0, pw - 1);
                    const int ya = iclip(sy + y, 0, ph - 1), yb = iclip(sy + y + 1, 0, ph - 1);
                    const int q00 = __ldcg(base + (int64_t)ya * stride + xa), q01 = __ldcg(base + (int64_t)ya * stride + xb);
                    const int q10 = __ldcg(base + (int64_t)yb * stride + xa), q11 = __ldcg(base + (int64_t)yb * stride + xb);
                    int v = q00;
                    if (mx && my) {
                        const int sh1 = 4 - ib, r1 = (1 << sh1) >> 1;
                        const int m0 = (16 * q00 + mx * (q01 - q00) + r1) >> sh1;
                        const int m1 = (16 * q10 + mx * (q11 - q10) + r1) >> sh1;
                        const int sh2 = 4 + ib;
                        v = clip_px<pixel>((16 * m0 + my * (m1 - m0) + ((1 << sh2) >> 1)) >> sh2, bdmax);
                    } else if (mx) {
                        const int sh1 = 4 - ib;
                        const int px = (16 * q00 + mx * (q01 - q00) + ((1 << sh1) >> 1)) >> sh1;
                        v = clip_px<pixel>((px + ((1 << ib) >> 1)) >> ib, bdmax);
                    } else if (my) {
                        v = clip_px<pixel>((16 * q00 + my * (q10 - q00) + 8) >> 4, bdmax);
                    }
                    tile[i] = (uint16_t)v;
                }
                P.pm = PM_TILE; P.tile = tile;
                grp_sync(g);
            }
            auto stage_b = [&](auto pwc) {
                constexpr int PW = decltype(pwc)::value;
#pragma unroll 1
                for (int i = g.gl * PW; i < n; i += g.G * PW) {
                    const int y = y0 + (i >> lw), x = i & (w - 1);
                    pixel *p = dst + y * stride + x;
                    int v[PW];
                    if (from_dst) load_px<pixel, PW>(p, v);
                    else ipred_seg<pixel, PW, uint16_t>(P, x, y, kind == 1 ? i : (y << lw) + x, bdmax, v);
                    if (bmask) {
                        // mc.blend (mc_tmpl.c:642-653) of the intra prediction onto the inter prediction
                        int c[PW];
                        load_px<pixel, PW>(p, c);
                        const uint8_t *mk = bmask + (y << lw) + x;
#pragma unroll
                        for (int k = 0; k < PW; k++) v[k] = (c[k] * (64 - mk[k]) + v[k] * mk[k] + 32) >> 6;
                    }
                    if (has_res) {
                        uint4 r = r0;
                        if (y0 || i != g.gl * PW) {
                            const int16_t *rp = res + y * rstride + x;
                            if (PW == 8) r = __ldcg((const uint4 *)rp);
                            else { const uint2 t = __ldcg((const uint2 *)rp); r.x = t.x; r.y = t.y; }
                        }
                        v[0] = clip_px<pixel>(v[0] + (int)(int16_t)(r.x & 0xffff), bdmax);
                        v[1] = clip_px<pixel>(v[1] + ((int)r.x >> 16), bdmax);
                        v[2] = clip_px<pixel>(v[2] + (int)(int16_t)(r.y & 0xffff), bdmax);
                        v[3] = clip_px<pixel>(v[3] + ((int)r.y >> 16), bdmax);
                        if (PW == 8) {
                            v[PW - 4] = clip_px<pixel>(v[PW - 4] + (int)(int16_t)(r.z & 0xffff), bdmax);
                            v[PW - 3] = clip_px<pixel>(v[PW - 3] + ((int)r.z >> 16), bdmax);
                            v[PW - 2] = clip_px<pixel>(v[PW - 2] + (int)(int16_t)(r.w & 0xffff), bdmax);
                            v[PW - 1] = clip_px<pixel>(v[PW - 1] + ((int)r.w >> 16), bdmax);
                        }
                    }
                    store_px<pixel, PW>(p, v);
                }
            };
            stage_b(IntC<4>());
            grp_sync(g);
        }
    }
    exec_done(g, a, d);
    return ok;
}

// Per-warp shared memory of the executor.  A warp works on one chunk at a time, split into groups
// of G lanes, one operation per group (the operations of a chunk share their size class):
//   G = 1   4x4                 edge[-8..8], 24 pixels of Z-mode scratch, tile of 16
//   G = 4   up to 64 pixels     edge[-32..32] in 72 pixels, 56 of scratch, tile of 64
//   G = 8   up to 256 pixels    edge[-64..64] in 136 pixels, 80 of scratch, tile of 256
//   G = 16  up to 512 pixels    as G = 8, tile of 512
//   G = 32  anything            EDGE_BUF + IPRED_SCRATCH pixels, tile of 1024
// The per-lane regions of G = 1 are an odd number of words apart (no bank conflicts when the 32
// lanes walk their own edges in step).
struct GrpSmem { int es_stride, centre, scr, z2c, tile_stride; };
template <typename pixel> DEV GrpSmem grp_smem(const int G) {
    if (G == 1) return GrpSmem{ sizeof(pixel) == 2 ? 50 : 52, 8, 26, 8, 20 };
    if (G == 4) return GrpSmem{ 128, 36, 72, 24, 64 };
    if (G == 8) return GrpSmem{ 216, 68, 136, 40, 256 };
    if (G == 16) return GrpSmem{ 216, 68, 136, 40, 512 };
    return GrpSmem{ 0, EDGE_C, EDGE_BUF, 128 + 8, 0 };
}
constexpr int ES_PX = 32 * 52;
template <typename pixel> struct __align__(16) ExecSmem {
    pixel es[ES_PX];
    uint16_t tile[32 * 32];
};
constexpr int R_WARPS = 8;
// operation id: frame (6 bits) | index inside the frame (20 bits); EMPTY = no operation in the slot
constexpr int OP_FRAME_SHIFT = 20;
constexpr unsigned OP_EMPTY = 0xffffffffu;
DEV int op_frame(const unsigned id) { return (int)((id >> OP_FRAME_SHIFT) & 63u); }
DEV int op_index(const unsigned id) { return (int)(id & ((1u << OP_FRAME_SHIFT) - 1u)); }
// Sort key: level * 256 + bin, bin = size class * 32 + predictor class.  Size classes: 0 4x4, 1 up
// to 64 pixels, 2 up to 256, 3 up to 512, 4 larger; predictor class: the DSP-table predictor 0..13,
// CfL, palette, intrabc, inter-intra, residual only - so the groups of a warp and the warps of an SM
// mostly run the same code path.
// Slots: the sorted list is cut into chunks of 32 slots, one chunk per warp and claim; an operation
// of size class s takes 1, 4, 8, 16 or 32 slots (the first one holds its id), i.e. as many slots as
// it gets lanes; the regions of a level are padded to whole chunks.  Levels with few operations
// (the thin tail of the wavefront, where latency is all that counts) give every operation a warp.
constexpr int N_BINS = 256;
constexpr int MAX_LEVELS = 8192;
constexpr int HIST_SMEM_LEVELS = 16;     // levels whose bins a block counts in shared memory
constexpr unsigned THIN_LEVEL_OPS = 512;
DEV int op_bin(const Dav1dCudaIntraDesc &d, const int cls) {
    const int c4 = d.tw4 * d.th4;
    // cls: 0..13 DSP-table predictor, 14 CfL, 15 palette, 16 intrabc, 17 inter-intra, 18 residual only
    return (c4 <= 1 ? 0 : c4 <= 4 ? 32 : c4 <= 16 ? 64 : c4 <= 32 ? 96 : 128) + cls;
}
// slots (= lanes) an operation of this bin takes
DEV unsigned bin_slots(const unsigned bin, const bool thin) {
    const unsigned sc = bin >> 5;
    return thin || sc >= 4 ? 32u : sc == 0 ? 1u : 2u << sc;
}

// control block of a group's executor (device memory, cleared before the mark launch)
struct ExecCtl {
    unsigned ticket;                    // next chunk to claim (a CTA takes R_WARPS at a time)
    unsigned n_chunks;                  // written by the scan
    unsigned max_level;
    int levels_done;                    // all chunks of the levels up to this one are complete (-1: none yet)
};
struct SchedArgs {
    Intra2Args g;
    ExecCtl *ctl;
    unsigned *key;                      // per operation of the group: level * 32 + bin
    unsigned *bins;                     // MAX_LEVELS * N_BINS counters -> slot offsets -> cursors
    uint8_t *thin;                      // per level: every operation gets a whole warp
    unsigned *lvl_end;                  // per level: one past its last chunk
    unsigned *lvl_left;                 // per level: chunks that are not complete yet
    unsigned *slots;                    // the sorted list: operation ids, OP_EMPTY padding
    unsigned n_slots_cap;
};

// ---- 1. mark: every operation adds one to each of its cells (launched after the maps are cleared)
__global__ void intra_clear_kernel(const __grid_constant__ Intra2Args a) {
    const Intra2Frame &f = a.f[blockIdx.y];
    if (f.n_ops <= 0) return;
    uint4 *p = (uint4 *)f.cnt;
    const unsigned n = f.map_words >> 2;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        p[i] = make_uint4(0u, 0u, 0u, 0u);
}
__global__ void intra_mark_kernel(const __grid_constant__ Intra2Args a) {
    const Intra2Frame &f = a.f[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < f.n_ops; i += gridDim.x * blockDim.x) {
        const Dav1dCudaIntraDesc &d = f.descs[i];
        const int pl = d.plane, W = f.mw[pl], H = f.mh[pl], S = f.ms[pl];
        const int x4 = d.x4, tw4 = d.tw4, y1 = imin(d.y4 + d.th4, H);
        if (tw4 >= 4 && !(x4 & 3)) {
            for (int y = d.y4; y < y1; y++)
                for (int x = x4; x < imin(x4 + tw4, W); x += 4) {
                    const int nb = imin(4, W - x);
                    atomicAdd((unsigned *)(f.cnt + f.mo[pl] + (size_t)y * S + x), 0x01010101u >> (8 * (4 - nb)));
                }
        } else {
            for (int y = d.y4; y < y1; y++)
                for (int x = x4; x < imin(x4 + tw4, W); x++) {
                    const size_t off = (size_t)f.mo[pl] + (size_t)y * S + x;
                    atomicAdd((unsigned *)(f.cnt + (off & ~(size_t)3)), 1u << (8 * (off & 3)));
                }
        }
    }
}

// ---- 2. levels.  A cluster of LV_CL blocks per frame, every thread owns the operations t, t + T,
// t + 2T, ... (T threads per frame).  Rounds: a thread looks at each of its unresolved operations;
// one whose cells all carry the level of their writer (or have no writer) takes its own level and
// publishes it; a cluster barrier ends the round.  Nobody spins: a round is a sweep over what is
// left, and the sweeps of all frames run side by side.
constexpr int LV_THREADS = 1024, LV_CL = 4;
constexpr unsigned LV_FLAG = 0x8000u;
__global__ void __cluster_dims__(LV_CL, 1, 1) __launch_bounds__(LV_THREADS, 1)
intra_levels_kernel(const __grid_constant__ SchedArgs a) {
    __shared__ unsigned s_flags[2];
    cg::cluster_group cluster = cg::this_cluster();
    const Intra2Frame &f = a.g.f[blockIdx.y];
    if (f.recorded_levels) return;       // (the whole cluster)
    constexpr int T = LV_THREADS * LV_CL;
    const int t = (int)(blockIdx.x * LV_THREADS + threadIdx.x);
    unsigned max_level = 0, bad = 0;
    int round = 0;
    // 64 operations per thread at a time (decode order is a topological order: a later part of the
    // frame never feeds an earlier one)
    for (int base = 0; base < f.n_ops; base += 64 * T) {
        unsigned long long todo = 0;
        for (int k = 0; k < 64; k++)
            if (base + k * T + t < f.n_ops) todo |= 1ull << k;
        bool first = true;
        for (;; round++) {
            bool progress = false;
            // an unresolved operation remembers one cell it was seen waiting for (in its key slot):
            // only when that cell carries its level is the operation looked at again
            unsigned long long cand = todo;
            if (!first) {
                for (unsigned long long m = todo; m; m &= m - 1) {
                    const int k = __ffsll((long long)m) - 1;
                    const unsigned kb = __ldcg(a.key + f.op_base + (unsigned)(base + k * T + t));
                    const unsigned v = __ldcg(f.lvl + (kb & 0x3fffffffu));
                    if (v == 0 || (!(kb & 0x40000000u) && (v & LV_FLAG))) cand &= ~(1ull << k);
                }
            }
            first = false;
            for (unsigned long long m = cand; m; m &= m - 1) {
                const int k = __ffsll((long long)m) - 1;
                const int i = base + k * T + t;
                const Dav1dCudaIntraDesc d = f.descs[i];
                int cls = 0;
                unsigned level = 0, blocker = 0;
                bool ready = true;
                op_deps(f, d, &cls, [&](const int pl, const int x0, const int x1, const int y0, const int y1, const bool own) {
                    if (!ready) return;
                    const unsigned cb = f.mo[pl];
                    const int S = f.ms[pl];
                    const int nx = x1 - x0, n = nx * (y1 - y0);
                    // four cells at a time: their loads are in flight together
                    for (int j0 = 0; j0 < n && ready; j0 += 4) {
                        unsigned cc[4], cn[4], v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            int x, y;
                            rect_cell(imin(j0 + u, n - 1), x0, nx, y0, y1 - y0, &x, &y);
                            const unsigned c = cb + (unsigned)(y * S + x);
                            cc[u] = c;
                            cn[u] = __ldcg(f.cnt + c);
                            v[u] = __ldcg(f.lvl + c);
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (cn[u] == 0 || (own && cn[u] == 1)) continue;     // nobody (else) writes the cell in this phase
                            if (cn[u] > 2) bad |= ST_CELLS;
                            if (v[u] == 0 || (!own && (v[u] & LV_FLAG))) { ready = false; blocker = cc[u] | (own ? 0x40000000u : 0u); }
                            level = max(level, v[u] & (LV_FLAG - 1u));
                        }
                    }
                });
                if (!ready) { a.key[f.op_base + (unsigned)i] = blocker; continue; }
                if (level >= (unsigned)MAX_LEVELS) { bad |= ST_DEPTH; level = MAX_LEVELS - 1; }
                // publish: level + 1 in the operation's cells; flagged where a residual-only operation follows
                const int pl = d.plane, W = f.mw[pl], H = f.mh[pl], S = f.ms[pl];
                const bool primary = d.mode != DAV1D_CUDA_INTRA_NONE;
                for (int y = d.y4; y < imin(d.y4 + d.th4, H); y++)
                    for (int x = d.x4; x < imin(d.x4 + d.tw4, W); x++) {
                        const unsigned c = f.mo[pl] + (unsigned)(y * S + x);
                        const bool more = primary && __ldcg(f.cnt + c) > 1;
                        f.lvl[c] = (uint16_t)((level + 1u) | (more ? LV_FLAG : 0u));
                    }
                a.key[f.op_base + (unsigned)i] = level * N_BINS + (unsigned)op_bin(d, cls);
                max_level = max(max_level, level);
                todo &= ~(1ull << k);
                progress = true;
            }
            // end of the round: does anybody in the cluster have work left, did anybody get on?
            const int any_p = __syncthreads_or(progress), any_t = __syncthreads_or(todo != 0);
            if (threadIdx.x == 0) s_flags[round & 1] = (any_p ? 1u : 0u) | (any_t ? 2u : 0u);
            __threadfence();
            cluster.sync();
            unsigned all = 0;
            for (int r = 0; r < LV_CL; r++) all |= *cluster.map_shared_rank(&s_flags[round & 1], r);
            if (!(all & 2u)) { round++; break; }
            if (!(all & 1u)) {              // operations left, none resolved: they wait for each other
                bad |= ST_STUCK;
                // their keys still have to be valid: level 0
                for (unsigned long long m = todo; m; m &= m - 1) {
                    const int i = base + (__ffsll((long long)m) - 1) * T + t;
                    const Dav1dCudaIntraDesc d = f.descs[i];
                    int cls = 0;
                    op_deps(f, d, &cls, [&](int, int, int, int, int, bool) {});
                    a.key[f.op_base + (unsigned)i] = (unsigned)op_bin(d, cls);
                }
                round++;
                break;
            }
        }
    }
    if (max_level > __ldcg(&a.ctl->max_level)) atomicMax(&a.ctl->max_level, max_level);
    if (bad) atomicOr(a.g.status, bad);
    cluster.sync();      // no block leaves while its flags may still be read
}

// ---- 2'. frames whose recorder assigned the levels (dav1d_cuda_intra_levels(), a linear pass in
// decode order on the host): only the sort keys are left to do
__global__ void __launch_bounds__(256) intra_keys_kernel(const __grid_constant__ SchedArgs a) {
    const Intra2Frame &f = a.g.f[blockIdx.y];
    if (!f.recorded_levels) return;
    unsigned max_level = 0, bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < f.n_ops; i += gridDim.x * blockDim.x) {
        const Dav1dCudaIntraDesc d = f.descs[i];
        int cls = 0;
        op_deps(f, d, &cls, [&](int, int, int, int, int, bool) {});
        unsigned level = (d.reserved & 0xffffu) - 1u;
        if ((d.reserved & 0xffffu) == 0) { bad |= ST_STUCK; level = 0; }       // not a recorded level
        if (level >= (unsigned)MAX_LEVELS) { bad |= ST_DEPTH; level = MAX_LEVELS - 1; }
        a.key[f.op_base + (unsigned)i] = level * N_BINS + (unsigned)op_bin(d, cls);
        max_level = max(max_level, level);
    }
    if (max_level > __ldcg(&a.ctl->max_level)) atomicMax(&a.ctl->max_level, max_level);
    if (bad) atomicOr(a.g.status, bad);
}

// ---- 3. sort.  COUNT: histogram of the keys; scan: counters -> first slot of every bin; !COUNT:
// every operation takes its slot.  Blocks aggregate the bins of the first levels in shared memory.
template <bool COUNT>
__global__ void __launch_bounds__(256) intra_sort_kernel(const __grid_constant__ SchedArgs a) {
    __shared__ unsigned s_cnt[HIST_SMEM_LEVELS * N_BINS];
    const Intra2Frame &f = a.g.f[blockIdx.y];
    const int i0 = (int)blockIdx.x * 1024;
    if (i0 >= f.n_ops) return;
    for (int q = threadIdx.x; q < HIST_SMEM_LEVELS * N_BINS; q += blockDim.x) s_cnt[q] = 0;
    __syncthreads();
    unsigned key[4], rank[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int i = i0 + u * 256 + (int)threadIdx.x;
        key[u] = OP_EMPTY;
        rank[u] = 0;
        if (i < f.n_ops) {
            key[u] = __ldcg(a.key + f.op_base + (unsigned)i);
            const unsigned step = COUNT ? 1u : bin_slots(key[u] & (N_BINS - 1), __ldcg(a.thin + key[u] / N_BINS) != 0);
            if (key[u] < HIST_SMEM_LEVELS * N_BINS) rank[u] = atomicAdd(&s_cnt[key[u]], step);
            else if (COUNT) atomicAdd(a.bins + key[u], 1u);
            else rank[u] = atomicAdd(a.bins + key[u], step);
        }
    }
    __syncthreads();
    // the block's share of every bin it met: counted (COUNT) or reserved (-> s_cnt = first slot)
    for (int q = threadIdx.x; q < HIST_SMEM_LEVELS * N_BINS; q += blockDim.x) {
        const unsigned c = s_cnt[q];
        if (c) s_cnt[q] = atomicAdd(a.bins + q, c);
    }
    if (COUNT) return;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; u++) {
        if (key[u] == OP_EMPTY) continue;
        const unsigned pos = rank[u] + (key[u] < HIST_SMEM_LEVELS * N_BINS ? s_cnt[key[u]] : 0u);
        const unsigned id = ((unsigned)blockIdx.y << OP_FRAME_SHIFT) | (unsigned)(i0 + u * 256 + (int)threadIdx.x);
        const unsigned ns = bin_slots(key[u] & (N_BINS - 1), __ldcg(a.thin + key[u] / N_BINS) != 0);
        if (pos + ns > a.n_slots_cap) continue;
        if (ns == 1) {
            a.slots[pos] = id;
        } else {
            *(uint4 *)(a.slots + pos) = make_uint4(id, OP_EMPTY, OP_EMPTY, OP_EMPTY);
            for (unsigned q = 4; q < ns; q += 4) *(uint4 *)(a.slots + pos + q) = make_uint4(OP_EMPTY, OP_EMPTY, OP_EMPTY, OP_EMPTY);
        }
    }
}
// Slots of a level: one region per size class, each padded to whole chunks with empty slots.
__global__ void __launch_bounds__(1024) intra_scan_kernel(const __grid_constant__ SchedArgs a) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_lev = (int)min(a.ctl->max_level + 1u, (unsigned)MAX_LEVELS);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int l0 = 0; l0 < n_lev; l0 += 1024) {
        const int l = l0 + tid;
        unsigned nc[5] = { 0, 0, 0, 0, 0 };
        if (l < n_lev) {
            const uint4 *p = (const uint4 *)(a.bins + (size_t)l * N_BINS);
#pragma unroll
            for (int sc = 0; sc < 5; sc++)
#pragma unroll 4
                for (int q = 0; q < 8; q++) { const uint4 v = p[sc * 8 + q]; nc[sc] += v.x + v.y + v.z + v.w; }
        }
        const bool thin = nc[0] + nc[1] + nc[2] + nc[3] + nc[4] < THIN_LEVEL_OPS;
        unsigned mine = 0;
#pragma unroll
        for (int sc = 0; sc < 5; sc++) mine += (bin_slots(sc * 32, thin) * nc[sc] + 31u) & ~31u;
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned before = s_carry;
        for (int q = 0; q < warp; q++) before += s_warp[q];
        unsigned run = before + incl - mine;          // first slot of level l
        if (l < n_lev) {
            a.thin[l] = thin ? 1 : 0;
            a.lvl_end[l] = (run + mine) >> 5;
            a.lvl_left[l] = mine >> 5;
            unsigned *b = a.bins + (size_t)l * N_BINS;
            for (int sc = 0; sc < 5; sc++) {
                const unsigned per = bin_slots(sc * 32, thin), end = run + ((per * nc[sc] + 31u) & ~31u);
                for (int q = sc * 32; q < sc * 32 + 32; q++) { const unsigned c = b[q]; b[q] = run; run += per * c; }
                for (; run < end; run++)
                    if (run < a.n_slots_cap) a.slots[run] = OP_EMPTY;
            }
        }
        __syncthreads();
        if (tid == 1023) s_carry = before + incl;
        __syncthreads();
    }
    if (tid == 0) {
        a.ctl->n_chunks = min(s_carry, a.n_slots_cap) >> 5;
        a.ctl->ticket = 0;
        a.ctl->levels_done = -1;
    }
}

// ---- 4. execute
template <typename pixel>
__global__ void __launch_bounds__(R_WARPS * 32, 3) intra_exec_kernel(const __grid_constant__ SchedArgs a) {
    extern __shared__ __align__(16) uint8_t exec_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ExecSmem<pixel> *sm = (ExecSmem<pixel> *)exec_smem_raw + warp;
    const unsigned n_chunks = __ldcg(&a.ctl->n_chunks);
    // a CTA claims R_WARPS consecutive chunks at a time (neighbours in the sorted list: the same size class and
    // predictor, the same frame), warp w takes the w-th; the next claim is in flight while the CTA works
    __shared__ unsigned s_base[2];
    if (threadIdx.x == 0) s_base[0] = atomicAdd(&a.ctl->ticket, (unsigned)R_WARPS);
    __syncthreads();
    unsigned round = 0;
    const int n_lev = (int)min(__ldcg(&a.ctl->max_level) + 1u, (unsigned)MAX_LEVELS);
    int level = 0, known_done = -1;
    unsigned level_end = __ldcg(a.lvl_end);
    for (;;) {
        const unsigned base = s_base[round & 1u];
        if (base >= n_chunks) break;                         // the same for every warp of the CTA
        if (threadIdx.x == 0) s_base[(round + 1u) & 1u] = atomicAdd(&a.ctl->ticket, (unsigned)R_WARPS);
        round++;
        const unsigned c = base + (unsigned)warp;
        if (c >= n_chunks) { __syncthreads(); continue; }
        const unsigned sl = __ldcg(a.slots + 32 * c + lane);
        // the chunk's level (claims only move forward); are all lower levels complete?  Then nothing
        // the chunk's operations read can still change and they need not look at the cell counts.
        while (c >= level_end && level + 1 < n_lev) level_end = __ldcg(a.lvl_end + ++level);
        if (known_done < level - 1) known_done = (int)ld_acquire_u32((const unsigned *)&a.ctl->levels_done);
        // far ahead of the wavefront (the thin tail: more warps than work): sleep on the one progress word
        // instead of polling cells; the cell counts are only looked at within one level of the front
        for (int naps = 0; known_done < level - 2 && naps < (1 << 15); naps++) {
            __nanosleep(1500);
            known_done = (int)ld_acquire_u32((const unsigned *)&a.ctl->levels_done);
        }
        const bool settled = known_done >= level - 1;
        bool ok = true;
        if (__ballot_sync(0xffffffffu, sl != OP_EMPTY)) {
            // lanes per operation = spacing of the used slots (a lone operation gets the whole warp)
            const unsigned at = __reduce_or_sync(0xffffffffu, sl != OP_EMPTY ? (unsigned)lane : 0u);
            const int G = at ? (int)(at & (0u - at)) : 32;
            const int gi = lane / G;
            const unsigned id = __shfl_sync(0xffffffffu, sl, gi * G);
            if (id != OP_EMPTY) {
                const Grp g = Grp{ lane & (G - 1), G, G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (lane & ~(G - 1)) };
                const GrpSmem gs = grp_smem<pixel>(G);
                const Intra2Frame &f = a.g.f[op_frame(id)];
                pixel *es = sm->es + gi * gs.es_stride;
                ok = intra_exec<pixel>(g, f, f.descs + op_index(id), es + gs.centre, es + gs.scr, gs.z2c,
                                       sm->tile + gi * gs.tile_stride, settled);
            }
        }
        if (!ok) atomicOr(a.g.status, ST_STUCK);
        __syncwarp();
        // the chunk is complete (every lane fenced its stores before its cells were counted down):
        // one chunk less in its level; whoever completes a level moves levels_done forward
        if (lane == 0 && atomicSub(a.lvl_left + level, 1u) == 1u) {
            __threadfence();
            int d = (int)ld_acquire_u32((const unsigned *)&a.ctl->levels_done);
            while (d + 1 < n_lev && ld_acquire_u32(a.lvl_left + d + 1) == 0u) d++;
            atomicMax(&a.ctl->levels_done, d);
        }
        __syncthreads();                                     // the round is over: the next claim is visible
    }
}

static int g_exec_blocks[2] = { 0, 0 };   // resident blocks of the executor per pixel type
template <typename pixel> static size_t exec_smem_bytes() { return R_WARPS * sizeof(ExecSmem<pixel>); }

void recon_init_attrs() {
    itx_init_attrs();
    cudaFuncSetAttribute(intra_exec_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)exec_smem_bytes<uint8_t>());
    cudaFuncSetAttribute(intra_exec_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)exec_smem_bytes<uint16_t>());
    cudaFuncSetAttribute(intra_exec_kernel<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(intra_exec_kernel<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, intra_exec_kernel<uint8_t>, R_WARPS * 32, exec_smem_bytes<uint8_t>());
    g_exec_blocks[0] = std::max(1, occ) * std::max(1, sms);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, intra_exec_kernel<uint16_t>, R_WARPS * 32, exec_smem_bytes<uint16_t>());
    g_exec_blocks[1] = std::max(1, occ) * std::max(1, sms);
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

// Workspace of a group's scheduling passes: control block, the sort's bins, one key per operation
// and the sorted slot list.  One per context, grown on demand outside any stream capture;
// submissions of a context are ordered on its stream.
static size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t ws_hdr_bytes() { return al256(sizeof(ExecCtl)) + al256((size_t)MAX_LEVELS * N_BINS * 4) + al256(MAX_LEVELS) + 2 * al256((size_t)MAX_LEVELS * 4); }
static size_t ws_slots_cap(size_t total) { return 32 * total + 160 * (size_t)MAX_LEVELS; }
static size_t ws_need(size_t total) { return ws_hdr_bytes() + al256(total * 4) + al256(ws_slots_cap(total) * 4); }
static size_t group_ops(const Dav1dCudaReconBatch *const *bs, int n) {
    size_t total = 0;
    for (int f = 0; f < n; f++)
        if (bs[f] && bs[f]->n_intra > 0) total += (size_t)bs[f]->n_intra;
    return total;
}
static int ensure_rounds_ws(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n) {
    const size_t need = ws_need(group_ops(bs, n));
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
        for (int k = 0; k < 4; k++)
            if (b->n_mc_scaled[k] < 0 || (b->n_mc_scaled[k] > 0 && !b->mc_scaled)) return -22;
        if (b->n_cf_esc < 0 || (b->n_cf_esc > 0 && !b->cf_esc)) return -22;
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
    constexpr int NS = 1 + Dav1dCudaContext::N_AUX;
    cudaStream_t ss[NS];
    ss[0] = st;
    for (int i = 0; i < Dav1dCudaContext::N_AUX; i++) ss[1 + i] = c->aux[i];
    const bool hbd = bs[0]->dst->bitdepth_max > 0xff;
    // intra: the scheduling passes (cell maps, levels, sort) touch nothing the other phases use and
    // run on their own branch next to them
    SchedArgs sa;
    memset(&sa, 0, sizeof(sa));
    Intra2Args &ia = sa.g;
    int n_ops = 0, n_rec = 0;
    size_t total = 0;
    if (mask & 16) {
        ia.nf = n;
        for (int f = 0; f < n; f++) {
            const Dav1dCudaReconBatch *b = bs[f];
            Intra2Frame &p = ia.f[f];
            p.pic = pic_view(b->dst); p.bw4 = b->bw4; p.bh4 = b->bh4; p.cf = b->cf;
            if (b->intra_res) p.res = pic_view(b->intra_res);
            p.descs = b->intra; p.pal = b->pal; p.pal_idx = b->pal_idx;
            p.n_ops = b->n_intra > 0 ? b->n_intra : 0;
            p.recorded_levels = b->intra_levels_recorded;
            n_rec += b->intra_levels_recorded ? 1 : 0;
            p.op_base = (unsigned)total;
            total += (size_t)p.n_ops;
            const MapGeo mg = map_geo(b->bw4, b->bh4, b->dst->ss_hor, b->dst->ss_ver);
            for (int pl = 0; pl < 3; pl++) { p.mw[pl] = mg.mw[pl]; p.mh[pl] = mg.mh[pl]; p.ms[pl] = mg.ms[pl]; p.mo[pl] = mg.mo[pl]; }
            p.cnt = b->intra_cellmap;
            p.lvl = (uint16_t *)(b->intra_cellmap + mg.cells);
            p.map_words = mg.cells * 3u / 4u;
            n_ops = std::max(n_ops, p.n_ops);
        }
        ia.status = c->status;
        if (total >= ((size_t)1 << 31)) return -22;
        if (ws_need(total) > c->rounds_ws_bytes) return -12;      // ensure_rounds_ws() comes first
        uint8_t *ws = (uint8_t *)c->rounds_ws;
        sa.ctl = (ExecCtl *)ws;
        sa.bins = (unsigned *)(ws + al256(sizeof(ExecCtl)));
        sa.thin = ws + al256(sizeof(ExecCtl)) + al256((size_t)MAX_LEVELS * N_BINS * 4);
        sa.lvl_end = (unsigned *)(sa.thin + al256(MAX_LEVELS));
        sa.lvl_left = sa.lvl_end + al256((size_t)MAX_LEVELS * 4) / 4;
        sa.key = (unsigned *)(ws + ws_hdr_bytes());
        sa.slots = (unsigned *)(ws + ws_hdr_bytes() + al256(total * 4));
        sa.n_slots_cap = (unsigned)std::min<size_t>(ws_slots_cap(total), 0xfffffff0u);
    }
    if (!fork_aux(c, st)) return -5;
    if ((mask & 16) && n_ops > 0) {
        cudaStream_t si = ss[NS - 1];
        D1_CHECK(cudaMemsetAsync(c->rounds_ws, 0, ws_hdr_bytes(), si));
        intra_clear_kernel<<<dim3(64u, (unsigned)n), 256, 0, si>>>(ia);
        intra_mark_kernel<<<dim3((unsigned)std::min((n_ops + 255) / 256, 64), (unsigned)n), 256, 0, si>>>(ia);
        int n_sched = 5;
        if (n_rec < n) { intra_levels_kernel<<<dim3((unsigned)LV_CL, (unsigned)n), LV_THREADS, 0, si>>>(sa); n_sched++; }   // a cluster per frame
        if (n_rec > 0) { intra_keys_kernel<<<dim3((unsigned)std::min((n_ops + 255) / 256, 64), (unsigned)n), 256, 0, si>>>(sa); n_sched++; }
        intra_sort_kernel<true><<<dim3((unsigned)((n_ops + 1023) / 1024), (unsigned)n), 256, 0, si>>>(sa);
        intra_scan_kernel<<<1, 1024, 0, si>>>(sa);
        intra_sort_kernel<false><<<dim3((unsigned)((n_ops + 1023) / 1024), (unsigned)n), 256, 0, si>>>(sa);
        count_launch(n_sched);
        if (!cuda_ok(cudaGetLastError(), "intra scheduling passes")) return -5;
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
        // predictions from references of another size: the two compound waves, then (below) the OBMC waves
        const Dav1dCudaMcScaledDesc *sc = (mask & 3) ? b->mc_scaled : nullptr;
        const int *nsc = b->n_mc_scaled;
        if (sc && (r = mc_scaled_launch_raw(dst, refs, sc, nsc[0], b->masks, s))) return r;
        if (sc && (r = mc_scaled_launch_raw(dst, refs, sc + nsc[0], nsc[1], b->masks, s))) return r;
        // OBMC blends onto the finished predictions: top neighbours, then left neighbours
        if ((mask & 1) && b->mc_obmc &&
            (r = mc_obmc_launch_raw(dst, refs, b->mc_obmc, b->mc_obmc_tiles, b->n_mc_obmc_tiles[0], s))) return r;
        if (sc && (r = mc_scaled_launch_raw(dst, refs, sc + nsc[0] + nsc[1], nsc[2], b->masks, s))) return r;
        if ((mask & 1) && b->mc_obmc &&
            (r = mc_obmc_launch_raw(dst, refs, b->mc_obmc, b->mc_obmc_tiles + b->n_mc_obmc_tiles[0],
                                    b->n_mc_obmc_tiles[1], s))) return r;
        if (sc && (r = mc_scaled_launch_raw(dst, refs, sc + nsc[0] + nsc[1] + nsc[2], nsc[3], b->masks, s))) return r;
        const CoefFmt fmt = { hbd && b->cf_int16, b->cf_esc, b->n_cf_esc };
        if ((mask & 8) && b->itx && b->itx_tasks) {
            if ((r = itx_task_launch(dst, nullptr, b->cf, b->itx, b->itx_tasks, b->n_itx_tasks[0], b->n_itx_tasks[1], 0, s, s,
                                     &fmt))) return r;
        } else if ((mask & 8) && b->itx && (r = itx_batch_launch(dst, nullptr, b->cf, b->itx, b->itx_class_count, 0, s, &fmt)))
            return r;
        // intra residual pre-pass -> int16 residual planes (independent of everything above)
        if ((mask & 16) && b->n_intra > 0 && b->intra_itx && b->intra_res) {
            const PicView rv = pic_view(b->intra_res);
            cudaStream_t s2 = ss[(f + NS / 2) % NS];
            if (b->intra_itx_tasks) {
                if ((r = itx_task_launch(dst, &rv, b->cf, b->intra_itx, b->intra_itx_tasks, b->n_intra_itx_tasks[0],
                                         b->n_intra_itx_tasks[1], 0, s2, s2, &fmt))) return r;
            } else if ((r = itx_batch_launch(dst, &rv, b->cf, b->intra_itx, b->intra_itx_class_count, 0, s2, &fmt))) return r;
        }
    }
    if (!join_aux(c, st)) return -5;
    if (!(mask & 16) || n_ops <= 0) return 0;
    // the executor: one persistent launch, as many blocks as the device holds at once (any number
    // would do: chunks are claimed by ticket)
    const size_t smem = hbd ? exec_smem_bytes<uint16_t>() : exec_smem_bytes<uint8_t>();
    const int grid = (int)std::min<size_t>((size_t)g_exec_blocks[hbd], (total + R_WARPS * 4 - 1) / (R_WARPS * 4));
    if (hbd) intra_exec_kernel<uint16_t><<<grid, R_WARPS * 32, smem, st>>>(sa);
    else intra_exec_kernel<uint8_t><<<grid, R_WARPS * 32, smem, st>>>(sa);
    count_launch();
    return cuda_ok(cudaGetLastError(), "intra_exec_kernel") ? 0 : -5;
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
    // a count byte and a 16-bit level per 4x4 cell of the three planes
    return ((size_t)map_geo(bw4, bh4, ss_hor, ss_ver).cells * 3 + 255) & ~(size_t)255;
}


// Recorder-side dependency levels: ONE pass over the descriptors in decode order (linear in the
// number of operations), the per-cell level map of the frame kept in host memory.  level + 1 goes to
// the descriptors' `reserved` field.  Returns the number of levels, or a negative errno.
int dav1d_cuda_intra_levels(Dav1dCudaIntraDesc *descs, int n, int bw4, int bh4, int ss_hor, int ss_ver) {
    if (!descs || n < 0 || bw4 <= 0 || bh4 <= 0) return -22;
    Intra2Frame f;
    memset(&f, 0, sizeof(f));
    f.bw4 = bw4; f.bh4 = bh4; f.pic.ss_hor = ss_hor; f.pic.ss_ver = ss_ver;
    const MapGeo mg = map_geo(bw4, bh4, ss_hor, ss_ver);
    for (int pl = 0; pl < 3; pl++) { f.mw[pl] = mg.mw[pl]; f.mh[pl] = mg.mh[pl]; f.ms[pl] = mg.ms[pl]; f.mo[pl] = mg.mo[pl]; }
    static thread_local std::vector<uint16_t> map;
    map.assign(mg.cells, 0);
    unsigned n_levels = 0;
    for (int i = 0; i < n; i++) {
        Dav1dCudaIntraDesc &d = descs[i];
        if (d.plane > 2 || d.x4 >= f.mw[d.plane] || d.y4 >= f.mh[d.plane]) return -22;
        int cls = 0;
        unsigned level = 0;
        op_deps(f, d, &cls, [&](const int pl, const int x0, const int x1, const int y0, const int y1, bool) {
            for (int y = y0; y < y1; y++) {
                const uint16_t *row = map.data() + f.mo[pl] + (size_t)y * f.ms[pl];
                for (int x = x0; x < x1; x++) level = std::max<unsigned>(level, row[x]);
            }
        });
        if (level >= 0xfffeu) return -34;
        d.reserved = (d.reserved & 0xffff0000u) | (level + 1u);
        n_levels = std::max(n_levels, level + 1u);
        const int pl = d.plane;
        for (int y = d.y4; y < imin(d.y4 + d.th4, f.mh[pl]); y++) {
            uint16_t *row = map.data() + f.mo[pl] + (size_t)y * f.ms[pl];
            for (int x = d.x4; x < imin(d.x4 + d.tw4, f.mw[pl]); x++) row[x] = (uint16_t)(level + 1u);
        }
    }
    return (int)n_levels;
}

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
