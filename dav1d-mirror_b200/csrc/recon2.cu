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
// launch per group of frames does the same at a finer grain:
//   * the recorder hands over the operations in DECODE order plus the offsets of the "units"
//     (superblocks) they belong to - nothing is scheduled, sorted or levelled on the host;
//   * a warp claims units in decode order from a counter (units of the group's frames
//     interleaved) and executes the unit's operations one after the other;
//   * a byte per 4x4 cell and plane counts the operations that still have to write the cell
//     (set up by a small marking launch, back at zero when the frame is done: the map needs no
//     clearing between frames).  Before an operation reads pixels OUTSIDE its own unit it waits
//     until their cells are at zero; every cell it waits for belongs to a unit that precedes it
//     in decode order, i.e. one that was claimed earlier by a warp that is running: no deadlock.
//   * prediction and residual of an operation are fused: the predictor writes a shared-memory
//     tile, the column pass of the inverse transform adds it and stores the final pixels once.
// A wait is bounded; a stalled dependency sets the context's status word, the remaining
// operations are abandoned, and dav1d_cuda_synchronize() reports the failure.
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
int itx_batch_launch(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs,
                     const int32_t *class_count, int zero_coefs, cudaStream_t st);
int itx_task_launch(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs, const uint32_t *tasks,
                    int n_small, int n_big, int zero_coefs, cudaStream_t st_small, cudaStream_t st_big);
int mc_obmc_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcDesc *descs,
                       const uint32_t *tiles, int n_tiles, cudaStream_t st);
void itx_init_attrs();
int mc_put_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcDesc *descs,
                      const uint32_t *tiles, int n_tiles, int n_small, uint8_t *masks, int16_t *tmp,
                      bool compound, cudaStream_t st);

constexpr int I2_WARPS = 4;
constexpr int EDGE_BUF = 288;
constexpr int EDGE_C = 144;
constexpr int I2_MAXF = DAV1D_CUDA_MAX_GROUP;
// bytes of the tile region: one 64-wide transform tile (32 x 65 ints), or - operations up to
// 32x32 - a 32 x 33 int transform tile, then the prediction tile, then the CfL ac / scratch tile
constexpr int I2_TILE_INTS = 32 * 65;
constexpr int I2_PRED_OFF = 32 * 33 * 4;                 // 4224
constexpr int I2_AC_OFF = I2_PRED_OFF + 32 * 32 * 2;     // 6272 (+ 2048 = 8320)

template <typename pixel> struct __align__(16) Intra2Smem {
    int tile[I2_TILE_INTS];
    pixel edge[EDGE_BUF];
    pixel scratch[IPRED_SCRATCH];
};

// one frame of the group
struct Intra2Frame {
    PicView pic;
    int bw4, bh4;
    void *cf;
    const Dav1dCudaIntraDesc *descs;     // decode order
    const void *pal;
    const uint8_t *pal_idx;
    const uint2 *units;                  // (first operation, count) in claim order
    int n_units, n_ops;
    uint8_t *map;                        // cell map: plane 0, 1, 2 one after the other
};
struct Intra2Args {
    Intra2Frame f[I2_MAXF];
    int nf, max_units;
    unsigned *claim;                     // claim counter of this launch (zeroed before)
    unsigned *status;                    // context status word: bit0 = a dependency wait timed out
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

DEV unsigned ld_acquire_u8(const uint8_t *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
DEV unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Wait until the cell (cx, cy) of plane pl is final.  Returns false when the wait was abandoned.
DEV bool wait_cell(const Intra2Frame &f, const int pl, const int cx, const int cy, unsigned *status) {
    const uint8_t *p = f.map + map_off(f, pl) + cy * map_w(f, pl) + cx;
    unsigned ns = 32, waited = 0;
    while (ld_acquire_u8(p) != 0) {
        if (ld_relaxed_u32(status) & 1u) return false;
        __nanosleep(ns);
        waited += ns;
        if (ns < 1024) ns <<= 1;
        if (waited > (1u << 28)) { atomicOr(status, 1u); return false; }      // ~0.27 s
    }
    return true;
}

// The cells whose pixels operation d reads and that lie outside d's own unit must be final.
// Exactly the pixels dav1d_prepare_intra_edges reads for the resolved mode (plus the source area
// of an intrabc block); cells of the operation's own unit were written by this warp.
DEV bool intra2_wait(const Intra2Frame &f, const Dav1dCudaIntraDesc &d, const int lane, unsigned *status) {
    const int mode = d.mode;
    if (mode == DAV1D_CUDA_INTRA_NONE || mode == DAV1D_CUDA_INTRA_PAL) return true;
    const int pl = d.plane;
    const int sh = pl ? f.pic.ss_hor : 0, sv = pl ? f.pic.ss_ver : 0;
    const int W = map_w(f, pl), H = map_h(f, pl);
    const int x0 = d.x4, y0 = d.y4;
    // cells known to be written by this warp: the operation's coding block (descriptor hint)
    int rx0 = x0, ry0 = y0, rx1 = x0, ry1 = y0;          // empty
    if (d.blk >> 16) {
        rx0 = x0 - (int)(d.blk & 15); ry0 = y0 - (int)((d.blk >> 4) & 15);
        rx1 = rx0 + (1 << ((d.blk >> 8) & 15)); ry1 = ry0 + (1 << ((d.blk >> 12) & 15));
    }
    bool ok = true;
    if (mode == DAV1D_CUDA_INTRA_IBC) {
        const int sx = (int16_t)(d.aux & 0xffff), sy = (int16_t)(d.aux >> 16);
        const int pw = 4 * W, ph = 4 * H;
        const int xa = iclip(sx, 0, pw - 1) >> 2, xb = iclip(sx + 4 * d.tw4 + (d.angle_delta ? 1 : 0) - 1, 0, pw - 1) >> 2;
        const int ya = iclip(sy, 0, ph - 1) >> 2, yb = iclip(sy + 4 * d.th4 + (d.flags ? 1 : 0) - 1, 0, ph - 1) >> 2;
        const int nx = xb - xa + 1, n = nx * (yb - ya + 1);
        for (int j = lane; j < n; j += 32) {
            const int cy = ya + j / nx, cx = xa + j % nx;
            if (cx >= rx0 && cx < rx1 && cy >= ry0 && cy < ry1) continue;
            ok &= wait_cell(f, pl, cx, cy, status);
        }
    } else {
        const int have_left = x0 > d.tile_x4_start, have_top = y0 > d.tile_y4_start;
        const int needs = mode == DAV1D_CUDA_INTRA_II ? intra_needs(d.angle_delta, 0, have_left, have_top)
                        : intra_needs(mode == DAV1D_CUDA_INTRA_CFL ? 0 : mode, d.angle_delta, have_left, have_top);
        // top row: [xs, xe) at y0 - 1
        int xs = 0, xe = 0;
        if (have_top && ((needs & 2) || (needs & 4) || ((needs & 1) && !have_left))) {
            const bool tr = (needs & 8) && (d.edge_flags & 1);
            xs = ((needs & 4) && have_left) ? x0 - 1 : x0;
            xe = (needs & 2) ? imin(x0 + d.tw4 + (tr ? d.tw4 : 0), d.tile_x4_end) : x0 + 1;
            xe = imin(xe, W);
        }
        // left column: [y0, ye) at x0 - 1
        int ye = y0;
        if (have_left && ((needs & 1) || ((needs & 2) && !have_top) || ((needs & 4) && !have_top))) {
            const bool bl = (needs & 16) && (d.edge_flags & 8);
            ye = (needs & 1) ? imin(y0 + d.th4 + (bl ? d.th4 : 0), d.tile_y4_end) : y0 + 1;
            ye = imin(ye, H);
        }
        const int nt = xe - xs, n = nt + (ye - y0);
        for (int j = lane; j < n; j += 32) {
            const int cx = j < nt ? xs + j : x0 - 1, cy = j < nt ? y0 - 1 : y0 + (j - nt);
            if (cx >= rx0 && cx < rx1 && cy >= ry0 && cy < ry1) continue;
            ok &= wait_cell(f, pl, cx, cy, status);
        }
        // CfL reads the co-located luma: the same block, i.e. the same unit (without any hint:
        // check the cells)
        if (mode == DAV1D_CUDA_INTRA_CFL && !(d.blk >> 16)) {
            const int lw4 = d.tw4 << sh, lh4 = d.th4 << sv;
            for (int j = lane; j < lw4 * lh4; j += 32) {
                const int cx = (x0 << sh) + j % lw4, cy = (y0 << sv) + j / lw4;
                if (cx < f.bw4 && cy < f.bh4) ok &= wait_cell(f, 0, cx, cy, status);
            }
        }
    }
    return __all_sync(0xffffffffu, ok);
}

// The operation's pixels are stored: one count less on each of its cells.
DEV void intra2_publish(const Intra2Frame &f, const Dav1dCudaIntraDesc &d, const int lane) {
    __syncwarp();
    __threadfence();
    const int pl = d.plane, W = map_w(f, pl), H = map_h(f, pl);
    uint8_t *m = f.map + map_off(f, pl);
    const int ltw = 31 - __clz((int)d.tw4);              // tw4 is a power of two
    const int n = d.th4 << ltw;
    for (int j = lane; j < n; j += 32) {
        const int cx = d.x4 + (j & (d.tw4 - 1)), cy = d.y4 + (j >> ltw);
        if (cx < W && cy < H) {
            uint8_t *p = m + cy * W + cx;
            *(volatile uint8_t *)p = (uint8_t)(*(volatile uint8_t *)p - 1);
        }
    }
}

// One intra-class operation (prediction [+ residual]) by one warp.
template <typename pixel>
__device__ __noinline__ void intra2_op(const Intra2Frame &a, const Dav1dCudaIntraDesc &d, Intra2Smem<pixel> *sm,
                                       const int lane) {
    typedef typename PxTraits<pixel>::coef coef;
    const int pl = d.plane;
    const int ss_hor = pl ? a.pic.ss_hor : 0, ss_ver = pl ? a.pic.ss_ver : 0;
    const PlaneView &pv = a.pic.p[pl];
    const int stride = (int)(pv.stride / (int)sizeof(pixel));
    pixel *dst = (pixel *)pv.data + (int64_t)d.y4 * 4 * stride + d.x4 * 4;
    const int w = d.tw4 * 4, h = d.th4 * 4;
    const int bdmax = a.pic.bdmax;
    pixel *edge = sm->edge + EDGE_C;
    const int have_left = d.x4 > d.tile_x4_start, have_top = d.y4 > d.tile_y4_start;
    const int mode = d.mode;
    const bool has_res = d.eob >= 0 && mode != DAV1D_CUDA_INTRA_PAL;
    pixel *ptile = (pixel *)((char *)sm->tile + I2_PRED_OFF);
    int16_t *ac = (int16_t *)((char *)sm->tile + I2_AC_OFF);
    // with a residual to follow, predictions of up to 32x32 go to the shared tile
    const bool to_tile = has_res && w <= 32 && h <= 32 &&
                         (mode <= DAV1D_CUDA_INTRA_FILTER || mode == DAV1D_CUDA_INTRA_CFL);
    pixel *pout = to_tile ? ptile : dst;
    const int pstride = to_tile ? w : stride;

    if (has_res) {
        // pull the block's coefficients towards the SM while the prediction runs
        const int ncoef = d.cw4 ? 16 * d.cw4 * d.ch4 : imin(w, 32) * imin(h, 32);
        const char *cp = (const char *)((const coef *)a.cf + d.coef_off);
        for (int o = lane * 128; o < ncoef * (int)sizeof(coef); o += 32 * 128)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(cp + o));
    }

    if (mode == DAV1D_CUDA_INTRA_PAL) {
        pal_pred_block<pixel>(dst, stride, (const pixel *)a.pal + d.aux, a.pal_idx + d.coef_off, w, h, lane, 32);
    } else if (mode == DAV1D_CUDA_INTRA_CFL) {
        const PlaneView &lv = a.pic.p[0];
        const int lstride = (int)(lv.stride / (int)sizeof(pixel));
        const pixel *luma = (const pixel *)lv.data + (int64_t)((d.y4 * 4) << ss_ver) * lstride + ((d.x4 * 4) << ss_hor);
        cfl_ac_block<pixel>(ac, luma, lstride, d.aux & 0xff, (d.aux >> 8) & 0xff, w, h, ss_hor, ss_ver, lane);
        int angle = 0;
        const int m = prepare_edges<pixel>(d.x4, have_left, d.y4, have_top, d.tile_x4_end, d.tile_y4_end, 0, dst,
                                           stride, nullptr, 0, &angle, d.tw4, d.th4, 0, edge, bdmax, lane);
        cfl_pred_block<pixel>(m, pout, pstride, edge, w, h, ac, d.angle_delta, bdmax, lane);
    } else if (mode == DAV1D_CUDA_INTRA_IBC) {
        // intrabc: put_bilin (mc_tmpl.c:395-450) from the current picture; coordinates clamped to
        // the 4*bw4 x 4*bh4 area (= emu_edge, recon_tmpl.c:974-995)
        const int sx = (int16_t)(d.aux & 0xffff), sy = (int16_t)(d.aux >> 16);
        const int mx = d.angle_delta, my = d.flags;
        const int pw = (4 * a.bw4) >> ss_hor, ph = (4 * a.bh4) >> ss_ver;
        const int ib = PxTraits<pixel>::inter_bits(bdmax);
        const pixel *base = (const pixel *)pv.data;
        const int lw = 31 - __clz(w);
        for (int i = lane; i < w * h; i += 32) {
            const int y = i >> lw, x = i & (w - 1);
            const int xa = iclip(sx + x, 0, pw - 1), xb = iclip(sx + x + 1, 0, pw - 1);
            const int ya = iclip(sy + y, 0, ph - 1), yb = iclip(sy + y + 1, 0, ph - 1);
            const int p00 = __ldcg(base + (int64_t)ya * stride + xa);
            int out;
            if (mx && my) {
                const int p01 = __ldcg(base + (int64_t)ya * stride + xb);
                const int p10 = __ldcg(base + (int64_t)yb * stride + xa);
                const int p11 = __ldcg(base + (int64_t)yb * stride + xb);
                const int sh1 = 4 - ib, r1 = (1 << sh1) >> 1;
                const int m0 = (16 * p00 + mx * (p01 - p00) + r1) >> sh1;
                const int m1 = (16 * p10 + mx * (p11 - p10) + r1) >> sh1;
                const int sh2 = 4 + ib;
                out = clip_px<pixel>((16 * m0 + my * (m1 - m0) + ((1 << sh2) >> 1)) >> sh2, bdmax);
            } else if (mx) {
                const int p01 = __ldcg(base + (int64_t)ya * stride + xb);
                const int sh1 = 4 - ib;
                const int px = (16 * p00 + mx * (p01 - p00) + ((1 << sh1) >> 1)) >> sh1;
                out = clip_px<pixel>((px + ((1 << ib) >> 1)) >> ib, bdmax);
            } else if (my) {
                const int p10 = __ldcg(base + (int64_t)yb * stride + xa);
                out = clip_px<pixel>((16 * p00 + my * (p10 - p00) + 8) >> 4, bdmax);
            } else {
                out = p00;
            }
            dst[y * stride + x] = (pixel)out;
        }
    } else if (mode == DAV1D_CUDA_INTRA_II) {
        // inter-intra: predict the whole block (<= 32x32) into scratch, then mc.blend onto the
        // inter prediction (mc_tmpl.c:642-653)
        int angle = 0;
        const int m = prepare_edges<pixel>(d.x4, have_left, d.y4, have_top, d.tile_x4_end, d.tile_y4_end, 0, dst,
                                           stride, nullptr, d.angle_delta, &angle, d.tw4, d.th4, 0, edge, bdmax, lane);
        pixel *tmp = (pixel *)ac;
        ipred_block<pixel>(m, tmp, w, edge, w, h, 0, 0, 0, bdmax, sm->scratch, lane);
        __syncwarp();
        const uint8_t *mask = a.pal_idx + d.coef_off;
        const int lw = 31 - __clz(w);
        for (int i = lane; i < w * h; i += 32) {
            const int y = i >> lw, x = i & (w - 1);
            const int mk = mask[i], p = dst[y * stride + x], q = tmp[i];
            dst[y * stride + x] = (pixel)((p * (64 - mk) + q * mk + 32) >> 6);
        }
    } else if (mode != DAV1D_CUDA_INTRA_NONE) {
        int angle = d.angle_delta;
        const int m = prepare_edges<pixel>(d.x4, have_left, d.y4, have_top, d.tile_x4_end, d.tile_y4_end,
                                           d.edge_flags, dst, stride, nullptr, mode, &angle, d.tw4, d.th4,
                                           (d.flags >> 10) & 1, edge, bdmax, lane);
        const int max_w = ((4 * a.bw4 + ss_hor) >> ss_hor) - 4 * d.x4;
        const int max_h = ((4 * a.bh4 + ss_ver) >> ss_ver) - 4 * d.y4;
        ipred_block<pixel>(m, pout, pstride, edge, w, h, angle | d.flags, max_w, max_h, bdmax, sm->scratch, lane);
    }
    __syncwarp();
    if (!has_res) return;
    itx2_block<pixel, 64>(true, lane, 32, sm->tile, (coef *)a.cf + d.coef_off, d.tx, d.txtp, d.eob, d.cw4, d.ch4,
                          to_tile ? ptile : dst, to_tile ? w : stride, dst, stride, bdmax, false);
}

// The persistent executor: see the header of this file.
template <typename pixel>
__global__ void __launch_bounds__(I2_WARPS * 32, 4) intra2_kernel(const __grid_constant__ Intra2Args a) {
    extern __shared__ __align__(16) uint8_t intra2_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Intra2Smem<pixel> *sm = (Intra2Smem<pixel> *)intra2_smem_raw + warp;
    const unsigned total = (unsigned)a.max_units * (unsigned)a.nf;
    for (;;) {
        unsigned k = 0;
        if (lane == 0) k = atomicAdd(a.claim, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= total) break;
        // claim k -> frame k % nf, unit k / nf: the frames of the group advance together
        const Intra2Frame &f = a.f[k % (unsigned)a.nf];
        const int u = (int)(k / (unsigned)a.nf);
        if (u >= f.n_units) continue;
        const uint2 un = f.units[u];
        const int i0 = (int)un.x, i1 = (int)(un.x + un.y);
        if (i0 >= i1) continue;
        Dav1dCudaIntraDesc d = f.descs[i0];
        for (int i = i0; i < i1; i++) {
            Dav1dCudaIntraDesc dn;
            if (i + 1 < i1) dn = f.descs[i + 1];             // next descriptor in flight during this operation
            if (!intra2_wait(f, d, lane, a.status)) return;  // abandoned: the host reports it
            intra2_op<pixel>(f, d, sm, lane);
            intra2_publish(f, d, lane);
            d = dn;
        }
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

static int g_i2_blocks[2] = { 0, 0 };   // resident blocks of the executor per pixel type

void recon_init_attrs() {
    itx_init_attrs();
    cudaFuncSetAttribute(intra2_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(I2_WARPS * sizeof(Intra2Smem<uint8_t>)));
    cudaFuncSetAttribute(intra2_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(I2_WARPS * sizeof(Intra2Smem<uint16_t>)));
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, intra2_kernel<uint8_t>, I2_WARPS * 32,
                                                  I2_WARPS * sizeof(Intra2Smem<uint8_t>));
    g_i2_blocks[0] = std::max(1, occ) * std::max(1, sms);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, intra2_kernel<uint16_t>, I2_WARPS * 32,
                                                  I2_WARPS * sizeof(Intra2Smem<uint16_t>));
    g_i2_blocks[1] = std::max(1, occ) * std::max(1, sms);
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
        if (b->n_intra > 0 && (!b->intra || !b->intra_units || b->n_intra_units < 1 || !b->intra_cellmap))
            return -22;
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
            p.descs = b->intra; p.pal = b->pal; p.pal_idx = b->pal_idx;
            p.units = (const uint2 *)b->intra_units;
            p.n_units = b->n_intra > 0 ? b->n_intra_units : 0;
            p.n_ops = b->n_intra > 0 ? b->n_intra : 0;
            p.map = b->intra_cellmap;
            ia.max_units = std::max(ia.max_units, p.n_units);
            n_ops = std::max(n_ops, p.n_ops);
        }
        ia.claim = c->claim + (c->claim_next++ % Dav1dCudaContext::N_CLAIM);
        ia.status = c->status;
    }
    if (!fork_aux(c, st)) return -5;
    if ((mask & 16) && n_ops > 0) {
        D1_CHECK(cudaMemsetAsync(ia.claim, 0, sizeof(unsigned), ss[NS - 1]));
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
            if ((r = itx_task_launch(dst, b->cf, b->itx, b->itx_tasks, b->n_itx_tasks[0], b->n_itx_tasks[1], 0, s, s)))
                return r;
        } else if ((mask & 8) && b->itx && (r = itx_batch_launch(dst, b->cf, b->itx, b->itx_class_count, 0, s))) return r;
    }
    if (!join_aux(c, st)) return -5;
    if (!(mask & 16) || n_ops <= 0) return 0;
    const size_t smem = I2_WARPS * (hbd ? sizeof(Intra2Smem<uint16_t>) : sizeof(Intra2Smem<uint8_t>));
    const long long claims = (long long)ia.max_units * n;
    const int grid = (int)std::min<long long>(g_i2_blocks[hbd], (claims + I2_WARPS - 1) / I2_WARPS);
    if (hbd) intra2_kernel<uint16_t><<<grid, I2_WARPS * 32, smem, st>>>(ia);
    else intra2_kernel<uint8_t><<<grid, I2_WARPS * 32, smem, st>>>(ia);
    count_launch();
    return cuda_ok(cudaGetLastError(), "intra2_kernel") ? 0 : -5;
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

// Host helper for recorders that do not track units themselves: the operations (decode order) are
// cut wherever the luma superblock (unit_log2 = 4: 64x64) changes; optionally wavefront order.
int dav1d_cuda_intra_units(const Dav1dCudaIntraDesc *descs, int n, int ss_hor, int ss_ver, int unit_log2,
                           int wave_gradient, uint32_t *units, int max_units)
{
    if (!descs || !units || n < 0 || unit_log2 < 1 || wave_gradient < 0) return -22;
    struct U { uint32_t first, count; int wave; };
    std::vector<U> us;
    int last_x = -1, last_y = -1;
    for (int i = 0; i < n; i++) {
        const Dav1dCudaIntraDesc &d = descs[i];
        const int ux = (d.plane ? d.x4 << ss_hor : d.x4) >> unit_log2, uy = (d.plane ? d.y4 << ss_ver : d.y4) >> unit_log2;
        if (us.empty() || ux != last_x || uy != last_y) {
            us.push_back({ (uint32_t)i, 0u, ux + wave_gradient * uy });
            last_x = ux; last_y = uy;
        }
        us.back().count++;
    }
    if ((int)us.size() > max_units) return -34;
    if (wave_gradient > 0)
        std::stable_sort(us.begin(), us.end(), [](const U &a, const U &b) { return a.wave < b.wave; });
    for (size_t k = 0; k < us.size(); k++) { units[2 * k] = us[k].first; units[2 * k + 1] = us[k].count; }
    return (int)us.size();
}

int dav1d_cuda_recon_submit(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b) {
    if (!c || !b) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    return group_submit_on(c, &b, 1, c->stream, 31);
}

int dav1d_cuda_recon_submit_phases(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b, int phase_mask) {
    if (!c || !b) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    return group_submit_on(c, &b, 1, c->stream, phase_mask & 31);
}

int dav1d_cuda_recon_group_submit(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n) {
    if (!c) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    return group_submit_on(c, bs, n, c->stream, 31);
}

int dav1d_cuda_recon_group_submit_phases(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n,
                                         int phase_mask)
{
    if (!c) return -22;
    D1_CHECK(cudaSetDevice(c->device));
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
