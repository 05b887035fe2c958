// Motion-compensation operator classes: batched kernels (put/prep, fused
// compound, warp, blend, scaled) and the Dav1dMCDSPContext overrides that run
// the same kernels on one staged block.  Reference: src/mc_tmpl.c.
#include <string.h>
#include <vector>
#include <algorithm>
#include <type_traits>
#include "ctx.h"
#include "stage.h"
#include "mc.cuh"

namespace d1 {

constexpr int MC_WARPS = 4;
#ifndef D1_PUT_MINB
#define D1_PUT_MINB 6
#endif
void mc_init_attrs();

struct McArgs {
    PicView dst;
    PicView refs[7];
    const Dav1dCudaMcDesc *descs;
    const uint32_t *tiles;      // desc_index * 16 + (ty * 4 + tx), 32x32 tiles
    int n_tiles;
    int n_small;                // leading tiles of at most 8x8 (grouped four per warp)
    uint8_t *masks;             // wedge / segmentation masks (device)
    int16_t *tmp;               // int16 pool for PREP output
    const void *tmaps[7];       // per reference: the picture's tensor maps (Dav1dCudaPicture.tma), or null
};

struct TileGeo { int x0, y0, tw, th; };
DEV TileGeo tile_geo(const Dav1dCudaMcDesc &d, const int t) {
    TileGeo g;
    g.x0 = (t & 3) * MC_T;
    g.y0 = (t >> 2) * MC_T;
    g.tw = imin(MC_T, d.w - g.x0);
    g.th = imin(MC_T, d.h - g.y0);
    return g;
}

// ---- put / prep: one warp per tile of up to 32x32, or (SMALL) four tiles of up to 8x8 per
// warp, one per group of 8 lanes (the groups may take different filter paths: every barrier
// inside mc_tile() is restricted to the group's lanes)
template <bool SMALL> struct McVar {
    static constexpr int G = SMALL ? 8 : 32, TMAX = SMALL ? 8 : 32, TPW = 32 / G;
};

// The grid is capped at the number of blocks that are resident at once (launch_mc) and every
// warp (group) strides over the tile list: warps are independent, so a slow tile (picture edge,
// large window) delays only its own warp instead of holding a block's slot, and the next
// tile's descriptor is fetched while the current tile is computed.
// WITH_PREP = false: the descriptor list holds no PREP entries (the frame path: a.tmp == nullptr) - the kernel then
// carries ONE instantiation of the filter code instead of two and fits the instruction caches (57 -> ~30 KB).
template <typename pixel, bool SMALL, bool WITH_PREP>
__global__ void __launch_bounds__(MC_WARPS * 32, D1_PUT_MINB) mc_put_kernel(const __grid_constant__ McArgs a) {
    extern __shared__ __align__(16) uint8_t mc_smem_raw[];
    typedef McVar<SMALL> V;
    const int wl = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = wl / V::G, lane = wl % V::G;
    const unsigned gmask = SMALL ? 0xffu << (8 * grp) : 0xffffffffu;
    const int stride = gridDim.x * MC_WARPS * V::TPW;
    int ti = (blockIdx.x * MC_WARPS + warp) * V::TPW + grp;
    if (ti >= a.n_tiles) return;
    McSmem<pixel, V::TMAX> *sm = (McSmem<pixel, V::TMAX> *)mc_smem_raw + (warp * V::TPW + grp);
    uint32_t tcode = a.tiles[ti];
    Dav1dCudaMcDesc d = a.descs[tcode >> 4];
    for (;;) {
        const int tn = ti + stride;
        uint32_t ncode = 0;
        Dav1dCudaMcDesc nd;
        if (tn < a.n_tiles) {
            ncode = a.tiles[tn];
            nd = a.descs[ncode >> 4];
        }
        const TileGeo g = tile_geo(d, tcode & 15);
        const Dav1dCudaMcSrc s = d.src[0];
        const PlaneView &ref = a.refs[s.ref].p[d.plane];
        if (WITH_PREP && d.kind == DAV1D_CUDA_MC_PREP) {
            int16_t *out = a.tmp + d.aux_off + g.y0 * d.w + g.x0;
            mc_tile<pixel, true, V::TMAX, V::G>(ref, s.x + g.x0, s.y + g.y0, g.tw, g.th, d.w, d.h, s.mx, s.my,
                                                s.filter_2d, a.dst.bdmax, sm, out, d.w, lane, gmask);
        } else {
            const PlaneView &dp = a.dst.p[d.plane];
            const int dstride = (int)(dp.stride / (int)sizeof(pixel));
            pixel *out = (pixel *)dp.data + (int64_t)(d.y + g.y0) * dstride + d.x + g.x0;
            mc_tile<pixel, false, V::TMAX, V::G>(ref, s.x + g.x0, s.y + g.y0, g.tw, g.th, d.w, d.h, s.mx, s.my,
                                                 s.filter_2d, a.dst.bdmax, sm, out, dstride, lane, gmask);
        }
        if (tn >= a.n_tiles) break;
        ti = tn; tcode = ncode; d = nd;
    }
}

// ---- fused compound: two preps into shared int16 tiles, then the combine
template <typename pixel, bool SMALL>
__global__ void __launch_bounds__(MC_WARPS * 32, 5) mc_compound_kernel(const __grid_constant__ McArgs a) {
    extern __shared__ __align__(16) uint8_t mc_smem_raw[];
    typedef McVar<SMALL> V;
    const int wl = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = wl / V::G, lane = wl % V::G;
    const unsigned gmask = SMALL ? 0xffu << (8 * grp) : 0xffffffffu;
    const int stride = gridDim.x * MC_WARPS * V::TPW;
    int ti = (blockIdx.x * MC_WARPS + warp) * V::TPW + grp;
    if (ti >= a.n_tiles) return;
    McSmemCompound<pixel, V::TMAX> *sm = (McSmemCompound<pixel, V::TMAX> *)mc_smem_raw + (warp * V::TPW + grp);
    uint32_t tcode = a.tiles[ti];
    Dav1dCudaMcDesc d = a.descs[tcode >> 4];
    for (;;) {
        const int tn = ti + stride;
        uint32_t ncode = 0;
        Dav1dCudaMcDesc nd;
        if (tn < a.n_tiles) {
            ncode = a.tiles[tn];
            nd = a.descs[ncode >> 4];
        }
        const TileGeo g = tile_geo(d, tcode & 15);
        for (int i = 0; i < 2; i++) {
            const Dav1dCudaMcSrc s = d.src[i];
            const PlaneView &ref = a.refs[s.ref].p[d.plane];
            mc_tile<pixel, true, V::TMAX, V::G>(ref, s.x + g.x0, s.y + g.y0, g.tw, g.th, d.w, d.h, s.mx, s.my,
                                                s.filter_2d, a.dst.bdmax, &sm->s, i ? sm->tb : sm->ta, V::TMAX, lane,
                                                gmask);
        }
        const PlaneView &dp = a.dst.p[d.plane];
        const int dstride = (int)(dp.stride / (int)sizeof(pixel));
        pixel *out = (pixel *)dp.data + (int64_t)(d.y + g.y0) * dstride + d.x + g.x0;
        uint8_t *mask = nullptr;
        int ms = 0;
        if (d.kind == DAV1D_CUDA_MC_MASK) {
            ms = d.w;
            mask = a.masks + d.aux_off + g.y0 * ms + g.x0;
        } else if (d.kind == DAV1D_CUDA_MC_W_MASK) {
            const int ssh = d.mask_ss >= 1, ssv = d.mask_ss == 2;
            ms = d.w >> ssh;
            mask = a.masks + d.aux_off + (g.y0 >> ssv) * ms + (g.x0 >> ssh);
        }
        mc_combine<pixel>(d.kind, sm->ta, sm->tb, V::TMAX, out, dstride, g.tw, g.th, d.weight, mask, ms, d.mask_ss,
                          a.dst.bdmax, lane, V::G);
        if (tn >= a.n_tiles) break;
        __syncwarp(gmask);      // the combine's reads of ta/tb end before the next tile's writes
        ti = tn; tcode = ncode; d = nd;
    }
}

// ---- the 32x32-tile kernels with TMA staging (references from dav1d_cuda_picture_alloc, which carry
// tensor maps): a warp walks its tiles as above; the window of the NEXT prediction is requested
// (one cp.async.bulk.tensor.2d by one lane) before the current one is filtered, the descriptor of the
// tile after that is fetched at the same time.  Windows at the picture border: mc_stage().
constexpr int MC_WARPS_TC = 5;            // warps per block of the compound variant (shared memory per SM / warp)
template <typename pixel>
DEV bool mc_tma_request(const McArgs &a, const Dav1dCudaMcDesc &d, const Dav1dCudaMcSrc &s, const TileGeo &g,
                        pixel *buf, unsigned long long *bar, const int lane)
{
    const PlaneView &ref = a.refs[s.ref].p[d.plane];
    const char *tm = (const char *)a.tmaps[s.ref];
    const int x = s.x + g.x0 - 3, y = s.y + g.y0 - 3;
    const bool ok = tm && x >= 0 && y >= 0 && x + g.tw + 7 <= ref.w && y + g.th + 7 <= ref.h;
    if (ok && lane == 0) {
        // the box starts at the 16-byte aligned pixel at or before the window's first column (a TMA
        // requirement on the innermost coordinate; the staged row is wide enough for the remainder)
        const int cls = mc_tma_class(g.th);
        mc_tma_issue(smem_u32(buf), tm + (d.plane * MC_TMA_CLASSES + cls) * 128, x & ~(McSrcGeo<pixel, MC_T>::VPX - 1), y,
                     (unsigned)(mc_tma_rows(cls) * McSrcGeo<pixel, MC_T>::STRIDE * (int)sizeof(pixel)), smem_u32(bar));
    }
    return ok;
}
// the window is in `buf` when this returns (off = its column offset)
template <typename pixel>
DEV int mc_tma_arrive(const McArgs &a, const Dav1dCudaMcDesc &d, const Dav1dCudaMcSrc &s, const TileGeo &g,
                      const bool requested, pixel *buf, unsigned long long *bar, const unsigned parity, const int lane)
{
    if (requested) {
        mbar_wait(smem_u32(bar), parity);
        return (s.x + g.x0 - 3) & (McSrcGeo<pixel, MC_T>::VPX - 1);
    }
    return mc_stage<pixel, MC_T, 32>(a.refs[s.ref].p[d.plane], s.x + g.x0, s.y + g.y0, g.tw, g.th, s.mx, s.my, buf, lane);
}

template <typename pixel, bool WITH_PREP>
__global__ void __launch_bounds__(MC_WARPS * 32, 5) mc_put_tma_kernel(const __grid_constant__ McArgs a) {
    extern __shared__ __align__(128) uint8_t mc_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = gridDim.x * MC_WARPS;
    int ti = blockIdx.x * MC_WARPS + warp;
    if (ti >= a.n_tiles) return;
    McSmemTma<pixel> *sm = (McSmemTma<pixel> *)mc_smem_raw + warp;
    mc_tma_init(sm, lane);
    uint32_t tcode = a.tiles[ti], ncode = 0;
    Dav1dCudaMcDesc d = a.descs[tcode >> 4], nd;
    TileGeo g = tile_geo(d, tcode & 15);
    if (ti + stride < a.n_tiles) { ncode = a.tiles[ti + stride]; nd = a.descs[ncode >> 4]; }
    unsigned par = 0;           // bit b: phase parity of buffer b's barrier
    int b = 0;
    bool cur = mc_tma_request<pixel>(a, d, d.src[0], g, sm->src[0], &sm->bar[0], lane);
    for (;;) {
        const int tn = ti + stride, tnn = tn + stride;
        uint32_t nncode = 0;
        Dav1dCudaMcDesc nnd;
        if (tnn < a.n_tiles) { nncode = a.tiles[tnn]; nnd = a.descs[nncode >> 4]; }      // used next time round
        TileGeo ng = g;
        bool nxt = false;
        if (tn < a.n_tiles) {
            ng = tile_geo(nd, ncode & 15);
            nxt = mc_tma_request<pixel>(a, nd, nd.src[0], ng, sm->src[b ^ 1], &sm->bar[b ^ 1], lane);
        }
        const Dav1dCudaMcSrc s = d.src[0];
        const int off = mc_tma_arrive<pixel>(a, d, s, g, cur, sm->src[b], &sm->bar[b], (par >> b) & 1u, lane);
        if (cur) par ^= 1u << b;
        if (WITH_PREP && d.kind == DAV1D_CUDA_MC_PREP) {
            int16_t *out = a.tmp + d.aux_off + g.y0 * d.w + g.x0;
            mc_filter<pixel, true, MC_T, 32>(sm->src[b], off, sm->mid, g.tw, g.th, d.w, d.h, s.mx, s.my, s.filter_2d,
                                             a.dst.bdmax, out, d.w, lane);
        } else {
            const PlaneView &dp = a.dst.p[d.plane];
            const int dstride = (int)(dp.stride / (int)sizeof(pixel));
            pixel *out = (pixel *)dp.data + (int64_t)(d.y + g.y0) * dstride + d.x + g.x0;
            mc_filter<pixel, false, MC_T, 32>(sm->src[b], off, sm->mid, g.tw, g.th, d.w, d.h, s.mx, s.my, s.filter_2d,
                                              a.dst.bdmax, out, dstride, lane);
        }
        if (tn >= a.n_tiles) break;
        ti = tn; tcode = ncode; d = nd; g = ng; cur = nxt; b ^= 1;
        ncode = nncode; nd = nnd;
    }
}

template <typename pixel>
__global__ void __launch_bounds__(MC_WARPS_TC * 32, 3) mc_compound_tma_kernel(const __grid_constant__ McArgs a) {
    extern __shared__ __align__(128) uint8_t mc_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = gridDim.x * MC_WARPS_TC;
    int ti = blockIdx.x * MC_WARPS_TC + warp;
    if (ti >= a.n_tiles) return;
    McSmemTmaCompound<pixel> *sm = (McSmemTmaCompound<pixel> *)mc_smem_raw + warp;
    mc_tma_init(&sm->s, lane);
    uint32_t tcode = a.tiles[ti], ncode = 0;
    Dav1dCudaMcDesc d = a.descs[tcode >> 4], nd;
    TileGeo g = tile_geo(d, tcode & 15);
    if (ti + stride < a.n_tiles) { ncode = a.tiles[ti + stride]; nd = a.descs[ncode >> 4]; }
    // the first source's window always lands in buffer 0, the second's in buffer 1; bit i of `par` / `req`:
    // phase parity of buffer i's barrier / its window was requested through TMA
    unsigned par = 0;
    unsigned req = mc_tma_request<pixel>(a, d, d.src[0], g, sm->s.src[0], &sm->s.bar[0], lane) ? 1u : 0u;
    for (;;) {
        const int tn = ti + stride, tnn = tn + stride;
        uint32_t nncode = 0;
        Dav1dCudaMcDesc nnd;
        if (tnn < a.n_tiles) { nncode = a.tiles[tnn]; nnd = a.descs[nncode >> 4]; }
        TileGeo ng = g;
        if (tn < a.n_tiles) ng = tile_geo(nd, ncode & 15);
        // the two predictions through ONE copy of the filter code (the loop stays rolled: two copies do not fit the
        // instruction caches, measured 51 instead of 40 us per frame): source i from window buffer i; before it is
        // filtered, the request for the next window goes out - (d, 1) while i == 0, (next tile, 0) while i == 1
#pragma unroll 1
        for (int i = 0; i < 2; i++) {
            const int j = i ^ 1;
            if (i == 0 || tn < a.n_tiles) {
                const Dav1dCudaMcDesc &pd = i ? nd : d;
                const bool ok = mc_tma_request<pixel>(a, pd, i ? pd.src[0] : pd.src[1], i ? ng : g, sm->s.src[j], &sm->s.bar[j], lane);
                req = (req & ~(1u << j)) | ((ok ? 1u : 0u) << j);
            }
            const Dav1dCudaMcSrc s = i ? d.src[1] : d.src[0];
            const bool requested = (req >> i) & 1u;
            const int off = mc_tma_arrive<pixel>(a, d, s, g, requested, sm->s.src[i], &sm->s.bar[i], (par >> i) & 1u, lane);
            if (requested) par ^= 1u << i;
            mc_filter<pixel, true, MC_T, 32>(sm->s.src[i], off, sm->s.mid, g.tw, g.th, d.w, d.h, s.mx, s.my, s.filter_2d,
                                             a.dst.bdmax, i ? sm->tb : sm->ta, MC_T, lane);
        }
        const PlaneView &dp = a.dst.p[d.plane];
        const int dstride = (int)(dp.stride / (int)sizeof(pixel));
        pixel *out = (pixel *)dp.data + (int64_t)(d.y + g.y0) * dstride + d.x + g.x0;
        uint8_t *mask = nullptr;
        int ms = 0;
        if (d.kind == DAV1D_CUDA_MC_MASK) {
            ms = d.w;
            mask = a.masks + d.aux_off + g.y0 * ms + g.x0;
        } else if (d.kind == DAV1D_CUDA_MC_W_MASK) {
            const int ssh = d.mask_ss >= 1, ssv = d.mask_ss == 2;
            ms = d.w >> ssh;
            mask = a.masks + d.aux_off + (g.y0 >> ssv) * ms + (g.x0 >> ssh);
        }
        mc_combine<pixel>(d.kind, sm->ta, sm->tb, MC_T, out, dstride, g.tw, g.th, d.weight, mask, ms, d.mask_ss,
                          a.dst.bdmax, lane, 32);
        if (tn >= a.n_tiles) break;
        __syncwarp();           // the combine's reads of ta/tb end before the next tile's writes
        ti = tn; tcode = ncode; d = nd; g = ng;
        ncode = nncode; nd = nnd;
    }
}

// ---- OBMC (obmc(), recon_tmpl.c:1071-1131): the neighbour's prediction of a tile into a shared
// pixel tile ("lap"), then blend_h (top neighbour: rows < 3/4 of the blend height, mask
// obmc_masks[bh + y]) or blend_v (left neighbour: columns < 3/4 of the width, obmc_masks[w + x])
// onto the block's own prediction (mc_tmpl.c:655-681).
template <typename pixel> struct __align__(16) McSmemObmc {
    McSmem<pixel, 32> s;
    pixel lap[32 * 32];
};

template <typename pixel>
__global__ void __launch_bounds__(MC_WARPS * 32) mc_obmc_kernel(const __grid_constant__ McArgs a) {
    extern __shared__ __align__(16) uint8_t mc_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ti = blockIdx.x * MC_WARPS + warp;
    if (ti >= a.n_tiles) return;
    McSmemObmc<pixel> *sm = (McSmemObmc<pixel> *)mc_smem_raw + warp;
    const uint32_t tcode = a.tiles[ti];
    const Dav1dCudaMcDesc d = a.descs[tcode >> 4];
    const TileGeo g = tile_geo(d, tcode & 15);
    const Dav1dCudaMcSrc s = d.src[0];
    const PlaneView &ref = a.refs[s.ref].p[d.plane];
    mc_tile<pixel, false, 32, 32>(ref, s.x + g.x0, s.y + g.y0, g.tw, g.th, d.w, d.h, s.mx, s.my, s.filter_2d,
                                  a.dst.bdmax, &sm->s, sm->lap, 32, lane);
    const PlaneView &dp = a.dst.p[d.plane];
    const int dstride = (int)(dp.stride / (int)sizeof(pixel));
    pixel *dst = (pixel *)dp.data + (int64_t)(d.y + g.y0) * dstride + d.x + g.x0;
    const bool horz = d.kind == DAV1D_CUDA_MC_OBMC_H;
    const int bh = horz ? d.aux16 : d.h;
    // region of the whole block that is blended, clipped to this tile
    const int lim_x = horz ? d.w : (d.w * 3) >> 2, lim_y = horz ? (bh * 3) >> 2 : d.h;
    const int nx = imin(g.tw, lim_x - g.x0), ny = imin(g.th, lim_y - g.y0);
    if (nx <= 0 || ny <= 0) return;
    for (int i = lane; i < ny * 32; i += 32) {
        const int y = i >> 5, x = i & 31;
        if (x >= nx) continue;
        const int m = horz ? g_obmc_masks[bh + g.y0 + y] : g_obmc_masks[d.w + g.x0 + x];
        const int p = dst[y * dstride + x], q = sm->lap[y * 32 + x];
        dst[y * dstride + x] = (pixel)((p * (64 - m) + q * m + 32) >> 6);
    }
}

// ---- stand-alone ops on one block (per-call surface + unfused batch use)
struct BlockOp {
    int kind;                 // combine: Dav1dCudaMcKind; blend: 0/1/2
    int w, h;
    int weight, mask_ss;
    void *dst; int dstride;   // pixels
    const void *a, *b;        // combine: int16 tmp1/tmp2 (stride w); blend: a = pixel tmp (stride w)
    uint8_t *mask;
    int bdmax;
};

template <typename pixel> __global__ void mc_combine_kernel(const BlockOp op) {
    const int ssh = op.mask_ss >= 1;
    const int ms = op.kind == DAV1D_CUDA_MC_W_MASK ? op.w >> ssh : op.w;
    mc_combine<pixel>(op.kind, (const int16_t *)op.a, (const int16_t *)op.b, op.w, (pixel *)op.dst, op.dstride,
                      op.w, op.h, op.weight, op.mask, ms, op.mask_ss, op.bdmax, threadIdx.x, blockDim.x);
}

template <typename pixel> __global__ void mc_blend_kernel(const BlockOp op) {
    mc_blend<pixel>(op.kind, (pixel *)op.dst, op.dstride, (const pixel *)op.a, op.w, op.h, op.mask,
                    threadIdx.x, blockDim.x);
}

struct WarpArgs {
    PlaneView ref;
    int sx, sy;
    int16_t abcd[4];
    int mx, my;
    void *out; int ostride;
    int bdmax;
};
template <typename pixel, bool PREP> __global__ void mc_warp_kernel(const WarpArgs w) {
    __shared__ int16_t mid[15 * 8];
    mc_warp8x8<pixel, PREP>(w.ref, w.sx, w.sy, w.abcd, w.mx, w.my, w.bdmax, mid,
                            (typename McOut<pixel, PREP>::type *)w.out, w.ostride, threadIdx.x);
}

struct EmuArgs { PlaneView ref; int x, y, bw, bh; void *dst; int dstride; };
template <typename pixel> __global__ void mc_emu_edge_kernel(const EmuArgs e) {
    const pixel *rp = (const pixel *)e.ref.data;
    const int64_t rs = e.ref.stride / (int64_t)sizeof(pixel);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < e.bw * e.bh; i += gridDim.x * blockDim.x) {
        const int y = i / e.bw, x = i % e.bw;
        ((pixel *)e.dst)[y * e.dstride + x] =
            rp[iclip(e.y + y, 0, e.ref.h - 1) * rs + iclip(e.x + x, 0, e.ref.w - 1)];
    }
}

// ---- scaled put/prep (reference scaling), one warp per 32x32 tile.
// mc_tmpl.c:173-221 (put_8tap_scaled), :284-328 (prep), :452-491,548-585 (bilin)
struct ScaledArgs {
    PlaneView ref;
    int sx, sy;               // integer position of the block's top-left
    int w, h, mx, my, dx, dy; // mx,my in 1/1024
    int filter_2d;
    void *out; int ostride;
    int bdmax;
};

constexpr int SC_MAX_ROWS = 72;

// One tile (at most 32x32, top-left (x0, y0) of a w x h block) of a scaled prediction, by one warp.
// (sx, sy): integer position of the block's top-left in `ref`, read with clamping; mx, my: the
// 1/1024 remainder; dx, dy: steps.  `mid` holds SC_MAX_ROWS * MC_T int16.  out: the tile's first
// element, stride ostride.
template <typename pixel, bool PREP, typename out_t>
DEV void mc_scaled_tile(const PlaneView &ref, const int sx, const int sy, const int w, const int h, const int mx,
                        const int my, const int dx, const int dy, const int filter_2d, const int bdmax, const int x0,
                        const int y0, const int tw, const int th, int16_t *mid, out_t *out, const int ostride,
                        const int lane)
{
    const bool bilin = filter_2d == 9;
    const int ib = PxTraits<pixel>::inter_bits(bdmax);
    const int th_t = (0x15A80 >> (2 * filter_2d)) & 3, tv_t = filter_2d % 3;
    const int hset = w > 4 ? th_t : 3 + (th_t & 1);
    const int vset = h > 4 ? tv_t : 3 + (tv_t & 1);
    const pixel *rp = (const pixel *)ref.data;
    const int64_t rs = ref.stride / (int64_t)sizeof(pixel);

    // rows of `mid` this tile needs, relative to the block's first mid row
    const int ypos0 = my + y0 * dy;
    const int row_first = ypos0 >> 10;
    const int row_last = ((my + (y0 + th - 1) * dy) >> 10) + (bilin ? 1 : 7);
    const int nrows = row_last - row_first + 1;
    const int roff = bilin ? 0 : -3;      // mid row r <-> source row sy + roff + r
    for (int i = lane; i < nrows * tw; i += 32) {
        const int r = i / tw, x = i % tw;
        const int pos = mx + (x0 + x) * dx;
        const int ioff = pos >> 10, fx = (pos & 0x3ff) >> 6;
        const int yy = iclip(sy + roff + row_first + r, 0, ref.h - 1);
        const pixel *row = rp + yy * rs;
        int v;
        if (bilin) {
            const int p0 = row[iclip(sx + ioff, 0, ref.w - 1)];
            const int p1 = row[iclip(sx + ioff + 1, 0, ref.w - 1)];
            const int sh = 4 - ib;
            v = (16 * p0 + fx * (p1 - p0) + ((1 << sh) >> 1)) >> sh;
        } else if (fx) {
            const int8_t *f = g_subpel_filters + (hset * 15 + fx - 1) * 8;
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += f[k] * row[iclip(sx + ioff + k - 3, 0, ref.w - 1)];
            const int sh = 6 - ib;
            v = (sum + ((1 << sh) >> 1)) >> sh;
        } else {
            v = row[iclip(sx + ioff, 0, ref.w - 1)] << ib;
        }
        mid[r * MC_T + x] = (int16_t)v;
    }
    __syncwarp();
    for (int i = lane; i < tw * th; i += 32) {
        const int y = i / tw, x = i % tw;
        const int ypos = my + (y0 + y) * dy;
        const int r = (ypos >> 10) - row_first, fy = (ypos & 0x3ff) >> 6;
        const int16_t *m = mid + r * MC_T + x;
        int res;
        if (bilin) {
            const int s = 16 * m[0] + fy * (m[MC_T] - m[0]);
            if (PREP) res = ((s + 8) >> 4) - PxTraits<pixel>::prep_bias;
            else res = clip_px<pixel>((s + ((1 << (4 + ib)) >> 1)) >> (4 + ib), bdmax);
        } else if (fy) {
            const int8_t *f = g_subpel_filters + (vset * 15 + fy - 1) * 8;
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += f[k] * m[k * MC_T];
            if (PREP) res = ((sum + 32) >> 6) - PxTraits<pixel>::prep_bias;
            else res = clip_px<pixel>((sum + ((1 << (6 + ib)) >> 1)) >> (6 + ib), bdmax);
        } else {
            const int c = m[3 * MC_T];
            if (PREP) res = c - PxTraits<pixel>::prep_bias;
            else res = clip_px<pixel>((c + ((1 << ib) >> 1)) >> ib, bdmax);
        }
        out[y * ostride + x] = (out_t)res;
    }
    __syncwarp();                           // `mid` may be rewritten by the caller's next tile
}

template <typename pixel, bool PREP>
__global__ void __launch_bounds__(MC_WARPS * 32) mc_scaled_kernel(const ScaledArgs a) {
    __shared__ int16_t mid_all[MC_WARPS][SC_MAX_ROWS * MC_T];
    typedef typename McOut<pixel, PREP>::type out_t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_x = (a.w + MC_T - 1) / MC_T, tiles_y = (a.h + MC_T - 1) / MC_T;
    const int t = blockIdx.x * MC_WARPS + warp;
    if (t >= tiles_x * tiles_y) return;
    const int x0 = (t % tiles_x) * MC_T, y0 = (t / tiles_x) * MC_T;
    const int tw = imin(MC_T, a.w - x0), th = imin(MC_T, a.h - y0);
    mc_scaled_tile<pixel, PREP, out_t>(a.ref, a.sx, a.sy, a.w, a.h, a.mx, a.my, a.dx, a.dy, a.filter_2d, a.bdmax, x0, y0,
                                       tw, th, mid_all[warp], (out_t *)a.out + y0 * a.ostride + x0, a.ostride, lane);
}

// ---- batched form (Dav1dCudaMcScaledDesc): one warp per descriptor, tile after tile; one or two
// scaled predictions into shared tiles, then the same combine / blend code as the same-size path
struct ScaledBatchArgs {
    PicView dst;
    PicView refs[7];
    const Dav1dCudaMcScaledDesc *descs;
    int n;
    uint8_t *masks;
};

template <typename pixel> struct __align__(16) ScaledSmem {
    int16_t mid[SC_MAX_ROWS * MC_T];
    int16_t ta[MC_T * MC_T], tb[MC_T * MC_T];
};

template <typename pixel>
__global__ void __launch_bounds__(MC_WARPS * 32) mc_scaled_batch_kernel(const __grid_constant__ ScaledBatchArgs a) {
    __shared__ ScaledSmem<pixel> sm_all[MC_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int di = blockIdx.x * MC_WARPS + warp;
    if (di >= a.n) return;
    ScaledSmem<pixel> *sm = &sm_all[warp];
    const Dav1dCudaMcScaledDesc d = a.descs[di];
    const PlaneView &dp = a.dst.p[d.plane];
    const int dstride = (int)(dp.stride / (int)sizeof(pixel));
    const bool obmc = d.kind == DAV1D_CUDA_MC_OBMC_H || d.kind == DAV1D_CUDA_MC_OBMC_V;
    const bool single = d.kind == DAV1D_CUDA_MC_PUT || obmc;
    for (int y0 = 0; y0 < d.h; y0 += MC_T) {
        for (int x0 = 0; x0 < d.w; x0 += MC_T) {
            const int tw = imin(MC_T, d.w - x0), th = imin(MC_T, d.h - y0);
            pixel *out = (pixel *)dp.data + (int64_t)(d.y + y0) * dstride + d.x + x0;
            if (single) {
                const Dav1dCudaMcScaledSrc s = d.src[0];
                const PlaneView &ref = a.refs[s.ref].p[d.plane];
                pixel *lap = (pixel *)sm->ta;
                if (obmc)
                    mc_scaled_tile<pixel, false, pixel>(ref, s.pos_x >> 10, s.pos_y >> 10, d.w, d.h, s.pos_x & 0x3ff,
                                                        s.pos_y & 0x3ff, s.step_x, s.step_y, s.filter_2d, a.dst.bdmax, x0,
                                                        y0, tw, th, sm->mid, lap, MC_T, lane);
                else
                    mc_scaled_tile<pixel, false, pixel>(ref, s.pos_x >> 10, s.pos_y >> 10, d.w, d.h, s.pos_x & 0x3ff,
                                                        s.pos_y & 0x3ff, s.step_x, s.step_y, s.filter_2d, a.dst.bdmax, x0,
                                                        y0, tw, th, sm->mid, out, dstride, lane);
                if (obmc) {
                    const bool horz = d.kind == DAV1D_CUDA_MC_OBMC_H;
                    const int bh = horz ? d.aux16 : d.h;
                    const int lim_x = horz ? d.w : (d.w * 3) >> 2, lim_y = horz ? (bh * 3) >> 2 : d.h;
                    const int nx = imin(tw, lim_x - x0), ny = imin(th, lim_y - y0);
                    for (int i = lane; i < ny * 32; i += 32) {
                        const int y = i >> 5, x = i & 31;
                        if (x >= nx) continue;
                        const int m = horz ? g_obmc_masks[bh + y0 + y] : g_obmc_masks[d.w + x0 + x];
                        const int p = out[y * dstride + x], q = lap[y * MC_T + x];
                        out[y * dstride + x] = (pixel)((p * (64 - m) + q * m + 32) >> 6);
                    }
                    __syncwarp();
                }
                continue;
            }
            for (int i = 0; i < 2; i++) {
                const Dav1dCudaMcScaledSrc s = d.src[i];
                const PlaneView &ref = a.refs[s.ref].p[d.plane];
                mc_scaled_tile<pixel, true, int16_t>(ref, s.pos_x >> 10, s.pos_y >> 10, d.w, d.h, s.pos_x & 0x3ff,
                                                     s.pos_y & 0x3ff, s.step_x, s.step_y, s.filter_2d, a.dst.bdmax, x0, y0,
                                                     tw, th, sm->mid, i ? sm->tb : sm->ta, MC_T, lane);
            }
            uint8_t *mask = nullptr;
            int ms = 0;
            if (d.kind == DAV1D_CUDA_MC_MASK) {
                ms = d.w;
                mask = a.masks + d.aux_off + y0 * ms + x0;
            } else if (d.kind == DAV1D_CUDA_MC_W_MASK) {
                const int ssh = d.mask_ss >= 1, ssv = d.mask_ss == 2;
                ms = d.w >> ssh;
                mask = a.masks + d.aux_off + (y0 >> ssv) * ms + (x0 >> ssh);
            }
            mc_combine<pixel>(d.kind, sm->ta, sm->tb, MC_T, out, dstride, tw, th, d.weight, mask, ms, d.mask_ss,
                              a.dst.bdmax, lane, 32);
            __syncwarp();
        }
    }
}

int mc_scaled_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcScaledDesc *descs, const int n,
                         uint8_t *masks, cudaStream_t s)
{
    if (n <= 0) return 0;
    ScaledBatchArgs a;
    a.dst = dst;
    for (int i = 0; i < 7; i++) a.refs[i] = refs[i];
    a.descs = descs; a.n = n; a.masks = masks;
    const int grid = (n + MC_WARPS - 1) / MC_WARPS;
    if (dst.bdmax > 0xff) mc_scaled_batch_kernel<uint16_t><<<grid, MC_WARPS * 32, 0, s>>>(a);
    else mc_scaled_batch_kernel<uint8_t><<<grid, MC_WARPS * 32, 0, s>>>(a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "mc_scaled_batch_kernel") ? 0 : -5;
}

// ------------------------------------------------------------------ launchers
static std::vector<uint32_t> tiles_for(int desc_idx, int w, int h) {
    std::vector<uint32_t> t;
    for (int ty = 0; ty * MC_T < h; ty++)
        for (int tx = 0; tx * MC_T < w; tx++)
            t.push_back((uint32_t)desc_idx * 16 + ty * 4 + tx);
    return t;
}

// a.tiles[0 .. a.n_small) are tiles of at most 8x8 (four per warp), the rest one per warp
// resident blocks per SM of a kernel variant x number of SMs (cached per variant)
template <typename K> static int resident_blocks(K kernel, size_t smem, int &cache, int threads = MC_WARPS * 32) {
    if (cache <= 0) {
        int dev = 0, sms = 148, per_sm = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        cache = sms * per_sm;
    }
    return cache;
}

// TMA staging of the 32x32-tile kernels where the references carry tensor maps (dav1d_cuda_set_mc_tma):
// 0 off, 1 single-reference predictions (default), 2 compound predictions as well.  Measured on B200 (4K
// 10-bit benchmark mix, per frame): put 20.6 us with TMA vs 21.6 with cp.async; compound 42.3 vs 41.0 us - the
// second window buffer costs a quarter of the resident warps (15 instead of 20 per SM), so compound stays on
// cp.async by default.  (A first version that called the filter code twice per tile ran at 54.5 us: two copies
// of it, 65 KB, do not fit the instruction caches - both compound kernels keep the two predictions in a rolled
// loop, and the put kernels are instantiated without the PREP branch for the frame path.)
static int g_mc_tma = 1;
void mc_set_tma(int mode) { g_mc_tma = mode < 0 ? 0 : mode > 2 ? 2 : mode; }
int mc_get_tma() { return g_mc_tma; }

template <typename pixel, bool COMPOUND>
static int launch_mc(McArgs a, cudaStream_t st) {
    static int cap_small = 0, cap_big = 0, cap_small_prep = 0, cap_big_prep = 0;
    const int n_small = a.n_small, n_big = a.n_tiles - a.n_small;
    const uint32_t *tiles = a.tiles;
    if (n_small > 0) {
        a.tiles = tiles; a.n_tiles = n_small;
        int grid = (n_small + MC_WARPS * 4 - 1) / (MC_WARPS * 4);
        if (COMPOUND) {
            const size_t smem = MC_WARPS * 4 * sizeof(McSmemCompound<pixel, 8>);
            grid = std::min(grid, resident_blocks(mc_compound_kernel<pixel, true>, smem, cap_small));
            mc_compound_kernel<pixel, true><<<grid, MC_WARPS * 32, smem, st>>>(a);
        } else {
            const size_t smem = MC_WARPS * 4 * sizeof(McSmem<pixel, 8>);
            if (a.tmp) {
                grid = std::min(grid, resident_blocks(mc_put_kernel<pixel, true, true>, smem, cap_small_prep));
                mc_put_kernel<pixel, true, true><<<grid, MC_WARPS * 32, smem, st>>>(a);
            } else {
                grid = std::min(grid, resident_blocks(mc_put_kernel<pixel, true, false>, smem, cap_small));
                mc_put_kernel<pixel, true, false><<<grid, MC_WARPS * 32, smem, st>>>(a);
            }
        }
        count_launch();
    }
    bool tma = g_mc_tma >= (COMPOUND ? 2 : 1);
    for (int i = 0; i < 7; i++) tma = tma && (a.tmaps[i] || !a.refs[i].p[0].data);
    if (n_big > 0 && tma) {
        // every reference carries tensor maps: TMA staging, two window buffers per warp
        static int cap_tma = 0, cap_tma_prep = 0;
        a.tiles = tiles + n_small; a.n_tiles = n_big;
        if (COMPOUND) {
            const size_t smem = MC_WARPS_TC * sizeof(McSmemTmaCompound<pixel>);
            const int grid = std::min((n_big + MC_WARPS_TC - 1) / MC_WARPS_TC,
                                      resident_blocks(mc_compound_tma_kernel<pixel>, smem, cap_tma, MC_WARPS_TC * 32));
            mc_compound_tma_kernel<pixel><<<grid, MC_WARPS_TC * 32, smem, st>>>(a);
        } else {
            const size_t smem = MC_WARPS * sizeof(McSmemTma<pixel>);
            if (a.tmp) {
                const int grid = std::min((n_big + MC_WARPS - 1) / MC_WARPS,
                                          resident_blocks(mc_put_tma_kernel<pixel, true>, smem, cap_tma_prep, MC_WARPS * 32));
                mc_put_tma_kernel<pixel, true><<<grid, MC_WARPS * 32, smem, st>>>(a);
            } else {
                const int grid = std::min((n_big + MC_WARPS - 1) / MC_WARPS,
                                          resident_blocks(mc_put_tma_kernel<pixel, false>, smem, cap_tma, MC_WARPS * 32));
                mc_put_tma_kernel<pixel, false><<<grid, MC_WARPS * 32, smem, st>>>(a);
            }
        }
        count_launch();
    } else if (n_big > 0) {
        a.tiles = tiles + n_small; a.n_tiles = n_big;
        int grid = (n_big + MC_WARPS - 1) / MC_WARPS;
        if (COMPOUND) {
            const size_t smem = MC_WARPS * sizeof(McSmemCompound<pixel, 32>);
            grid = std::min(grid, resident_blocks(mc_compound_kernel<pixel, false>, smem, cap_big));
            mc_compound_kernel<pixel, false><<<grid, MC_WARPS * 32, smem, st>>>(a);
        } else {
            const size_t smem = MC_WARPS * sizeof(McSmem<pixel, 32>);
            if (a.tmp) {
                grid = std::min(grid, resident_blocks(mc_put_kernel<pixel, false, true>, smem, cap_big_prep));
                mc_put_kernel<pixel, false, true><<<grid, MC_WARPS * 32, smem, st>>>(a);
            } else {
                grid = std::min(grid, resident_blocks(mc_put_kernel<pixel, false, false>, smem, cap_big));
                mc_put_kernel<pixel, false, false><<<grid, MC_WARPS * 32, smem, st>>>(a);
            }
        }
        count_launch();
    }
    return cuda_ok(cudaGetLastError(), COMPOUND ? "mc_compound_kernel" : "mc_put_kernel") ? 0 : -5;
}

int mc_obmc_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcDesc *descs,
                       const uint32_t *tiles, int n_tiles, cudaStream_t st)
{
    if (n_tiles <= 0 || !descs || !tiles) return 0;
    McArgs a;
    memset(&a, 0, sizeof(a));
    a.dst = dst;
    for (int i = 0; i < 7; i++) a.refs[i] = refs[i];
    a.descs = descs; a.tiles = tiles; a.n_tiles = n_tiles;
    const int grid = (n_tiles + MC_WARPS - 1) / MC_WARPS;
    if (dst.bdmax > 0xff) mc_obmc_kernel<uint16_t><<<grid, MC_WARPS * 32, MC_WARPS * sizeof(McSmemObmc<uint16_t>), st>>>(a);
    else mc_obmc_kernel<uint8_t><<<grid, MC_WARPS * 32, MC_WARPS * sizeof(McSmemObmc<uint8_t>), st>>>(a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "mc_obmc_kernel") ? 0 : -5;
}

int mc_put_launch(const McArgs &a, cudaStream_t st) {
    if (a.n_tiles <= 0) return 0;
    return a.dst.bdmax > 0xff ? launch_mc<uint16_t, false>(a, st) : launch_mc<uint8_t, false>(a, st);
}
int mc_compound_launch(const McArgs &a, cudaStream_t st) {
    if (a.n_tiles <= 0) return 0;
    return a.dst.bdmax > 0xff ? launch_mc<uint16_t, true>(a, st) : launch_mc<uint8_t, true>(a, st);
}

int mc_put_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcDesc *descs,
                      const uint32_t *tiles, int n_tiles, int n_small, uint8_t *masks, int16_t *tmp,
                      bool compound, cudaStream_t st)
{
    if (n_tiles <= 0 || !descs || !tiles) return 0;
    McArgs a;
    memset(&a, 0, sizeof(a));
    a.dst = dst;
    for (int i = 0; i < 7; i++) { a.refs[i] = refs[i]; a.tmaps[i] = refs[i].tma; }
    a.descs = descs;
    a.tiles = tiles;
    a.n_tiles = n_tiles;
    a.n_small = n_small;
    a.masks = masks;
    a.tmp = tmp;
    return compound ? mc_compound_launch(a, st) : mc_put_launch(a, st);
}

void mc_init_attrs() {
    cudaFuncSetAttribute(mc_compound_tma_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(MC_WARPS_TC * sizeof(McSmemTmaCompound<uint16_t>)));
    cudaFuncSetAttribute(mc_compound_tma_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(MC_WARPS_TC * sizeof(McSmemTmaCompound<uint8_t>)));
    cudaFuncSetAttribute(mc_put_tma_kernel<uint16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(MC_WARPS * sizeof(McSmemTma<uint16_t>)));
    cudaFuncSetAttribute(mc_put_tma_kernel<uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(MC_WARPS * sizeof(McSmemTma<uint8_t>)));
    cudaFuncSetAttribute(mc_put_tma_kernel<uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(MC_WARPS * sizeof(McSmemTma<uint16_t>)));
    cudaFuncSetAttribute(mc_put_tma_kernel<uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(MC_WARPS * sizeof(McSmemTma<uint8_t>)));
    cudaFuncSetAttribute(mc_compound_kernel<uint16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(MC_WARPS * sizeof(McSmemCompound<uint16_t, 32>)));
    cudaFuncSetAttribute(mc_compound_kernel<uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(MC_WARPS * sizeof(McSmemCompound<uint8_t, 32>)));
}

// ------------------------------------------------------------ per-call surface
static void plane_of(PlaneView &p, void *data, size_t stride, int w, int h) {
    p.data = data; p.stride = (int64_t)stride; p.w = w; p.h = h;
}

// mc_fn / mct_fn (src/mc.h:38-63)
template <typename pixel, bool PREP>
static void mc_single(const int filter, void *out_host, const ptrdiff_t out_stride_bytes,
                      const pixel *src, const ptrdiff_t src_stride, const int w, const int h,
                      const int mx, const int my, const int bdmax)
{
    const bool bil = filter == 9;
    const int cl = mx ? (bil ? 0 : 3) : 0, cr = mx ? (bil ? 1 : 4) : 0;
    const int rl = my ? (bil ? 0 : 3) : 0, rb = my ? (bil ? 1 : 4) : 0;
    const int ww = w + cl + cr, wh = h + rl + rb;
    const size_t sstride = ((size_t)ww * sizeof(pixel) + 63) & ~(size_t)63;
    const size_t ostride = PREP ? (size_t)w * 2 : (((size_t)w * sizeof(pixel) + 63) & ~(size_t)63);
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t o_src = st.reserve(sstride * wh);
    const size_t o_desc = st.reserve(sizeof(Dav1dCudaMcDesc));
    std::vector<uint32_t> tiles = tiles_for(0, w, h);
    const size_t o_tiles = st.reserve(tiles.size() * 4);
    const size_t o_out = st.reserve(ostride * h);
    if (!st.commit()) return;
    const ptrdiff_t spx = src_stride / (ptrdiff_t)sizeof(pixel);
    st.put2d(o_src, sstride, src - rl * spx - cl, src_stride, (size_t)ww * sizeof(pixel), wh);
    Dav1dCudaMcDesc d;
    memset(&d, 0, sizeof(d));
    d.w = (uint8_t)w; d.h = (uint8_t)h;
    d.kind = PREP ? DAV1D_CUDA_MC_PREP : DAV1D_CUDA_MC_PUT;
    d.src[0].x = cl; d.src[0].y = rl;
    d.src[0].filter_2d = (uint8_t)filter; d.src[0].mx = (uint8_t)mx; d.src[0].my = (uint8_t)my;
    memcpy(st.host(o_desc), &d, sizeof(d));
    memcpy(st.host(o_tiles), tiles.data(), tiles.size() * 4);
    if (!st.upload()) return;
    McArgs a;
    memset(&a, 0, sizeof(a));
    plane_of(a.refs[0].p[0], st.dev(o_src), sstride, ww, wh);
    plane_of(a.dst.p[0], st.dev(o_out), ostride, w, h);
    a.dst.bdmax = bdmax;
    a.descs = (const Dav1dCudaMcDesc *)st.dev(o_desc);
    a.tiles = (const uint32_t *)st.dev(o_tiles);
    a.n_tiles = (int)tiles.size();
    a.n_small = (w <= 8 && h <= 8) ? a.n_tiles : 0;
    a.tmp = (int16_t *)st.dev(o_out);
    if (mc_put_launch(a, st.stream())) return;
    if (!st.download(o_out, ostride * h) || !st.sync()) return;
    st.get2d(o_out, ostride, out_host, PREP ? (ptrdiff_t)w * 2 : out_stride_bytes,
             PREP ? (size_t)w * 2 : (size_t)w * sizeof(pixel), h);
}

// mc_scaled_fn / mct_scaled_fn (src/mc.h:45-69)
template <typename pixel, bool PREP>
static void mc_scaled_single(const int filter, void *out_host, const ptrdiff_t out_stride_bytes,
                             const pixel *src, const ptrdiff_t src_stride, const int w, const int h,
                             const int mx, const int my, const int dx, const int dy, const int bdmax)
{
    const bool bil = filter == 9;
    // extent the reference touches (mc_tmpl.c:182-201, 459-478)
    const int cl = bil ? 0 : 3, rl = bil ? 0 : 3;
    const int last_x = ((mx + (w - 1) * dx) >> 10) + (bil ? 1 : 4);
    const int last_y = ((my + (h - 1) * dy) >> 10) + (bil ? 1 : 4);
    const int ww = cl + last_x + 1, wh = rl + last_y + 1;
    const size_t sstride = ((size_t)ww * sizeof(pixel) + 63) & ~(size_t)63;
    const size_t ostride = PREP ? (size_t)w * 2 : (((size_t)w * sizeof(pixel) + 63) & ~(size_t)63);
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t o_src = st.reserve(sstride * wh);
    const size_t o_out = st.reserve(ostride * h);
    if (!st.commit()) return;
    const ptrdiff_t spx = src_stride / (ptrdiff_t)sizeof(pixel);
    st.put2d(o_src, sstride, src - rl * spx - cl, src_stride, (size_t)ww * sizeof(pixel), wh);
    if (!st.upload()) return;
    ScaledArgs a;
    plane_of(a.ref, st.dev(o_src), sstride, ww, wh);
    a.sx = cl; a.sy = rl;
    a.w = w; a.h = h; a.mx = mx; a.my = my; a.dx = dx; a.dy = dy;
    a.filter_2d = filter;
    a.out = st.dev(o_out);
    a.ostride = (int)(PREP ? w : ostride / sizeof(pixel));
    a.bdmax = bdmax;
    const int ntiles = ((w + MC_T - 1) / MC_T) * ((h + MC_T - 1) / MC_T);
    mc_scaled_kernel<pixel, PREP><<<(ntiles + MC_WARPS - 1) / MC_WARPS, MC_WARPS * 32, 0, st.stream()>>>(a);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "mc_scaled_kernel")) return;
    if (!st.download(o_out, ostride * h) || !st.sync()) return;
    st.get2d(o_out, ostride, out_host, PREP ? (ptrdiff_t)w * 2 : out_stride_bytes,
             PREP ? (size_t)w * 2 : (size_t)w * sizeof(pixel), h);
}

// avg_fn / w_avg_fn / mask_fn / w_mask_fn (src/mc.h:71-93)
template <typename pixel>
static void combine_single(const int kind, pixel *dst, const ptrdiff_t dst_stride, const int16_t *t1,
                           const int16_t *t2, const int w, const int h, const int weight_or_sign,
                           uint8_t *mask, const int mask_ss, const int bdmax)
{
    const size_t tb = (size_t)w * h * 2;
    const size_t ostride = ((size_t)w * sizeof(pixel) + 63) & ~(size_t)63;
    const int ssh = mask_ss >= 1, ssv = mask_ss == 2;
    const size_t mbytes = kind == DAV1D_CUDA_MC_MASK ? (size_t)w * h
                        : kind == DAV1D_CUDA_MC_W_MASK ? (size_t)(w >> ssh) * (h >> ssv) : 0;
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t o1 = st.reserve(tb), o2 = st.reserve(tb);
    const size_t om = st.reserve(mbytes ? mbytes : 1);
    const size_t o_out = st.reserve(ostride * h);
    if (!st.commit()) return;
    memcpy(st.host(o1), t1, tb);
    memcpy(st.host(o2), t2, tb);
    if (kind == DAV1D_CUDA_MC_MASK) memcpy(st.host(om), mask, mbytes);
    if (!st.upload()) return;
    BlockOp op;
    memset(&op, 0, sizeof(op));
    op.kind = kind; op.w = w; op.h = h; op.weight = weight_or_sign; op.mask_ss = mask_ss;
    op.dst = st.dev(o_out); op.dstride = (int)(ostride / sizeof(pixel));
    op.a = st.dev(o1); op.b = st.dev(o2);
    op.mask = st.dev(om);
    op.bdmax = bdmax;
    mc_combine_kernel<pixel><<<1, 256, 0, st.stream()>>>(op);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "mc_combine_kernel")) return;
    if (!st.download(om, (size_t)(o_out - om) + ostride * h) || !st.sync()) return;
    st.get2d(o_out, ostride, dst, dst_stride, (size_t)w * sizeof(pixel), h);
    if (kind == DAV1D_CUDA_MC_W_MASK) memcpy(mask, st.host(om), mbytes);
}

// blend_fn / blend_dir_fn (src/mc.h:95-102)
template <typename pixel>
static void blend_single(const int kind, pixel *dst, const ptrdiff_t dst_stride, const pixel *tmp,
                         const int w, const int h, const uint8_t *mask)
{
    const size_t tbytes = (size_t)w * h * sizeof(pixel);
    const size_t ostride = ((size_t)w * sizeof(pixel) + 63) & ~(size_t)63;
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t ot = st.reserve(tbytes), om = st.reserve((size_t)w * h);
    const size_t o_out = st.reserve(ostride * h);
    if (!st.commit()) return;
    memcpy(st.host(ot), tmp, tbytes);
    if (kind == 0) memcpy(st.host(om), mask, (size_t)w * h);
    st.put2d(o_out, ostride, dst, dst_stride, (size_t)w * sizeof(pixel), h);
    if (!st.upload()) return;
    BlockOp op;
    memset(&op, 0, sizeof(op));
    op.kind = kind; op.w = w; op.h = h;
    op.dst = st.dev(o_out); op.dstride = (int)(ostride / sizeof(pixel));
    op.a = st.dev(ot);
    op.mask = st.dev(om);
    mc_blend_kernel<pixel><<<1, 256, 0, st.stream()>>>(op);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "mc_blend_kernel")) return;
    if (!st.download(o_out, ostride * h) || !st.sync()) return;
    st.get2d(o_out, ostride, dst, dst_stride, (size_t)w * sizeof(pixel), h);
}

// warp8x8_fn / warp8x8t_fn (src/mc.h:52-69)
template <typename pixel, bool PREP>
static void warp_single(void *out_host, const ptrdiff_t out_stride_bytes, const pixel *src,
                        const ptrdiff_t src_stride, const int16_t *abcd, const int mx, const int my,
                        const int bdmax)
{
    const size_t sstride = 64;   // 15 pixels
    const size_t obytes = PREP ? 2 : sizeof(pixel);
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t o_src = st.reserve(sstride * 15);
    const size_t o_out = st.reserve(8 * 8 * obytes);
    if (!st.commit()) return;
    const ptrdiff_t spx = src_stride / (ptrdiff_t)sizeof(pixel);
    st.put2d(o_src, sstride, src - 3 * spx - 3, src_stride, 15 * sizeof(pixel), 15);
    if (!st.upload()) return;
    WarpArgs w;
    plane_of(w.ref, st.dev(o_src), sstride, 15, 15);
    w.sx = 3; w.sy = 3;
    memcpy(w.abcd, abcd, sizeof(w.abcd));
    w.mx = mx; w.my = my;
    w.out = st.dev(o_out); w.ostride = 8;
    w.bdmax = bdmax;
    mc_warp_kernel<pixel, PREP><<<1, 32, 0, st.stream()>>>(w);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "mc_warp_kernel")) return;
    if (!st.download(o_out, 8 * 8 * obytes) || !st.sync()) return;
    st.get2d(o_out, 8 * obytes, out_host, out_stride_bytes, 8 * obytes, 8);
}

// emu_edge_fn (src/mc.h:104-107)
template <typename pixel>
static void emu_edge_single(const intptr_t bw, const intptr_t bh, const intptr_t iw, const intptr_t ih,
                            const intptr_t x, const intptr_t y, pixel *dst, const ptrdiff_t dst_stride,
                            const pixel *ref, const ptrdiff_t ref_stride)
{
    // visible sub-rectangle of the reference that the block maps to
    const int cx0 = iclip((int)x, 0, (int)iw - 1), cx1 = iclip((int)(x + bw - 1), 0, (int)iw - 1);
    const int cy0 = iclip((int)y, 0, (int)ih - 1), cy1 = iclip((int)(y + bh - 1), 0, (int)ih - 1);
    const int rw = cx1 - cx0 + 1, rh = cy1 - cy0 + 1;
    const size_t sstride = ((size_t)rw * sizeof(pixel) + 63) & ~(size_t)63;
    const size_t ostride = ((size_t)bw * sizeof(pixel) + 63) & ~(size_t)63;
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t o_src = st.reserve(sstride * rh);
    const size_t o_out = st.reserve(ostride * bh);
    if (!st.commit()) return;
    const ptrdiff_t rpx = ref_stride / (ptrdiff_t)sizeof(pixel);
    st.put2d(o_src, sstride, ref + cy0 * rpx + cx0, ref_stride, (size_t)rw * sizeof(pixel), rh);
    if (!st.upload()) return;
    EmuArgs e;
    plane_of(e.ref, st.dev(o_src), sstride, rw, rh);
    e.x = (int)x - cx0; e.y = (int)y - cy0;
    e.bw = (int)bw; e.bh = (int)bh;
    e.dst = st.dev(o_out); e.dstride = (int)(ostride / sizeof(pixel));
    mc_emu_edge_kernel<pixel><<<8, 256, 0, st.stream()>>>(e);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "mc_emu_edge_kernel")) return;
    if (!st.download(o_out, ostride * bh) || !st.sync()) return;
    st.get2d(o_out, ostride, dst, dst_stride, (size_t)bw * sizeof(pixel), (int)bh);
}

// resize_fn (src/mc.h:110-114, mc_tmpl.c:877-903): horizontal 8-tap upscaling of super-resolution.  The
// reference walks a row with a running position (mx += dx; src_x += mx >> 14; mx &= 0x3fff); with
// T = mx0 + x * dx that is filter (T & 0x3fff) >> 8 at source column -1 + (T >> 14): a thread per output pixel.
struct ResizeArgs {
    const void *src; int64_t sstride;      // pixels
    void *dst; int64_t dstride;
    int dst_w, h, src_w, dx, mx0, bdmax;
};
template <typename pixel> __global__ void __launch_bounds__(256) mc_resize_kernel(const ResizeArgs a) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= a.dst_w) return;
    const int64_t T = (int64_t)a.mx0 + (int64_t)x * a.dx;       // mx0 = get_upscale_x0() & 0x3fff (decode.c:3582)
    const int mx = (int)(T & 0x3fff);
    const int src_x = -1 + (int)(T >> 14);
    const int8_t *const F = g_resize_filter + (mx >> 8) * 8;
    // positions and taps are the column's: in registers once, then RESIZE_ROWS rows per thread
    int xs[8], f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { xs[k] = iclip(src_x - 3 + k, 0, a.src_w - 1); f[k] = F[k]; }
    for (int y = blockIdx.y; y < a.h; y += gridDim.y) {
        const pixel *s = (const pixel *)a.src + y * a.sstride;
        int sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) sum += f[k] * (int)s[xs[k]];
        ((pixel *)a.dst)[y * a.dstride + x] = (pixel)clip_px<pixel>((-sum + 64) >> 7, a.bdmax);
    }
}
static int resize_launch(const ResizeArgs &a, const bool hbd, cudaStream_t st) {
    if (a.dst_w <= 0 || a.h <= 0) return 0;
    constexpr int RESIZE_ROWS = 8;
    const dim3 grid((unsigned)((a.dst_w + 255) / 256), (unsigned)std::min((a.h + RESIZE_ROWS - 1) / RESIZE_ROWS, 4096));
    if (hbd) mc_resize_kernel<uint16_t><<<grid, 256, 0, st>>>(a);
    else mc_resize_kernel<uint8_t><<<grid, 256, 0, st>>>(a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "mc_resize_kernel") ? 0 : -5;
}
// super-resolution of a whole picture (dav1d_filter_sbrow_resize over every superblock row,
// recon_tmpl.c:2104-2137: rows are independent, so the frame is one launch per plane)
int mc_resize_frame(const PicView &dst, const PicView &src, const int step[2], const int start[2], cudaStream_t st) {
    for (int pl = 0; pl < 3; pl++) {
        if (!dst.p[pl].data || !src.p[pl].data) continue;
        const int px = dst.bdmax > 0xff ? 2 : 1;
        ResizeArgs a;
        a.src = src.p[pl].data; a.sstride = src.p[pl].stride / px;
        a.dst = dst.p[pl].data; a.dstride = dst.p[pl].stride / px;
        a.dst_w = dst.p[pl].w; a.h = std::min(dst.p[pl].h, src.p[pl].h); a.src_w = src.p[pl].w;
        a.dx = step[!!pl]; a.mx0 = start[!!pl]; a.bdmax = dst.bdmax;
        const int r = resize_launch(a, px == 2, st);
        if (r) return r;
    }
    return 0;
}
template <typename pixel>
static void resize_single(pixel *dst, const ptrdiff_t dst_stride, const pixel *src, const ptrdiff_t src_stride,
                          const int dst_w, const int h, const int src_w, const int dx, const int mx0, const int bdmax)
{
    if (dst_w <= 0 || h <= 0 || src_w <= 0) return;
    const size_t sstride = ((size_t)src_w * sizeof(pixel) + 63) & ~(size_t)63;
    const size_t ostride = ((size_t)dst_w * sizeof(pixel) + 63) & ~(size_t)63;
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t o_src = st.reserve(sstride * h);
    const size_t o_out = st.reserve(ostride * h);
    if (!st.commit()) return;
    st.put2d(o_src, sstride, src, src_stride, (size_t)src_w * sizeof(pixel), h);
    if (!st.upload()) return;
    ResizeArgs a;
    a.src = st.dev(o_src); a.sstride = (int64_t)(sstride / sizeof(pixel));
    a.dst = st.dev(o_out); a.dstride = (int64_t)(ostride / sizeof(pixel));
    a.dst_w = dst_w; a.h = h; a.src_w = src_w; a.dx = dx; a.mx0 = mx0; a.bdmax = bdmax;
    if (resize_launch(a, sizeof(pixel) == 2, st.stream())) return;
    if (!st.download(o_out, ostride * h) || !st.sync()) return;
    st.get2d(o_out, ostride, dst, dst_stride, (size_t)dst_w * sizeof(pixel), h);
}
static void resize8(uint8_t *d, ptrdiff_t ds, const uint8_t *s, ptrdiff_t ss, int dst_w, int h, int src_w, int dx, int mx0)
{ resize_single<uint8_t>(d, ds, s, ss, dst_w, h, src_w, dx, mx0, 0xff); }
static void resize16(uint16_t *d, ptrdiff_t ds, const uint16_t *s, ptrdiff_t ss, int dst_w, int h, int src_w, int dx, int mx0,
                     int bitdepth_max)
{ resize_single<uint16_t>(d, ds, s, ss, dst_w, h, src_w, dx, mx0, bitdepth_max); }

// ---- typed entry points filled into the table
#define HBD_ARGS , int bitdepth_max
template <int F> static void put8(uint8_t *d, ptrdiff_t ds, const uint8_t *s, ptrdiff_t ss, int w, int h, int mx, int my)
{ mc_single<uint8_t, false>(F, d, ds, s, ss, w, h, mx, my, 0xff); }
template <int F> static void put16(uint16_t *d, ptrdiff_t ds, const uint16_t *s, ptrdiff_t ss, int w, int h, int mx, int my HBD_ARGS)
{ mc_single<uint16_t, false>(F, d, ds, s, ss, w, h, mx, my, bitdepth_max); }
template <int F> static void prep8(int16_t *t, const uint8_t *s, ptrdiff_t ss, int w, int h, int mx, int my)
{ mc_single<uint8_t, true>(F, t, 0, s, ss, w, h, mx, my, 0xff); }
template <int F> static void prep16(int16_t *t, const uint16_t *s, ptrdiff_t ss, int w, int h, int mx, int my HBD_ARGS)
{ mc_single<uint16_t, true>(F, t, 0, s, ss, w, h, mx, my, bitdepth_max); }
template <int F> static void puts8(uint8_t *d, ptrdiff_t ds, const uint8_t *s, ptrdiff_t ss, int w, int h, int mx, int my, int dx, int dy)
{ mc_scaled_single<uint8_t, false>(F, d, ds, s, ss, w, h, mx, my, dx, dy, 0xff); }
template <int F> static void puts16(uint16_t *d, ptrdiff_t ds, const uint16_t *s, ptrdiff_t ss, int w, int h, int mx, int my, int dx, int dy HBD_ARGS)
{ mc_scaled_single<uint16_t, false>(F, d, ds, s, ss, w, h, mx, my, dx, dy, bitdepth_max); }
template <int F> static void preps8(int16_t *t, const uint8_t *s, ptrdiff_t ss, int w, int h, int mx, int my, int dx, int dy)
{ mc_scaled_single<uint8_t, true>(F, t, 0, s, ss, w, h, mx, my, dx, dy, 0xff); }
template <int F> static void preps16(int16_t *t, const uint16_t *s, ptrdiff_t ss, int w, int h, int mx, int my, int dx, int dy HBD_ARGS)
{ mc_scaled_single<uint16_t, true>(F, t, 0, s, ss, w, h, mx, my, dx, dy, bitdepth_max); }

static void avg8(uint8_t *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h)
{ combine_single<uint8_t>(DAV1D_CUDA_MC_AVG, d, ds, a, b, w, h, 0, nullptr, 0, 0xff); }
static void avg16(uint16_t *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h HBD_ARGS)
{ combine_single<uint16_t>(DAV1D_CUDA_MC_AVG, d, ds, a, b, w, h, 0, nullptr, 0, bitdepth_max); }
static void wavg8(uint8_t *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h, int wt)
{ combine_single<uint8_t>(DAV1D_CUDA_MC_W_AVG, d, ds, a, b, w, h, wt, nullptr, 0, 0xff); }
static void wavg16(uint16_t *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h, int wt HBD_ARGS)
{ combine_single<uint16_t>(DAV1D_CUDA_MC_W_AVG, d, ds, a, b, w, h, wt, nullptr, 0, bitdepth_max); }
static void mask8(uint8_t *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h, const uint8_t *m)
{ combine_single<uint8_t>(DAV1D_CUDA_MC_MASK, d, ds, a, b, w, h, 0, (uint8_t *)m, 0, 0xff); }
static void mask16(uint16_t *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h, const uint8_t *m HBD_ARGS)
{ combine_single<uint16_t>(DAV1D_CUDA_MC_MASK, d, ds, a, b, w, h, 0, (uint8_t *)m, 0, bitdepth_max); }
template <int SS> static void wmask8(uint8_t *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h, uint8_t *m, int sign)
{ combine_single<uint8_t>(DAV1D_CUDA_MC_W_MASK, d, ds, a, b, w, h, sign, m, SS, 0xff); }
template <int SS> static void wmask16(uint16_t *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h, uint8_t *m, int sign HBD_ARGS)
{ combine_single<uint16_t>(DAV1D_CUDA_MC_W_MASK, d, ds, a, b, w, h, sign, m, SS, bitdepth_max); }
template <typename pixel> static void blend_p(pixel *d, ptrdiff_t ds, const pixel *t, int w, int h, const uint8_t *m)
{ blend_single<pixel>(0, d, ds, t, w, h, m); }
template <typename pixel> static void blend_v_p(pixel *d, ptrdiff_t ds, const pixel *t, int w, int h)
{ blend_single<pixel>(1, d, ds, t, w, h, nullptr); }
template <typename pixel> static void blend_h_p(pixel *d, ptrdiff_t ds, const pixel *t, int w, int h)
{ blend_single<pixel>(2, d, ds, t, w, h, nullptr); }
static void warp8(uint8_t *d, ptrdiff_t ds, const uint8_t *s, ptrdiff_t ss, const int16_t *abcd, int mx, int my)
{ warp_single<uint8_t, false>(d, ds, s, ss, abcd, mx, my, 0xff); }
static void warp16(uint16_t *d, ptrdiff_t ds, const uint16_t *s, ptrdiff_t ss, const int16_t *abcd, int mx, int my HBD_ARGS)
{ warp_single<uint16_t, false>(d, ds, s, ss, abcd, mx, my, bitdepth_max); }
static void warpt8(int16_t *t, ptrdiff_t ts, const uint8_t *s, ptrdiff_t ss, const int16_t *abcd, int mx, int my)
{ warp_single<uint8_t, true>(t, ts * 2, s, ss, abcd, mx, my, 0xff); }
static void warpt16(int16_t *t, ptrdiff_t ts, const uint16_t *s, ptrdiff_t ss, const int16_t *abcd, int mx, int my HBD_ARGS)
{ warp_single<uint16_t, true>(t, ts * 2, s, ss, abcd, mx, my, bitdepth_max); }
template <typename pixel> static void emu_p(intptr_t bw, intptr_t bh, intptr_t iw, intptr_t ih, intptr_t x, intptr_t y,
                                           pixel *d, ptrdiff_t ds, const pixel *r, ptrdiff_t rs)
{ emu_edge_single<pixel>(bw, bh, iw, ih, x, y, d, ds, r, rs); }

template <bool HBD, int F> static void fill_filter(Dav1dCudaMCDSPContext *c) {
    c->mc[F] = HBD ? (void *)put16<F> : (void *)put8<F>;
    c->mct[F] = HBD ? (void *)prep16<F> : (void *)prep8<F>;
    c->mc_scaled[F] = HBD ? (void *)puts16<F> : (void *)puts8<F>;
    c->mct_scaled[F] = HBD ? (void *)preps16<F> : (void *)preps8<F>;
}

template <bool HBD> static void fill_mc(Dav1dCudaMCDSPContext *c) {
    Staging &s = staging();
    {
        std::lock_guard<std::mutex> lk(s.mu);
        if (!s.ensure(1 << 20)) return;
    }
    fill_filter<HBD, 0>(c); fill_filter<HBD, 1>(c); fill_filter<HBD, 2>(c); fill_filter<HBD, 3>(c);
    fill_filter<HBD, 4>(c); fill_filter<HBD, 5>(c); fill_filter<HBD, 6>(c); fill_filter<HBD, 7>(c);
    fill_filter<HBD, 8>(c); fill_filter<HBD, 9>(c);
    typedef typename std::conditional<HBD, uint16_t, uint8_t>::type pixel;
    c->avg = HBD ? (void *)avg16 : (void *)avg8;
    c->w_avg = HBD ? (void *)wavg16 : (void *)wavg8;
    c->mask = HBD ? (void *)mask16 : (void *)mask8;
    c->w_mask[0] = HBD ? (void *)wmask16<0> : (void *)wmask8<0>;
    c->w_mask[1] = HBD ? (void *)wmask16<1> : (void *)wmask8<1>;
    c->w_mask[2] = HBD ? (void *)wmask16<2> : (void *)wmask8<2>;
    c->blend = (void *)blend_p<pixel>;
    c->blend_v = (void *)blend_v_p<pixel>;
    c->blend_h = (void *)blend_h_p<pixel>;
    c->warp8x8 = HBD ? (void *)warp16 : (void *)warp8;
    c->warp8x8t = HBD ? (void *)warpt16 : (void *)warpt8;
    c->emu_edge = (void *)emu_p<pixel>;
    c->resize = HBD ? (void *)resize16 : (void *)resize8;
}

}  // namespace d1

using namespace d1;

extern "C" {

void dav1d_cuda_mc_dsp_init_8bpc(Dav1dCudaMCDSPContext *c) { fill_mc<false>(c); }
void dav1d_cuda_mc_dsp_init_16bpc(Dav1dCudaMCDSPContext *c) { fill_mc<true>(c); }

int dav1d_cuda_resize_frame(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, const Dav1dCudaPicture *src,
                            const int32_t resize_step[2], const int32_t resize_start[2])
{
    if (!c || !dst || !src || !resize_step || !resize_start) return -22;
    if (dst->bitdepth_max != src->bitdepth_max || dst->ss_hor != src->ss_hor || dst->ss_ver != src->ss_ver ||
        dst->p[0].h != src->p[0].h || dst->p[0].data == src->p[0].data) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    const int step[2] = { resize_step[0], resize_step[1] }, start[2] = { resize_start[0], resize_start[1] };
    return mc_resize_frame(pic_view(dst), pic_view(src), step, start, c->stream);
}

static int mc_batch_common(Dav1dCudaContext *c, const Dav1dCudaPicture *dst,
                           const Dav1dCudaPicture *const refs[7], const Dav1dCudaMcDesc *descs,
                           const uint32_t *tiles, int n_tiles, int n_small, uint8_t *masks, int16_t *tmp,
                           bool compound)
{
    if (!c || !dst || !descs || !tiles) return -22;
    McArgs a;
    memset(&a, 0, sizeof(a));
    a.dst = pic_view(dst);
    for (int i = 0; i < 7; i++)
        if (refs[i]) a.refs[i] = pic_view(refs[i]);
    a.descs = descs;
    a.tiles = tiles;
    a.n_tiles = n_tiles;
    a.n_small = n_small;
    a.masks = masks;
    a.tmp = tmp;
    return compound ? mc_compound_launch(a, c->stream) : mc_put_launch(a, c->stream);
}

int dav1d_cuda_mc_tiles(uint32_t desc_index, int w, int h, uint32_t *out) {
    int n = 0;
    for (int ty = 0; ty * MC_T < h; ty++)
        for (int tx = 0; tx * MC_T < w; tx++)
            out[n++] = desc_index * 16 + ty * 4 + tx;
    return n;
}

int dav1d_cuda_mc_put_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst,
                            const Dav1dCudaPicture *const refs[7], const Dav1dCudaMcDesc *descs,
                            const uint32_t *tiles, int n_tiles, int n_small, int16_t *tmp)
{
    return mc_batch_common(c, dst, refs, descs, tiles, n_tiles, n_small, nullptr, tmp, false);
}

int dav1d_cuda_mc_compound_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst,
                                 const Dav1dCudaPicture *const refs[7], const Dav1dCudaMcDesc *descs,
                                 const uint32_t *tiles, int n_tiles, int n_small, uint8_t *masks)
{
    return mc_batch_common(c, dst, refs, descs, tiles, n_tiles, n_small, masks, nullptr, true);
}

}  // extern "C"
