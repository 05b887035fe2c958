// Frame-level batched reconstruction: the fused intra-class kernel (edge
// preparation + prediction + residual, one warp per transform block, one
// launch per dependency level), the host-side level scheduler, and the
// per-frame submit / CUDA-graph entry points.
//
// Reference call sites replaced: dav1d_recon_b_intra (src/recon_tmpl.c:1195-1596)
// and, through the MC / ITX launches, dav1d_recon_b_inter (:1598-2036).
#include <string.h>
#include <algorithm>
#include <vector>
#include "ctx.h"
#include "itx.cuh"
#include "ipred.cuh"
#include "mc.cuh"

namespace d1 {

// defined in itx.cu / mc.cu
int itx_batch_launch(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs,
                     const int32_t *class_count, int zero_coefs, cudaStream_t st);
struct McArgs;
int mc_put_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcDesc *descs,
                      const uint32_t *tiles, int n_tiles, uint8_t *masks, int16_t *tmp, bool compound,
                      cudaStream_t st);

constexpr int INTRA_WARPS = 4;
constexpr int EDGE_BUF = 288;
constexpr int EDGE_C = 144;
constexpr int INTRA_TILE_INTS = 32 * 65;

template <typename pixel> struct IntraSmem {
    int tile[INTRA_TILE_INTS];
    int16_t ac[32 * 32];
    pixel edge[EDGE_BUF];
    pixel scratch[IPRED_SCRATCH];
};

struct IntraArgs {
    PicView pic;
    int bw4, bh4;
    void *cf;
    const Dav1dCudaIntraDesc *descs;
    int n;
    const void *pal;
    const uint8_t *pal_idx;
};

template <typename pixel, int W, int H>
DEV void intra_residual(int *tile, void *cf, const Dav1dCudaIntraDesc &d, pixel *dst, const int stride,
                        const int bdmax, const int lane)
{
    typedef typename PxTraits<pixel>::coef coef;
    itx_block<pixel, W, H, 32>(true, lane, tile, (coef *)cf + d.coef_off, d.eob, d.txtp, dst, stride, bdmax, false);
}

template <typename pixel>
__global__ void __launch_bounds__(INTRA_WARPS * 32) intra_level_kernel(const __grid_constant__ IntraArgs a) {
    extern __shared__ __align__(16) uint8_t intra_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * INTRA_WARPS + warp;
    if (idx >= a.n) return;
    IntraSmem<pixel> *sm = (IntraSmem<pixel> *)intra_smem_raw + warp;
    const Dav1dCudaIntraDesc d = a.descs[idx];
    const int pl = d.plane;
    const int ss_hor = pl ? a.pic.ss_hor : 0, ss_ver = pl ? a.pic.ss_ver : 0;
    const PlaneView &pv = a.pic.p[pl];
    const int stride = (int)(pv.stride / (int)sizeof(pixel));
    pixel *dst = (pixel *)pv.data + (int64_t)d.y4 * 4 * stride + d.x4 * 4;
    const int w = d.tw4 * 4, h = d.th4 * 4;
    const int bdmax = a.pic.bdmax;
    pixel *edge = sm->edge + EDGE_C;
    const int have_left = d.x4 > d.tile_x4_start, have_top = d.y4 > d.tile_y4_start;

    if (d.mode == DAV1D_CUDA_INTRA_PAL) {
        pal_pred_block<pixel>(dst, stride, (const pixel *)a.pal + d.aux, a.pal_idx + d.coef_off, w, h, lane, 32);
    } else if (d.mode == DAV1D_CUDA_INTRA_CFL) {
        const PlaneView &lv = a.pic.p[0];
        const int lstride = (int)(lv.stride / (int)sizeof(pixel));
        const pixel *luma = (const pixel *)lv.data + (int64_t)((d.y4 * 4) << ss_ver) * lstride + ((d.x4 * 4) << ss_hor);
        cfl_ac_block<pixel>(sm->ac, luma, lstride, d.aux & 0xff, (d.aux >> 8) & 0xff, w, h, ss_hor, ss_ver, lane);
        int angle = 0;
        const int m = prepare_edges<pixel>(d.x4, have_left, d.y4, have_top, d.tile_x4_end, d.tile_y4_end, 0, dst,
                                           stride, nullptr, 0, &angle, d.tw4, d.th4, 0, edge, bdmax, lane);
        cfl_pred_block<pixel>(m, dst, stride, edge, w, h, sm->ac, d.angle_delta, bdmax, lane);
    } else if (d.mode != DAV1D_CUDA_INTRA_NONE) {
        int angle = d.angle_delta;
        const int m = prepare_edges<pixel>(d.x4, have_left, d.y4, have_top, d.tile_x4_end, d.tile_y4_end,
                                           d.edge_flags, dst, stride, nullptr, d.mode, &angle, d.tw4, d.th4,
                                           (d.flags >> 10) & 1, edge, bdmax, lane);
        const int max_w = ((4 * a.bw4 + ss_hor) >> ss_hor) - 4 * d.x4;
        const int max_h = ((4 * a.bh4 + ss_ver) >> ss_ver) - 4 * d.y4;
        ipred_block<pixel>(m, dst, stride, edge, w, h, angle | d.flags, max_w, max_h, bdmax, sm->scratch, lane);
    }
    __syncwarp();
    if (d.eob < 0) return;
    switch (d.tx) {
    case 0: intra_residual<pixel, 4, 4>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 1: intra_residual<pixel, 8, 8>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 2: intra_residual<pixel, 16, 16>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 3: intra_residual<pixel, 32, 32>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 4: intra_residual<pixel, 64, 64>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 5: intra_residual<pixel, 4, 8>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 6: intra_residual<pixel, 8, 4>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 7: intra_residual<pixel, 8, 16>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 8: intra_residual<pixel, 16, 8>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 9: intra_residual<pixel, 16, 32>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 10: intra_residual<pixel, 32, 16>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 11: intra_residual<pixel, 32, 64>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 12: intra_residual<pixel, 64, 32>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 13: intra_residual<pixel, 4, 16>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 14: intra_residual<pixel, 16, 4>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 15: intra_residual<pixel, 8, 32>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 16: intra_residual<pixel, 32, 8>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    case 17: intra_residual<pixel, 16, 64>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    default: intra_residual<pixel, 64, 16>(sm->tile, a.cf, d, dst, stride, bdmax, lane); break;
    }
}

template <typename pixel>
static int launch_intra_level(const IntraArgs &a, cudaStream_t st) {
    const int grid = (a.n + INTRA_WARPS - 1) / INTRA_WARPS;
    intra_level_kernel<pixel><<<grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraSmem<pixel>), st>>>(a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "intra_level_kernel") ? 0 : -5;
}

static int intra_batch_launch(const PicView &pic, int bw4, int bh4, void *cf, const Dav1dCudaIntraDesc *descs,
                              const int32_t *level_start, int n_levels, const void *pal, const uint8_t *pal_idx,
                              cudaStream_t st)
{
    for (int l = 0; l < n_levels; l++) {
        const int n = level_start[l + 1] - level_start[l];
        if (n <= 0) continue;
        IntraArgs a;
        a.pic = pic; a.bw4 = bw4; a.bh4 = bh4; a.cf = cf;
        a.descs = descs + level_start[l];
        a.n = n;
        a.pal = pal; a.pal_idx = pal_idx;
        const int r = pic.bdmax > 0xff ? launch_intra_level<uint16_t>(a, st) : launch_intra_level<uint8_t>(a, st);
        if (r) return r;
    }
    return 0;
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

static int recon_submit_on(const Dav1dCudaReconBatch *b, cudaStream_t st) {
    const PicView dst = pic_view(b->dst);
    PicView refs[7];
    refs_view(refs, b->refs);
    int r;
    // phase A: prediction from reference frames
    if ((r = mc_put_launch_raw(dst, refs, b->mc_put, b->mc_put_tiles, b->n_mc_put_tiles, nullptr, nullptr, false, st)))
        return r;
    if ((r = mc_put_launch_raw(dst, refs, b->mc_comp, b->mc_comp_tiles, b->n_mc_comp_tiles[0], b->masks, nullptr,
                               true, st)))
        return r;
    if ((r = mc_put_launch_raw(dst, refs, b->mc_comp, b->mc_comp_tiles + b->n_mc_comp_tiles[0],
                               b->n_mc_comp_tiles[1], b->masks, nullptr, true, st)))
        return r;
    if ((r = warp_batch_launch(dst, refs, b->warp, b->n_warp, st))) return r;
    // phase B: inter residuals
    if (b->itx && (r = itx_batch_launch(dst, b->cf, b->itx, b->itx_class_count, 0, st))) return r;
    // phase C: intra-class operations, level by level
    if (b->intra && (r = intra_batch_launch(dst, b->bw4, b->bh4, b->cf, b->intra, b->intra_level_start, b->n_levels,
                                            b->pal, b->pal_idx, st)))
        return r;
    return 0;
}

void recon_init_attrs() {
    cudaFuncSetAttribute(intra_level_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint16_t>)));
    cudaFuncSetAttribute(intra_level_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint8_t>)));
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
    PicView rv[7];
    refs_view(rv, refs);
    return warp_batch_launch(pic_view(dst), rv, descs, n, c->stream);
}

// Level assignment, see include/dav1d_cuda.h.  Per plane a map of 4x4 cells
// holds the level at which the cell's pixels become final (0 = produced by
// the inter phases).
int dav1d_cuda_intra_schedule(Dav1dCudaIntraDesc *descs, int n, int bw4, int bh4, int ss_hor, int ss_ver,
                              int32_t *order, int32_t *level_start, int max_levels)
{
    if (!descs || n < 0 || !order || !level_start) return -22;
    const int pw[3] = { bw4, (bw4 + ss_hor) >> ss_hor, (bw4 + ss_hor) >> ss_hor };
    const int ph[3] = { bh4, (bh4 + ss_ver) >> ss_ver, (bh4 + ss_ver) >> ss_ver };
    std::vector<int32_t> map[3];
    for (int p = 0; p < 3; p++) map[p].assign((size_t)pw[p] * ph[p], 0);
    int n_levels = 0;
    for (int i = 0; i < n; i++) {
        Dav1dCudaIntraDesc &d = descs[i];
        const int p = d.plane, W = pw[p], H = ph[p];
        const int x0 = d.x4, y0 = d.y4, x1 = std::min<int>(x0 + d.tw4, W), y1 = std::min<int>(y0 + d.th4, H);
        int lv = 0;
        auto dep = [&](int pl, int x, int y) {
            if (x >= 0 && y >= 0 && x < pw[pl] && y < ph[pl]) lv = std::max(lv, map[pl][(size_t)y * pw[pl] + x]);
        };
        if (d.mode == DAV1D_CUDA_INTRA_NONE) {
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) dep(p, x, y);
        } else if (d.mode != DAV1D_CUDA_INTRA_PAL) {
            const int have_left = x0 > d.tile_x4_start, have_top = y0 > d.tile_y4_start;
            if (have_top) {
                const int xe = std::min<int>(x0 + d.tw4 + ((d.edge_flags & 1) ? d.tw4 : 0), d.tile_x4_end);
                for (int x = x0 - have_left; x < xe; x++) dep(p, x, y0 - 1);
            }
            if (have_left) {
                const int ye = std::min<int>(y0 + d.th4 + ((d.edge_flags & 8) ? d.th4 : 0), d.tile_y4_end);
                for (int y = y0; y < ye; y++) dep(p, x0 - 1, y);
            }
            if (d.mode == DAV1D_CUDA_INTRA_CFL) {
                const int sh = p ? ss_hor : 0, sv = p ? ss_ver : 0;
                for (int y = y0 << sv; y < ((y0 + d.th4) << sv); y++)
                    for (int x = x0 << sh; x < ((x0 + d.tw4) << sh); x++) dep(0, x, y);
            }
        }
        lv += 1;
        d.level = (uint32_t)lv;
        n_levels = std::max(n_levels, lv);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) map[p][(size_t)y * W + x] = lv;
    }
    if (n_levels > max_levels) return -34;
    std::vector<int32_t> cnt(n_levels + 2, 0);
    for (int i = 0; i < n; i++) cnt[descs[i].level]++;          // levels are 1-based
    level_start[0] = 0;
    for (int l = 1; l <= n_levels; l++) level_start[l] = level_start[l - 1] + cnt[l];
    std::vector<int32_t> pos(level_start, level_start + n_levels + 1);
    for (int i = 0; i < n; i++) order[pos[descs[i].level - 1]++] = i;
    return n_levels;
}

int dav1d_cuda_intra_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, int bw4, int bh4, void *cf,
                           const Dav1dCudaIntraDesc *descs, const int32_t *level_start, int n_levels,
                           const void *pal, const uint8_t *pal_idx)
{
    if (!c || !dst || !descs || !level_start) return -22;
    return intra_batch_launch(pic_view(dst), bw4, bh4, cf, descs, level_start, n_levels, pal, pal_idx, c->stream);
}

int dav1d_cuda_recon_submit(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b) {
    if (!c || !b || !b->dst) return -22;
    return recon_submit_on(b, c->stream);
}

int dav1d_cuda_recon_graph_build(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b, Dav1dCudaReconGraph **out) {
    if (!c || !b || !out) return -22;
    *out = nullptr;
    cudaStream_t cap;
    D1_CHECK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    D1_CHECK(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    const int r = recon_submit_on(b, cap);
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

int dav1d_cuda_recon_graph_launch(Dav1dCudaContext *c, Dav1dCudaReconGraph *g) {
    if (!c || !g) return -22;
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
