// Frame-level batched reconstruction: the fused intra-class kernel (edge
// preparation + prediction + residual, one warp per transform block, one
// launch per dependency level), the host-side level scheduler, and the
// per-frame submit / CUDA-graph entry points.
//
// Reference call sites replaced: dav1d_recon_b_intra (src/recon_tmpl.c:1195-1596)
// and, through the MC / ITX launches, dav1d_recon_b_inter (:1598-2036).
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "ctx.h"
#define D1_ITX_PASS_NOINLINE
#include "itx.cuh"
#include "ipred.cuh"
#include "mc.cuh"

namespace d1 {

// defined in itx.cu / mc.cu
int itx_batch_launch(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs,
                     const int32_t *class_count, int zero_coefs, cudaStream_t st);
int itx_batch_launch_multi(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs,
                           const int32_t *class_count, int zero_coefs, cudaStream_t *streams, int n_streams);
int itx_task_launch(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs, const uint32_t *tasks,
                    int n_small, int n_big, int zero_coefs, cudaStream_t st_small, cudaStream_t st_big);
int itx_build_tasks(const Dav1dCudaItxDesc *descs, int n, int index_base, uint32_t *tasks, int *n_small, int *n_big);
int mc_obmc_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcDesc *descs,
                       const uint32_t *tiles, int n_tiles, cudaStream_t st);
int itx_multi_task_launch(const ItxFrameRef *frames, const Dav1dCudaItxDesc *descs, const uint2 *tasks, int n_small,
                          int n_big, bool hbd, cudaStream_t st_small, cudaStream_t st_big);
void itx_init_attrs();
struct McArgs;
int mc_put_launch_raw(const PicView &dst, const PicView *refs, const Dav1dCudaMcDesc *descs,
                      const uint32_t *tiles, int n_tiles, int n_small, uint8_t *masks, int16_t *tmp,
                      bool compound, cudaStream_t st);

constexpr int INTRA_WARPS = 4;
constexpr int EDGE_BUF = 288;
constexpr int EDGE_C = 144;
constexpr int INTRA_TILE_INTS = 32 * 65;

// Size classes of intra-class operations: each class gets its own kernel
// instantiation (smaller code footprint -> instruction cache, fewer registers
// and less shared memory for the small sizes that dominate the count).
//   0 = any size (dataflow / multi-frame variants), 1 = up to 8x8,
//   2 = up to 16x16, 3 = larger, 4 = prediction only (residuals run as transform tasks)
template <int CLS> struct IntraCls {
    static constexpr int TILE_INTS = CLS == 4 ? 4 : CLS == 1 ? 8 * 9 : CLS == 2 ? 16 * 17 : INTRA_TILE_INTS;
    static constexpr int AC_N = CLS == 1 ? 8 * 8 : CLS == 2 ? 16 * 16 : 32 * 32;
    static constexpr int MIN_BLOCKS = CLS == 1 || CLS == 4 ? 8 : CLS == 2 ? 6 : 4;
};
HD int intra_size_class(const int w, const int h) { return (w <= 8 && h <= 8) ? 1 : (w <= 16 && h <= 16) ? 2 : 3; }

template <typename pixel, int CLS = 0> struct IntraSmem {
    int tile[IntraCls<CLS>::TILE_INTS];
    int16_t ac[IntraCls<CLS>::AC_N];
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
    // dataflow variant
    const int32_t *dep_start;
    const int32_t *deps;
    unsigned *sync;
    int opw;               // level kernels: operations per warp
};

template <typename pixel, int W, int H>
DEV void intra_residual(int *tile, void *cf, const Dav1dCudaIntraDesc &d, pixel *dst, const int stride,
                        const int bdmax, const int lane)
{
    typedef typename PxTraits<pixel>::coef coef;
    itx_block<pixel, W, H, 32>(true, lane, tile, (coef *)cf + d.coef_off, d.eob, d.txtp, dst, stride, bdmax, false,
                               d.cw4, d.ch4);
}

// One intra-class operation (prediction [+ residual]) by one warp.
template <typename pixel, int CLS>
__device__ __noinline__ void intra_op(const IntraArgs &a, const Dav1dCudaIntraDesc &d, IntraSmem<pixel, CLS> *sm,
                                      const int lane) {
    const int pl = d.plane;
    const int ss_hor = pl ? a.pic.ss_hor : 0, ss_ver = pl ? a.pic.ss_ver : 0;
    const PlaneView &pv = a.pic.p[pl];
    const int stride = (int)(pv.stride / (int)sizeof(pixel));
    pixel *dst = (pixel *)pv.data + (int64_t)d.y4 * 4 * stride + d.x4 * 4;
    const int w = d.tw4 * 4, h = d.th4 * 4;
    const int bdmax = a.pic.bdmax;
    pixel *edge = sm->edge + EDGE_C;
    const int have_left = d.x4 > d.tile_x4_start, have_top = d.y4 > d.tile_y4_start;

    if (CLS != 4 && d.eob >= 0) {
        // pull the block's coefficients towards the SM while the prediction runs
        typedef typename PxTraits<pixel>::coef coef;
        const int ncoef = d.cw4 ? 16 * d.cw4 * d.ch4 : imin(w, 32) * imin(h, 32);   // packed box or dense block
        const char *cp = (const char *)((const coef *)a.cf + d.coef_off);
        for (int o = lane * 128; o < ncoef * (int)sizeof(coef); o += 32 * 128)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(cp + o));
    }

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
    } else if (d.mode == DAV1D_CUDA_INTRA_IBC) {
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
            const int x0 = iclip(sx + x, 0, pw - 1), x1 = iclip(sx + x + 1, 0, pw - 1);
            const int y0 = iclip(sy + y, 0, ph - 1), y1 = iclip(sy + y + 1, 0, ph - 1);
            const int p00 = __ldcg(base + (int64_t)y0 * stride + x0);
            int out;
            if (mx && my) {
                const int p01 = __ldcg(base + (int64_t)y0 * stride + x1);
                const int p10 = __ldcg(base + (int64_t)y1 * stride + x0);
                const int p11 = __ldcg(base + (int64_t)y1 * stride + x1);
                const int sh1 = 4 - ib, r1 = (1 << sh1) >> 1;
                const int m0 = (16 * p00 + mx * (p01 - p00) + r1) >> sh1;
                const int m1 = (16 * p10 + mx * (p11 - p10) + r1) >> sh1;
                const int sh2 = 4 + ib;
                out = clip_px<pixel>((16 * m0 + my * (m1 - m0) + ((1 << sh2) >> 1)) >> sh2, bdmax);
            } else if (mx) {
                const int p01 = __ldcg(base + (int64_t)y0 * stride + x1);
                const int sh1 = 4 - ib;
                const int px = (16 * p00 + mx * (p01 - p00) + ((1 << sh1) >> 1)) >> sh1;
                out = clip_px<pixel>((px + ((1 << ib) >> 1)) >> ib, bdmax);
            } else if (my) {
                const int p10 = __ldcg(base + (int64_t)y1 * stride + x0);
                out = clip_px<pixel>((16 * p00 + my * (p10 - p00) + 8) >> 4, bdmax);
            } else {
                out = p00;
            }
            dst[y * stride + x] = (pixel)out;
        }
    } else if (d.mode == DAV1D_CUDA_INTRA_II) {
        // inter-intra: predict the whole block into scratch (the ac buffer holds w*h pixels of the
        // operation's size class), then mc.blend onto the inter prediction (mc_tmpl.c:642-653)
        int angle = 0;
        const int m = prepare_edges<pixel>(d.x4, have_left, d.y4, have_top, d.tile_x4_end, d.tile_y4_end, 0, dst,
                                           stride, nullptr, d.angle_delta, &angle, d.tw4, d.th4, 0, edge, bdmax, lane);
        pixel *tmp = (pixel *)sm->ac;
        ipred_block<pixel>(m, tmp, w, edge, w, h, 0, 0, 0, bdmax, sm->scratch, lane);
        __syncwarp();
        const uint8_t *mask = a.pal_idx + d.coef_off;
        const int lw = 31 - __clz(w);
        for (int i = lane; i < w * h; i += 32) {
            const int y = i >> lw, x = i & (w - 1);
            const int mk = mask[i], p = dst[y * stride + x], q = tmp[i];
            dst[y * stride + x] = (pixel)((p * (64 - mk) + q * mk + 32) >> 6);
        }
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
    if (CLS == 4 || d.eob < 0) return;
#define D1_TXCASE(T, W, H) \
    case T: \
        if constexpr (CLS == 0 || CLS == ((W <= 8 && H <= 8) ? 1 : (W <= 16 && H <= 16) ? 2 : 3)) \
            intra_residual<pixel, W, H>(sm->tile, a.cf, d, dst, stride, bdmax, lane); \
        break;
    switch (d.tx) {
    D1_TXCASE(0, 4, 4)
    D1_TXCASE(1, 8, 8)
    D1_TXCASE(2, 16, 16)
    D1_TXCASE(3, 32, 32)
    D1_TXCASE(4, 64, 64)
    D1_TXCASE(5, 4, 8)
    D1_TXCASE(6, 8, 4)
    D1_TXCASE(7, 8, 16)
    D1_TXCASE(8, 16, 8)
    D1_TXCASE(9, 16, 32)
    D1_TXCASE(10, 32, 16)
    D1_TXCASE(11, 32, 64)
    D1_TXCASE(12, 64, 32)
    D1_TXCASE(13, 4, 16)
    D1_TXCASE(14, 16, 4)
    D1_TXCASE(15, 8, 32)
    D1_TXCASE(16, 32, 8)
    D1_TXCASE(17, 16, 64)
    D1_TXCASE(18, 64, 16)
    default: break;
    }
#undef D1_TXCASE
}

// Level-synchronous variant: one launch per dependency level.
template <typename pixel, int CLS>
__global__ void __launch_bounds__(INTRA_WARPS * 32, IntraCls<CLS>::MIN_BLOCKS)
intra_level_kernel(const __grid_constant__ IntraArgs a) {
    extern __shared__ __align__(16) uint8_t intra_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * INTRA_WARPS + warp;
    pdl_launch_dependents();
    if (idx >= a.n) return;
    IntraSmem<pixel, CLS> *sm = (IntraSmem<pixel, CLS> *)intra_smem_raw + warp;
    const Dav1dCudaIntraDesc d = a.descs[idx];
    pdl_wait();
    intra_op<pixel, CLS>(a, d, sm, lane);
}

// Multi-frame variant: dependency level l of SEVERAL frames (independent
// streams) in one launch, so the per-level latency is shared by all of them.
// per-frame arguments in a device table: the kernel hands intra_op() a reference INTO the
// table (a per-thread copy would live in local memory: intra_op is not inlined)
typedef IntraArgs IntraFrameParams;
struct IntraMultiArgs {
    const IntraFrameParams *frames;
    const Dav1dCudaIntraDesc *items;   // this level's operations: descriptor copies, `pad` = frame index
    int n;
};

template <typename pixel, int CLS>
__global__ void __launch_bounds__(INTRA_WARPS * 32, IntraCls<CLS>::MIN_BLOCKS) intra_multi_kernel(const IntraMultiArgs m) {
    extern __shared__ __align__(16) uint8_t intra_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    IntraSmem<pixel, CLS> *sm = (IntraSmem<pixel, CLS> *)intra_smem_raw + warp;
    const int i = blockIdx.x * INTRA_WARPS + warp;
    if (i >= m.n) return;
    // the merged level carries descriptor COPIES (consecutive warps read consecutive descriptors,
    // one dependent load less than an index into the frame's own array)
    const Dav1dCudaIntraDesc d = m.items[i];
    const IntraFrameParams &a = m.frames[d.pad];
    intra_op<pixel, CLS>(a, d, sm, lane);
}

// Dataflow variant: ONE persistent launch for the whole intra phase.  Warps
// claim operations in level-sorted (= topological) order from a global
// counter and wait on the completion flags of exactly the operations whose
// pixels they read (dependency lists built by dav1d_cuda_intra_schedule()).
// Every dependency has a smaller sorted index, i.e. it was claimed earlier by
// a warp that is already running, so waiting cannot deadlock.
// sync[0] = claim counter, sync[1 + i] = completion flag of operation i
// (zeroed by a memset node before the launch).
DEV unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
DEV void st_release(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

template <typename pixel>
__global__ void __launch_bounds__(INTRA_WARPS * 32, 4) intra_flow_kernel(const __grid_constant__ IntraArgs a) {
    extern __shared__ __align__(16) uint8_t intra_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    IntraSmem<pixel> *sm = (IntraSmem<pixel> *)intra_smem_raw + warp;
    unsigned *counter = a.sync, *flags = a.sync + 1;
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = (int)atomicAdd(counter, 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= a.n) break;
        const Dav1dCudaIntraDesc d = a.descs[idx];
        const int d0 = a.dep_start[idx], d1 = a.dep_start[idx + 1];
        for (int k = d0 + lane; k < d1; k += 32) {
            const unsigned *f = flags + a.deps[k];
            unsigned ns = 32;
            while (ld_acquire(f) == 0) { __nanosleep(ns); if (ns < 1024) ns <<= 1; }
        }
        __syncwarp();
        intra_op<pixel, 0>(a, d, sm, lane);
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release(flags + idx, 1u);
    }
}

// Dataflow launch for the TAIL of a group's wavefront (multi-frame graphs): the operations of the
// last levels of all frames in level order, dependencies as indices into that array.  A wait is
// bounded (about a second): a scheduling bug must not hang the GPU; it raises flags[-1 + 0] = sync[0]'s
// top bit instead, which the host never expects to see.
struct IntraFlowMultiArgs {
    const IntraFrameParams *frames;
    const Dav1dCudaIntraDesc *items;     // descriptor copies, pad = frame
    const int32_t *dep_start, *deps;
    unsigned *sync;                      // [0] claim counter, [1 + i] completion flag of operation i
    int n;
};

template <typename pixel>
__global__ void __launch_bounds__(INTRA_WARPS * 32, 4) intra_flow_multi_kernel(const __grid_constant__ IntraFlowMultiArgs m) {
    extern __shared__ __align__(16) uint8_t intra_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    IntraSmem<pixel> *sm = (IntraSmem<pixel> *)intra_smem_raw + warp;
    unsigned *counter = m.sync, *flags = m.sync + 1;
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = (int)(atomicAdd(counter, 1u) & 0x7fffffffu);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= m.n) break;
        const Dav1dCudaIntraDesc d = m.items[idx];
        const IntraFrameParams &a = m.frames[d.pad];
        const int d0 = m.dep_start[idx], d1 = m.dep_start[idx + 1];
        for (int k = d0 + lane; k < d1; k += 32) {
            const unsigned *f = flags + m.deps[k];
            unsigned ns = 32, waited = 0;
            while (ld_acquire(f) == 0) {
                __nanosleep(ns);
                waited += ns;
                if (ns < 1024) ns <<= 1;
                if (waited > (1u << 30)) { atomicOr(counter, 0x80000000u); break; }
            }
        }
        __syncwarp();
        intra_op<pixel, 0>(a, d, sm, lane);
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release(flags + idx, 1u);
    }
}

template <typename pixel, int CLS>
static int launch_intra_level_cls(const IntraArgs &a, cudaStream_t st) {
    const int grid = (a.n + INTRA_WARPS - 1) / INTRA_WARPS;
    launch_pdl(intra_level_kernel<pixel, CLS>, grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraSmem<pixel, CLS>), st, a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "intra_level_kernel") ? 0 : -5;
}
template <typename pixel>
static int launch_intra_level(const IntraArgs &a, cudaStream_t st, const int cls = 0) {
    switch (cls) {
    case 1: return launch_intra_level_cls<pixel, 1>(a, st);
    case 2: return launch_intra_level_cls<pixel, 2>(a, st);
    case 3: return launch_intra_level_cls<pixel, 3>(a, st);
    case 4: return launch_intra_level_cls<pixel, 4>(a, st);
    default: return launch_intra_level_cls<pixel, 0>(a, st);
    }
}

static int g_flow_blocks[2] = { 0, 0 };   // persistent grid size per pixel type (set in recon_init_attrs)

static int intra_flow_launch(const PicView &pic, int bw4, int bh4, void *cf, const Dav1dCudaIntraDesc *descs, int n,
                             const int32_t *dep_start, const int32_t *deps, unsigned *sync, const void *pal,
                             const uint8_t *pal_idx, cudaStream_t st)
{
    if (n <= 0) return 0;
    IntraArgs a;
    a.pic = pic; a.bw4 = bw4; a.bh4 = bh4; a.cf = cf;
    a.descs = descs; a.n = n; a.pal = pal; a.pal_idx = pal_idx;
    a.dep_start = dep_start; a.deps = deps; a.sync = sync; a.opw = 1;
    if (!cuda_ok(cudaMemsetAsync(sync, 0, (size_t)(n + 1) * sizeof(unsigned), st), "memset(intra sync)")) return -5;
    const bool hbd = pic.bdmax > 0xff;
    // experiment knob: D1_FLOW_BLOCKS caps the persistent grid (share of the GPU per stream)
    static const int cap = getenv("D1_FLOW_BLOCKS") ? atoi(getenv("D1_FLOW_BLOCKS")) : 1 << 30;
    const int grid = std::min(std::min(g_flow_blocks[hbd], cap), (n + INTRA_WARPS - 1) / INTRA_WARPS);
    if (hbd) intra_flow_kernel<uint16_t><<<grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraSmem<uint16_t>), st>>>(a);
    else intra_flow_kernel<uint8_t><<<grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraSmem<uint8_t>), st>>>(a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "intra_flow_kernel") ? 0 : -5;
}

static bool ensure_aux(Dav1dCudaContext *c);
static bool fork_aux(Dav1dCudaContext *c, cudaStream_t st);
static bool join_aux(Dav1dCudaContext *c, cudaStream_t st);

// Per level the three size classes run as parallel launches (class_start holds
// 3 * n_levels + 1 offsets into the level- and class-sorted descriptor array).
static int intra_batch_launch_classes(Dav1dCudaContext *c, const PicView &pic, int bw4, int bh4, void *cf,
                                      const Dav1dCudaIntraDesc *descs, const int32_t *class_start, int n_levels,
                                      const void *pal, const uint8_t *pal_idx, cudaStream_t st)
{
    if (!ensure_aux(c)) return -5;
    for (int l = 0; l < n_levels; l++) {
        int nz = 0;
        for (int k = 0; k < 3; k++) nz += class_start[3 * l + k + 1] > class_start[3 * l + k];
        const bool par = nz > 1;
        if (par && !fork_aux(c, st)) return -5;
        int used = 0;
        for (int k = 0; k < 3; k++) {
            const int n = class_start[3 * l + k + 1] - class_start[3 * l + k];
            if (n <= 0) continue;
            IntraArgs a;
            a.pic = pic; a.bw4 = bw4; a.bh4 = bh4; a.cf = cf;
            a.descs = descs + class_start[3 * l + k];
            a.n = n;
            a.pal = pal; a.pal_idx = pal_idx;
            a.dep_start = nullptr; a.deps = nullptr; a.sync = nullptr;
            cudaStream_t s = used == 0 ? st : c->aux[used - 1];
            used++;
            const int r = pic.bdmax > 0xff ? launch_intra_level<uint16_t>(a, s, k + 1)
                                           : launch_intra_level<uint8_t>(a, s, k + 1);
            if (r) return r;
        }
        if (par && !join_aux(c, st)) return -5;
    }
    return 0;
}

// Fused task variant (default): per level ONE launch (two when operations larger than 16x16
// exist).  A task = up to 32/G consecutive level-sorted operations whose residuals have the
// same transform size: the warp predicts them one after the other (whole warp per block), then
// runs all their residuals at once in groups of G lanes like itx_task_kernel; an operation
// without residual is a task of its own (tx code 31).
struct IntraTaskArgs {
    IntraArgs a;
    const uint32_t *tasks;
    int n_tasks;
};

template <typename pixel, bool BIG> struct IntraTaskSmem {
    IntraSmem<pixel, 4> pred;
    int tiles[BIG ? INTRA_TILE_INTS : 2 * 16 * 17];
};

template <typename pixel, int W, int H>
DEV void intra_task_residual(const IntraArgs &a, const int first, const int cnt, int *tiles, const int lane) {
    typedef ItxGeom<W, H> Geo;
    typedef typename PxTraits<pixel>::coef coef;
    constexpr int G = Geo::GMIN;
    const int grp = lane / G, gl = lane % G;
    const bool active = grp < cnt;
    Dav1dCudaIntraDesc d;
    if (active) d = a.descs[first + grp];
    else { d.coef_off = 0; d.x4 = d.y4 = 0; d.eob = 0; d.plane = 0; d.txtp = 0; d.cw4 = d.ch4 = 0; }
    const PlaneView &pv = a.pic.p[d.plane];
    const int stride = (int)(pv.stride / (int)sizeof(pixel));
    pixel *dst = (pixel *)pv.data + (int64_t)d.y4 * 4 * stride + d.x4 * 4;
    itx_block<pixel, W, H, G>(active, gl, tiles + grp * Geo::TILE_INTS, (coef *)a.cf + d.coef_off, d.eob, d.txtp,
                              dst, stride, a.pic.bdmax, false, d.cw4, d.ch4);
}

template <typename pixel, bool BIG>
__global__ void __launch_bounds__(INTRA_WARPS * 32, BIG ? 4 : 8)
intra_task_kernel(const __grid_constant__ IntraTaskArgs m) {
    extern __shared__ __align__(16) uint8_t intra_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * INTRA_WARPS + warp;
    if (t >= m.n_tasks) return;
    IntraTaskSmem<pixel, BIG> *sm = (IntraTaskSmem<pixel, BIG> *)intra_smem_raw + warp;
    const uint32_t code = m.tasks[t];
    const int first = (int)(code >> 8), tx = (code >> 3) & 31, cnt = (int)(code & 7) + 1;
    for (int k = 0; k < cnt; k++) {
        const Dav1dCudaIntraDesc d = m.a.descs[first + k];
        intra_op<pixel, 4>(m.a, d, &sm->pred, lane);
        __syncwarp();
    }
    if (tx == 31) return;
#define D1_ITASK(T, W, H) \
    case T: \
        if constexpr (BIG == (W > 16 || H > 16)) intra_task_residual<pixel, W, H>(m.a, first, cnt, sm->tiles, lane); \
        break;
    switch (tx) {
    D1_ITASK(0, 4, 4) D1_ITASK(1, 8, 8) D1_ITASK(2, 16, 16) D1_ITASK(3, 32, 32) D1_ITASK(4, 64, 64)
    D1_ITASK(5, 4, 8) D1_ITASK(6, 8, 4) D1_ITASK(7, 8, 16) D1_ITASK(8, 16, 8) D1_ITASK(9, 16, 32)
    D1_ITASK(10, 32, 16) D1_ITASK(11, 32, 64) D1_ITASK(12, 64, 32) D1_ITASK(13, 4, 16) D1_ITASK(14, 16, 4)
    D1_ITASK(15, 8, 32) D1_ITASK(16, 32, 8) D1_ITASK(17, 16, 64) D1_ITASK(18, 64, 16)
    default: break;
    }
#undef D1_ITASK
}

static int intra_task_launch(const PicView &pic, int bw4, int bh4, void *cf, const Dav1dCudaIntraDesc *descs,
                             const void *pal, const uint8_t *pal_idx, const uint32_t *tasks,
                             const int32_t *task_start, int n_levels, cudaStream_t st)
{
    IntraTaskArgs m;
    m.a.pic = pic; m.a.bw4 = bw4; m.a.bh4 = bh4; m.a.cf = cf; m.a.descs = descs; m.a.n = 0;
    m.a.pal = pal; m.a.pal_idx = pal_idx;
    m.a.dep_start = nullptr; m.a.deps = nullptr; m.a.sync = nullptr; m.a.opw = 1;
    const bool hbd = pic.bdmax > 0xff;
    for (int l = 0; l < n_levels; l++) {
        const int ns = task_start[2 * l + 1] - task_start[2 * l], nb = task_start[2 * l + 2] - task_start[2 * l + 1];
        if (ns > 0) {
            m.tasks = tasks + task_start[2 * l]; m.n_tasks = ns;
            const int grid = (ns + INTRA_WARPS - 1) / INTRA_WARPS;
            if (hbd) intra_task_kernel<uint16_t, false><<<grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraTaskSmem<uint16_t, false>), st>>>(m);
            else intra_task_kernel<uint8_t, false><<<grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraTaskSmem<uint8_t, false>), st>>>(m);
            count_launch();
        }
        if (nb > 0) {
            m.tasks = tasks + task_start[2 * l + 1]; m.n_tasks = nb;
            const int grid = (nb + INTRA_WARPS - 1) / INTRA_WARPS;
            if (hbd) intra_task_kernel<uint16_t, true><<<grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraTaskSmem<uint16_t, true>), st>>>(m);
            else intra_task_kernel<uint8_t, true><<<grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraTaskSmem<uint8_t, true>), st>>>(m);
            count_launch();
        }
    }
    return cuda_ok(cudaGetLastError(), "intra_task_kernel") ? 0 : -5;
}

// Split variant: per level a prediction-only launch followed by the level's residuals as
// transform tasks (small sizes, then the rare large ones).
static int intra_batch_launch_split(const PicView &pic, int bw4, int bh4, void *cf, const Dav1dCudaIntraDesc *descs,
                                    const int32_t *level_start, int n_levels, const void *pal,
                                    const uint8_t *pal_idx, const Dav1dCudaItxDesc *itx, const uint32_t *tasks,
                                    const int32_t *task_start, cudaStream_t st)
{
    for (int l = 0; l < n_levels; l++) {
        const int n = level_start[l + 1] - level_start[l];
        if (n <= 0) continue;
        IntraArgs a;
        a.pic = pic; a.bw4 = bw4; a.bh4 = bh4; a.cf = cf;
        a.descs = descs + level_start[l];
        a.n = n;
        a.pal = pal; a.pal_idx = pal_idx;
        a.dep_start = nullptr; a.deps = nullptr; a.sync = nullptr; a.opw = 1;
        static const int part = getenv("D1_INTRA_PART") ? atoi(getenv("D1_INTRA_PART")) : 3;   // experiment knob
        // Levels with few operations (the long tail of the wavefront) are latency-bound: one
        // fused launch (prediction + residual by the same warp) instead of up to three.
        static const int fuse_below = getenv("D1_INTRA_FUSE_BELOW") ? atoi(getenv("D1_INTRA_FUSE_BELOW")) : 512;
        int r = 0;
        if (n < fuse_below) {
            r = pic.bdmax > 0xff ? launch_intra_level<uint16_t>(a, st, 0) : launch_intra_level<uint8_t>(a, st, 0);
            if (r) return r;
            continue;
        }
        if (part & 1) r = pic.bdmax > 0xff ? launch_intra_level<uint16_t>(a, st, 4) : launch_intra_level<uint8_t>(a, st, 4);
        if (r) return r;
        const int ns = task_start[2 * l + 1] - task_start[2 * l], nb = task_start[2 * l + 2] - task_start[2 * l + 1];
        if ((part & 2) && ns + nb > 0 &&
            (r = itx_task_launch(pic, cf, itx, tasks + task_start[2 * l], ns, nb, 0, st, st))) return r;
    }
    return 0;
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
        a.dep_start = nullptr; a.deps = nullptr; a.sync = nullptr;
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

// One frame.  Launch classes that touch disjoint pixels run as parallel
// branches: {put} | {compound wave 0 -> wave 1} | {warp}, then the 19 transform
// size classes spread over 4 streams, then the intra phase.
static int recon_submit_on(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b, cudaStream_t st,
                           const int phase_mask = 31) {
    const PicView dst = pic_view(b->dst);
    PicView refs[7];
    refs_view(refs, b->refs);
    if (!ensure_aux(c)) return -5;
    int r;
    // experiment knob (tools/exp_frame.py): D1_PHASE_MASK selects launch classes
    // bit0 put, bit1 compound, bit2 warp, bit3 itx, bit4 intra; default all
    static const int env_mask = getenv("D1_PHASE_MASK") ? atoi(getenv("D1_PHASE_MASK")) : 31;
    const int mask = env_mask & phase_mask;
    // phase A: prediction from reference frames
    if (!fork_aux(c, st)) return -5;
    if ((mask & 1) && (r = mc_put_launch_raw(dst, refs, b->mc_put, b->mc_put_tiles, b->n_mc_put_tiles, b->n_mc_put_small, nullptr,
                                             nullptr, false, st)))
        return r;
    if ((mask & 2) && (r = mc_put_launch_raw(dst, refs, b->mc_comp, b->mc_comp_tiles, b->n_mc_comp_tiles[0],
                                             b->n_mc_comp_small[0], b->masks, nullptr, true, c->aux[0])))
        return r;
    if ((mask & 2) && (r = mc_put_launch_raw(dst, refs, b->mc_comp, b->mc_comp_tiles + b->n_mc_comp_tiles[0],
                                             b->n_mc_comp_tiles[1], b->n_mc_comp_small[1], b->masks, nullptr, true, c->aux[0])))
        return r;
    if ((mask & 4) && (r = warp_batch_launch(dst, refs, b->warp, b->n_warp, c->aux[1]))) return r;
    if (!join_aux(c, st)) return -5;
    // OBMC blends onto the finished predictions: top neighbours, then left neighbours
    if ((mask & 1) && b->mc_obmc) {
        if ((r = mc_obmc_launch_raw(dst, refs, b->mc_obmc, b->mc_obmc_tiles, b->n_mc_obmc_tiles[0], st))) return r;
        if ((r = mc_obmc_launch_raw(dst, refs, b->mc_obmc, b->mc_obmc_tiles + b->n_mc_obmc_tiles[0],
                                    b->n_mc_obmc_tiles[1], st))) return r;
    }
    // phase B: inter residuals
    if (b->itx && b->itx_tasks && (mask & 8)) {
        if (!fork_aux(c, st)) return -5;
        if ((r = itx_task_launch(dst, b->cf, b->itx, b->itx_tasks, b->n_itx_tasks[0], b->n_itx_tasks[1], 0, st,
                                 c->aux[0])))
            return r;
        if (!join_aux(c, st)) return -5;
    } else if (b->itx && (mask & 8)) {
        if (!fork_aux(c, st)) return -5;
        cudaStream_t ss[1 + Dav1dCudaContext::N_AUX] = { st, c->aux[0], c->aux[1], c->aux[2] };
        if ((r = itx_batch_launch_multi(dst, b->cf, b->itx, b->itx_class_count, 0, ss, 1 + Dav1dCudaContext::N_AUX)))
            return r;
        if (!join_aux(c, st)) return -5;
    }
    // phase C: intra-class operations
    if (!(mask & 16)) return 0;
    if (b->intra && b->intra_deps && b->intra_sync) {
        const int n = b->intra_level_start[b->n_levels];
        if ((r = intra_flow_launch(dst, b->bw4, b->bh4, b->cf, b->intra, n, b->intra_dep_start, b->intra_deps,
                                   (unsigned *)b->intra_sync, b->pal, b->pal_idx, st)))
            return r;
    } else if (b->intra && b->intra_tasks && b->intra_task_start) {
        if ((r = intra_task_launch(dst, b->bw4, b->bh4, b->cf, b->intra, b->pal, b->pal_idx, b->intra_tasks,
                                   b->intra_task_start, b->n_levels, st)))
            return r;
    } else if (b->intra && b->intra_itx && b->intra_itx_tasks && b->intra_itx_task_start) {
        if ((r = intra_batch_launch_split(dst, b->bw4, b->bh4, b->cf, b->intra, b->intra_level_start, b->n_levels,
                                          b->pal, b->pal_idx, b->intra_itx, b->intra_itx_tasks,
                                          b->intra_itx_task_start, st)))
            return r;
    } else if (b->intra && b->intra_class_start) {
        if ((r = intra_batch_launch_classes(c, dst, b->bw4, b->bh4, b->cf, b->intra, b->intra_class_start, b->n_levels,
                                            b->pal, b->pal_idx, st)))
            return r;
    } else if (b->intra && (r = intra_batch_launch(dst, b->bw4, b->bh4, b->cf, b->intra, b->intra_level_start,
                                                   b->n_levels, b->pal, b->pal_idx, st)))
        return r;
    return 0;
}

// Frames of several independent streams in one submission.  Phases A and B are
// launched per frame (spread over the fork/join streams); phase C runs one
// launch per dependency level covering that level of every frame.
// `tab` = device scratch for the per-frame parameter table + segment tables
// (uploaded here from `tab_host`, which must stay valid until the copy ran).
struct MultiTables {
    std::vector<IntraFrameParams> frames;
    std::vector<Dav1dCudaIntraDesc> items;   // all levels, concatenated: descriptor copies, pad = frame
    std::vector<int> level_start;        // per level: first item; size n_levels + 1
    // split execution (frames that carry intra_itx): residual tasks (code, frame) of all
    // frames per level, small sizes then large ones
    bool split = false;
    std::vector<ItxFrameRef> itx_frames;
    std::vector<Dav1dCudaItxDesc> ritx;  // residual descriptors of all frames, concatenated (copies)
    std::vector<uint2> rtasks;           // code (first index into ritx) , frame
    // tail of the wavefront as one dataflow launch: levels [tail_level, n_levels) = items
    // [level_start[tail_level], end) with dependency lists relative to that range
    int tail_level = -1;
    std::vector<int32_t> tail_dep_start, tail_deps;
    std::vector<int> rtask_start;        // 2 * n_levels + 1
};

static uint64_t intra_code_key(const Dav1dCudaIntraDesc &d) {
    const uint64_t res = d.eob >= 0 ? 1 + d.tx : 0;
    const uint64_t cls = (uint64_t)intra_size_class(d.tw4 * 4, d.th4 * 4);
    return (cls << 28) | (res << 16) | ((uint64_t)d.mode << 8) | (d.eob >= 0 ? d.txtp : 0);
}

// Level l of every frame merged and sorted by code path (needs the host copy of
// each frame's sorted descriptors, Dav1dCudaReconBatch.intra_host).
static int build_multi_tables(const Dav1dCudaReconBatch *const *bs, int n, MultiTables &t) {
    int max_levels = 0;
    t.frames.resize(n);
    for (int f = 0; f < n; f++) {
        const Dav1dCudaReconBatch *b = bs[f];
        IntraFrameParams &p = t.frames[f];
        p.pic = pic_view(b->dst); p.bw4 = b->bw4; p.bh4 = b->bh4; p.cf = b->cf;
        p.descs = b->intra; p.n = 0; p.pal = b->pal; p.pal_idx = b->pal_idx;
        p.dep_start = nullptr; p.deps = nullptr; p.sync = nullptr; p.opw = 1;
        if (b->intra && b->n_levels > 0) {
            if (!b->intra_host) return -22;
            max_levels = std::max(max_levels, (int)b->n_levels);
        }
    }
    t.level_start.assign(1, 0);
    std::vector<int32_t> item_src;           // per item: sorted index inside its frame
    std::vector<std::pair<uint64_t, uint32_t>> lv;
    for (int l = 0; l < max_levels; l++) {
        lv.clear();
        for (int f = 0; f < n; f++) {
            const Dav1dCudaReconBatch *b = bs[f];
            if (!b->intra || !b->intra_host || l >= b->n_levels) continue;
            for (int i = b->intra_level_start[l]; i < b->intra_level_start[l + 1]; i++)
                lv.push_back({ intra_code_key(b->intra_host[i]), ((uint32_t)f << 24) | (uint32_t)i });
        }
        std::stable_sort(lv.begin(), lv.end(),
                         [](const std::pair<uint64_t, uint32_t> &x, const std::pair<uint64_t, uint32_t> &y) {
                             return x.first < y.first;
                         });
        for (auto &e : lv) {
            Dav1dCudaIntraDesc d = bs[e.second >> 24]->intra_host[e.second & 0xffffff];
            d.pad = (uint16_t)(e.second >> 24);
            t.items.push_back(d);
            item_src.push_back((int32_t)(e.second & 0xffffff));
        }
        t.level_start.push_back((int)t.items.size());
    }
    // ---- tail: the last levels, each below the fusing threshold, when every frame brings its
    // dependency lists (experiment knob D1_INTRA_TAIL=0 keeps one fused launch per level)
    {
        static const int fuse_below = getenv("D1_INTRA_FUSE_BELOW") ? atoi(getenv("D1_INTRA_FUSE_BELOW")) : 512;
        static const bool tail_on = !getenv("D1_INTRA_TAIL") || atoi(getenv("D1_INTRA_TAIL")) != 0;
        bool have = tail_on && max_levels > 0;
        for (int f = 0; f < n; f++)
            if (bs[f]->intra && bs[f]->n_levels > 0 && (!bs[f]->intra_dep_start_host || !bs[f]->intra_deps_host)) have = false;
        int tl = max_levels;
        while (have && tl > 0 && t.level_start[tl] - t.level_start[tl - 1] < fuse_below) tl--;
        if (have && max_levels - tl >= 3) {            // worth it from three launches on
            const int t0 = t.level_start[tl], t1 = (int)t.items.size();
            // (frame, sorted index inside the frame) -> index inside the tail
            std::vector<std::vector<int32_t>> pos(n);
            for (int f = 0; f < n; f++)
                if (bs[f]->intra && bs[f]->n_levels > 0) pos[f].assign(bs[f]->intra_level_start[bs[f]->n_levels], -1);
            for (int j = t0; j < t1; j++) pos[t.items[j].pad][item_src[j]] = j - t0;
            t.tail_dep_start.assign(1, 0);
            for (int j = t0; j < t1; j++) {
                const int f = t.items[j].pad;
                const int32_t i = item_src[j];
                const int32_t *ds = bs[f]->intra_dep_start_host, *dd = bs[f]->intra_deps_host;
                for (int k = ds[i]; k < ds[i + 1]; k++) {
                    const int32_t q = pos[f][dd[k]];
                    if (q < 0) continue;               // produced by an earlier launch of this graph
                    if (q >= j - t0) return -22;       // would wait for a later operation: scheduler bug
                    t.tail_deps.push_back(q);
                }
                t.tail_dep_start.push_back((int32_t)t.tail_deps.size());
            }
            t.tail_level = tl;
        }
    }
    // residual tasks: regenerated on the host from each frame's sorted descriptors (the same
    // deterministic routine that produced the frame's device arrays intra_itx / intra_itx_tasks)
    t.split = n > 0;
    for (int f = 0; f < n; f++)
        if (bs[f]->intra && bs[f]->n_levels > 0 && !bs[f]->intra_itx) t.split = false;
    if (!t.split) return 0;
    t.itx_frames.resize(n);
    struct Key { uint32_t key; uint2 tk; };
    std::vector<std::vector<Key>> per_level(2 * (size_t)max_levels);
    for (int f = 0; f < n; f++) {
        const Dav1dCudaReconBatch *b = bs[f];
        t.itx_frames[f].pic = pic_view(b->dst);
        t.itx_frames[f].cf = b->cf;
        t.itx_frames[f].descs = b->intra_itx;
        if (!b->intra || b->n_levels <= 0) continue;
        const int n_ops = b->intra_level_start[b->n_levels];
        std::vector<Dav1dCudaItxDesc> itx((size_t)std::max(n_ops, 1));
        std::vector<uint32_t> tasks((size_t)std::max(n_ops, 1));
        std::vector<int32_t> tstart(2 * (size_t)b->n_levels + 1);
        int32_t nt = 0;
        const int n_itx = dav1d_cuda_intra_residual_tasks(b->intra_host, b->intra_level_start, b->n_levels, itx.data(),
                                                          tasks.data(), tstart.data(), &nt);
        if (n_itx < 0) return -22;
        const uint32_t off = (uint32_t)t.ritx.size();
        if (off + (uint32_t)n_itx >= (1u << 24)) return -22;        // task codes hold 24-bit indices
        t.ritx.insert(t.ritx.end(), itx.begin(), itx.begin() + n_itx);
        for (int l = 0; l < b->n_levels; l++)
            for (int half = 0; half < 2; half++)
                for (int k = tstart[2 * l + half]; k < tstart[2 * l + half + 1]; k++) {
                    const Dav1dCudaItxDesc &d0 = itx[tasks[k] >> 8];
                    const uint32_t tp = d0.eob == 0 && d0.txtp == 0 ? 0 : 1 + d0.txtp;
                    const uint32_t code = (((tasks[k] >> 8) + off) << 8) | (tasks[k] & 0xff);
                    per_level[2 * (size_t)l + half].push_back(
                        { (((tasks[k] >> 3) & 31) << 8) | tp, make_uint2(code, (unsigned)f) });
                }
    }
    t.rtask_start.assign(1, 0);
    for (auto &v : per_level) {
        std::stable_sort(v.begin(), v.end(), [](const Key &x, const Key &y) { return x.key < y.key; });
        for (auto &e : v) t.rtasks.push_back(e.tk);
        t.rtask_start.push_back((int)t.rtasks.size());
    }
    // TIMING-ONLY experiment knob (results are WRONG: dependencies are ignored): D1_INTRA_FLAT=1 runs
    // all intra-class operations of the group as ONE level = the throughput floor of the kernels
    // without the wavefront's launch chain (tools/exp_frame.py)
    static const bool flat = getenv("D1_INTRA_FLAT") && atoi(getenv("D1_INTRA_FLAT")) != 0;
    if (flat && max_levels > 1) {
        std::stable_sort(t.items.begin(), t.items.end(), [](const Dav1dCudaIntraDesc &x, const Dav1dCudaIntraDesc &y) {
            return intra_code_key(x) < intra_code_key(y);
        });
        t.level_start = { 0, (int)t.items.size() };
        t.tail_level = -1;
        std::vector<Key> all[2];
        for (size_t i = 0; i < per_level.size(); i++)
            all[i & 1].insert(all[i & 1].end(), per_level[i].begin(), per_level[i].end());
        t.rtasks.clear();
        t.rtask_start.assign(1, 0);
        for (auto &v : all) {
            std::stable_sort(v.begin(), v.end(), [](const Key &x, const Key &y) { return x.key < y.key; });
            for (auto &e : v) t.rtasks.push_back(e.tk);
            t.rtask_start.push_back((int)t.rtasks.size());
        }
    }
    return 0;
}

template <typename pixel, int CLS>
static int launch_intra_multi(const IntraMultiArgs &m, cudaStream_t st) {
    const int grid = (m.n + INTRA_WARPS - 1) / INTRA_WARPS;
    intra_multi_kernel<pixel, CLS><<<grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraSmem<pixel, CLS>), st>>>(m);
    count_launch();
    return cuda_ok(cudaGetLastError(), "intra_multi_kernel") ? 0 : -5;
}

static int recon_submit_multi_on(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n,
                                 const IntraFrameParams *d_frames, const Dav1dCudaIntraDesc *d_items,
                                 const ItxFrameRef *d_itx_frames, const Dav1dCudaItxDesc *d_ritx, const uint2 *d_rtasks,
                                 const int32_t *d_tail_dep_start, const int32_t *d_tail_deps, unsigned *d_tail_sync,
                                 const MultiTables &t,
                                 cudaStream_t st, const int phase_mask)
{
    if (!ensure_aux(c)) return -5;
    int r;
    static const int env_mask = getenv("D1_PHASE_MASK") ? atoi(getenv("D1_PHASE_MASK")) : 31;
    const int mask = env_mask & phase_mask;
    cudaStream_t ss[1 + Dav1dCudaContext::N_AUX] = { st, c->aux[0], c->aux[1], c->aux[2] };
    constexpr int NS = 1 + Dav1dCudaContext::N_AUX;
    // phases A + B per frame, frame f on stream f % NS (MC then residual of a frame stay ordered)
    if (!fork_aux(c, st)) return -5;
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
    if (!(mask & 16)) return 0;
    // phase C: per level one set of launches over all frames: prediction, then the residual
    // tasks (small / large sizes); levels with few operations run fused (one launch)
    const bool hbd = bs[0]->dst->bitdepth_max > 0xff;
    static const int fuse_below = getenv("D1_INTRA_FUSE_BELOW") ? atoi(getenv("D1_INTRA_FUSE_BELOW")) : 512;
    const size_t n_level_launches = t.tail_level >= 0 ? (size_t)t.tail_level : t.level_start.size() - 1;
    for (size_t l = 0; l < n_level_launches; l++) {
        const int s0 = t.level_start[l], s1 = t.level_start[l + 1];
        if (s1 <= s0) continue;
        IntraMultiArgs m;
        m.frames = d_frames; m.items = d_items + s0; m.n = s1 - s0;
        if (!t.split || m.n < fuse_below) {
            if ((r = hbd ? launch_intra_multi<uint16_t, 0>(m, st) : launch_intra_multi<uint8_t, 0>(m, st))) return r;
            continue;
        }
        if ((r = hbd ? launch_intra_multi<uint16_t, 4>(m, st) : launch_intra_multi<uint8_t, 4>(m, st))) return r;
        const int a0 = t.rtask_start[2 * l], a1 = t.rtask_start[2 * l + 1], a2 = t.rtask_start[2 * l + 2];
        if (a2 > a0) {
            // the level's small and large transform tasks touch disjoint blocks: parallel branches
            const bool par = a1 > a0 && a2 > a1;
            if (par && !fork_aux(c, st)) return -5;
            if ((r = itx_multi_task_launch(d_itx_frames, d_ritx, d_rtasks + a0, a1 - a0, a2 - a1, hbd, st,
                                           par ? c->aux[0] : st))) return r;
            if (par && !join_aux(c, st)) return -5;
        }
    }
    if (t.tail_level >= 0) {
        // the tail of the wavefront: one dataflow launch for all remaining levels of the group
        IntraFlowMultiArgs fm;
        fm.frames = d_frames;
        fm.items = d_items + t.level_start[t.tail_level];
        fm.n = (int)t.items.size() - t.level_start[t.tail_level];
        fm.dep_start = d_tail_dep_start; fm.deps = d_tail_deps; fm.sync = d_tail_sync;
        if (!cuda_ok(cudaMemsetAsync(d_tail_sync, 0, (size_t)(fm.n + 1) * sizeof(unsigned), st), "memset(tail sync)"))
            return -5;
        const int grid = std::min(g_flow_blocks[hbd], (fm.n + INTRA_WARPS - 1) / INTRA_WARPS);
        if (hbd) intra_flow_multi_kernel<uint16_t><<<grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraSmem<uint16_t>), st>>>(fm);
        else intra_flow_multi_kernel<uint8_t><<<grid, INTRA_WARPS * 32, INTRA_WARPS * sizeof(IntraSmem<uint8_t>), st>>>(fm);
        count_launch();
        if (!cuda_ok(cudaGetLastError(), "intra_flow_multi_kernel")) return -5;
    }
    return 0;
}

void recon_init_attrs() {
    itx_init_attrs();
    cudaFuncSetAttribute(intra_task_kernel<uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraTaskSmem<uint16_t, true>)));
    cudaFuncSetAttribute(intra_task_kernel<uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraTaskSmem<uint8_t, true>)));
    cudaFuncSetAttribute(intra_multi_kernel<uint16_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint16_t>)));
    cudaFuncSetAttribute(intra_multi_kernel<uint8_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint8_t>)));
    cudaFuncSetAttribute(intra_flow_multi_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint16_t>)));
    cudaFuncSetAttribute(intra_flow_multi_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint8_t>)));
    cudaFuncSetAttribute(intra_flow_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint16_t>)));
    cudaFuncSetAttribute(intra_flow_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint8_t>)));
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, intra_flow_kernel<uint8_t>, INTRA_WARPS * 32,
                                                  INTRA_WARPS * sizeof(IntraSmem<uint8_t>));
    g_flow_blocks[0] = std::max(1, occ) * std::max(1, sms);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, intra_flow_kernel<uint16_t>, INTRA_WARPS * 32,
                                                  INTRA_WARPS * sizeof(IntraSmem<uint16_t>));
    g_flow_blocks[1] = std::max(1, occ) * std::max(1, sms);
    cudaFuncSetAttribute(intra_level_kernel<uint16_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint16_t, 0>)));
    cudaFuncSetAttribute(intra_level_kernel<uint8_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint8_t, 0>)));
    cudaFuncSetAttribute(intra_level_kernel<uint16_t, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint16_t, 3>)));
    cudaFuncSetAttribute(intra_level_kernel<uint8_t, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INTRA_WARPS * sizeof(IntraSmem<uint8_t, 3>)));
}

}  // namespace d1

using namespace d1;

struct Dav1dCudaReconGraph {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    int n_nodes;
    void *tables;          // device: multi-frame parameter/segment tables (multi graphs only)
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
// Edges the resolved predictor needs: bit0 left, bit1 top, bit2 topleft, bit3 topright,
// bit4 bottomleft (host mirror of the table in csrc/ipred.cuh prepare_edges()).
static int intra_needs(int mode, int angle_delta, int have_left, int have_top) {
    if (mode >= 1 && mode <= 8) {
        static const int base[8] = { 90, 180, 45, 135, 113, 157, 203, 67 };
        const int a = base[mode - 1] + 3 * angle_delta;
        if (a <= 90) return (a < 90 && have_top) ? (2 | 8 | 4) : 2;             // Z1 : VERT
        if (a < 180) return 1 | 2 | 4;                                          // Z2
        return (a > 180 && have_left) ? (1 | 16 | 4) : 1;                       // Z3 : HOR
    }
    if (mode == 0) return have_left ? (have_top ? 3 : 1) : (have_top ? 2 : 0);  // DC family
    if (mode == 12) return have_left ? (have_top ? 7 : 1) : (have_top ? 2 : 0); // PAETH -> HOR / VERT / DC_128
    if (mode >= 9 && mode <= 11) return 3;                                      // SMOOTH*
    return 1 | 2 | 4;                                                           // FILTER
}

int dav1d_cuda_intra_schedule(Dav1dCudaIntraDesc *descs, int n, int bw4, int bh4, int ss_hor, int ss_ver,
                              int32_t *order, int32_t *level_start, int max_levels)
{
    return dav1d_cuda_intra_schedule_deps(descs, n, bw4, bh4, ss_hor, ss_ver, order, level_start, max_levels,
                                          nullptr, nullptr, 0, nullptr);
}

int dav1d_cuda_intra_schedule_deps(Dav1dCudaIntraDesc *descs, int n, int bw4, int bh4, int ss_hor, int ss_ver,
                                   int32_t *order, int32_t *level_start, int max_levels,
                                   int32_t *dep_start, int32_t *deps, int max_deps, int32_t *class_start)
{
    if (!descs || n < 0 || !order || !level_start) return -22;
    const int pw[3] = { bw4, (bw4 + ss_hor) >> ss_hor, (bw4 + ss_hor) >> ss_hor };
    const int ph[3] = { bh4, (bh4 + ss_ver) >> ss_ver, (bh4 + ss_ver) >> ss_ver };
    // per 4x4 cell: decode-order index of the operation that produced its final pixels (-1: inter phases)
    std::vector<int32_t> prod[3];
    for (int p = 0; p < 3; p++) prod[p].assign((size_t)pw[p] * ph[p], -1);
    std::vector<int32_t> dlist;              // dependency lists in decode order (CSR)
    std::vector<int32_t> dstart(n + 1, 0);
    std::vector<int32_t> cur;
    int n_levels = 0;
    for (int i = 0; i < n; i++) {
        Dav1dCudaIntraDesc &d = descs[i];
        const int p = d.plane, W = pw[p], H = ph[p];
        const int x0 = d.x4, y0 = d.y4, x1 = std::min<int>(x0 + d.tw4, W), y1 = std::min<int>(y0 + d.th4, H);
        cur.clear();
        auto dep = [&](int pl, int x, int y) {
            if (x >= 0 && y >= 0 && x < pw[pl] && y < ph[pl]) {
                const int32_t q = prod[pl][(size_t)y * pw[pl] + x];
                if (q >= 0 && (cur.empty() || cur.back() != q)) cur.push_back(q);
            }
        };
        if (d.mode == DAV1D_CUDA_INTRA_NONE) {
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) dep(p, x, y);
        } else if (d.mode == DAV1D_CUDA_INTRA_IBC) {
            // every operation that produced a pixel of the (clamped) source area
            const int sx = (int16_t)(d.aux & 0xffff), sy = (int16_t)(d.aux >> 16);
            const int px_w = 4 * W, px_h = 4 * H;
            const int xa = std::max(0, std::min(sx, px_w - 1)) >> 2;
            const int xb = std::max(0, std::min(sx + 4 * d.tw4 + (d.angle_delta ? 1 : 0) - 1, px_w - 1)) >> 2;
            const int ya = std::max(0, std::min(sy, px_h - 1)) >> 2;
            const int yb = std::max(0, std::min(sy + 4 * d.th4 + (d.flags ? 1 : 0) - 1, px_h - 1)) >> 2;
            for (int y = ya; y <= yb; y++)
                for (int x = xa; x <= xb; x++) dep(p, x, y);
        } else if (d.mode != DAV1D_CUDA_INTRA_PAL) {
            // exactly the pixels dav1d_prepare_intra_edges reads for the resolved mode
            // (ipred_prepare_tmpl.c:50-74 needs_* table, :94-117 mode resolution)
            const int have_left = x0 > d.tile_x4_start, have_top = y0 > d.tile_y4_start;
            const int needs = d.mode == DAV1D_CUDA_INTRA_II ? intra_needs(d.angle_delta, 0, have_left, have_top)
                              : intra_needs(d.mode == DAV1D_CUDA_INTRA_CFL ? 0 : d.mode, d.angle_delta, have_left,
                                            have_top);
            // bit0 left, bit1 top, bit2 topleft, bit3 topright, bit4 bottomleft
            const bool rd_top = have_top && ((needs & 2) || (needs & 4) || ((needs & 1) && !have_left));
            if (rd_top) {
                const bool tr = (needs & 8) && (d.edge_flags & 1);
                const int xs = ((needs & 4) && have_left) ? x0 - 1 : x0;
                const int xe = (needs & 2) ? std::min<int>(x0 + d.tw4 + (tr ? d.tw4 : 0), d.tile_x4_end) : x0 + 1;
                for (int x = xs; x < xe; x++) dep(p, x, y0 - 1);
            }
            const bool rd_left = have_left && ((needs & 1) || ((needs & 2) && !have_top) || ((needs & 4) && !have_top));
            if (rd_left) {
                const bool bl = (needs & 16) && (d.edge_flags & 8);
                const int ye = (needs & 1) ? std::min<int>(y0 + d.th4 + (bl ? d.th4 : 0), d.tile_y4_end) : y0 + 1;
                for (int y = y0; y < ye; y++) dep(p, x0 - 1, y);
            }
            if (d.mode == DAV1D_CUDA_INTRA_CFL) {
                const int sh = p ? ss_hor : 0, sv = p ? ss_ver : 0;
                for (int y = y0 << sv; y < ((y0 + d.th4) << sv); y++)
                    for (int x = x0 << sh; x < ((x0 + d.tw4) << sh); x++) dep(0, x, y);
            }
        }
        std::sort(cur.begin(), cur.end());
        cur.erase(std::unique(cur.begin(), cur.end()), cur.end());
        int lv = 0;
        for (int32_t q : cur) lv = std::max<int>(lv, (int)descs[q].level);
        lv += 1;
        d.level = (uint32_t)lv;
        n_levels = std::max(n_levels, lv);
        dlist.insert(dlist.end(), cur.begin(), cur.end());
        dstart[i + 1] = (int32_t)dlist.size();
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) prod[p][(size_t)y * W + x] = i;
    }
    if (n_levels > max_levels) return -34;
    std::vector<int32_t> cnt(n_levels + 2, 0);
    for (int i = 0; i < n; i++) cnt[descs[i].level]++;          // levels are 1-based
    level_start[0] = 0;
    for (int l = 1; l <= n_levels; l++) level_start[l] = level_start[l - 1] + cnt[l];
    // Within a level, group operations that run the same code (residual size, prediction
    // mode, transform type): the fused kernel is ~1.5 MB of SASS and co-resident warps that
    // execute different paths thrash the instruction cache (measured 4x on a 4K frame).
    std::vector<int32_t> inv(n);
    {
        std::vector<std::pair<uint64_t, int32_t>> keyed(n);
        for (int i = 0; i < n; i++) {
            const Dav1dCudaIntraDesc &d = descs[i];
            const uint64_t res = d.eob >= 0 ? 1 + d.tx : 0;
            const uint64_t cls = (uint64_t)intra_size_class(d.tw4 * 4, d.th4 * 4);
            const uint64_t key = ((uint64_t)d.level << 32) | (cls << 28) | (res << 16) | ((uint64_t)d.mode << 8) |
                                 (d.eob >= 0 ? d.txtp : 0);
            keyed[i] = { key, i };
        }
        std::stable_sort(keyed.begin(), keyed.end(),
                         [](const std::pair<uint64_t, int32_t> &a, const std::pair<uint64_t, int32_t> &b) {
                             return a.first < b.first;
                         });
        for (int s2 = 0; s2 < n; s2++) { order[s2] = keyed[s2].second; inv[keyed[s2].second] = s2; }
        if (class_start) {
            // offsets of (level, class) runs in the sorted order: 3 * n_levels + 1 entries
            int k = 0;
            for (int l = 1; l <= n_levels; l++)
                for (uint64_t c2 = 1; c2 <= 3; c2++) {
                    class_start[3 * (l - 1) + (int)c2 - 1] = k;
                    while (k < n && (keyed[k].first >> 32) == (uint64_t)l && ((keyed[k].first >> 28) & 15) == c2) k++;
                }
            class_start[3 * n_levels] = k;
        }
    }
    if (dep_start && deps) {
        if ((int)dlist.size() > max_deps) return -28;
        int k = 0;
        for (int s = 0; s < n; s++) {          // sorted order, indices translated to sorted space
            const int i = order[s];
            dep_start[s] = k;
            for (int j = dstart[i]; j < dstart[i + 1]; j++) deps[k++] = inv[dlist[j]];
        }
        dep_start[n] = k;
    }
    return n_levels;
}

// Task codes for the fused task kernel, see include/dav1d_cuda.h.
int dav1d_cuda_intra_tasks(const Dav1dCudaIntraDesc *sd, const int32_t *level_start, int n_levels,
                           uint32_t *tasks, int32_t *task_start, int32_t *n_tasks)
{
    if (!sd || !level_start || !tasks || !task_start || !n_tasks) return -22;
    static const uint8_t w4[19] = { 1, 2, 4, 8, 16, 1, 2, 2, 4, 4, 8, 8, 16, 1, 4, 2, 8, 4, 16 };
    static const uint8_t h4[19] = { 1, 2, 4, 8, 16, 2, 1, 4, 2, 8, 4, 16, 8, 4, 1, 8, 2, 16, 4 };
    int k = 0;
    for (int l = 0; l < n_levels; l++) {
        task_start[2 * l] = k;
        for (int pass = 0; pass < 2; pass++) {       // small operations first, then those larger than 16x16
            if (pass == 1) task_start[2 * l + 1] = k;
            int i = level_start[l];
            while (i < level_start[l + 1]) {
                const Dav1dCudaIntraDesc &d = sd[i];
                const bool big = d.tw4 > 4 || d.th4 > 4;
                const bool res = d.eob >= 0 && d.mode != DAV1D_CUDA_INTRA_PAL;
                int j = i + 1, bpw = 1;
                if (res) {
                    const int sw = w4[d.tx] * 4 < 32 ? w4[d.tx] * 4 : 32, sh = h4[d.tx] * 4 < 32 ? h4[d.tx] * 4 : 32;
                    bpw = 32 / (sh > sw ? sh : sw);
                    while (j < level_start[l + 1] && j - i < bpw && sd[j].eob >= 0 &&
                           sd[j].mode != DAV1D_CUDA_INTRA_PAL && sd[j].tx == d.tx) j++;
                }
                if (big == (pass == 1))
                    tasks[k++] = ((uint32_t)i << 8) | ((uint32_t)(res ? d.tx : 31) << 3) | (uint32_t)(j - i - 1);
                i = j;
            }
        }
    }
    task_start[2 * n_levels] = k;
    *n_tasks = k;
    return k;
}

int dav1d_cuda_intra_residual_tasks(const Dav1dCudaIntraDesc *sd, const int32_t *level_start, int n_levels,
                                    Dav1dCudaItxDesc *itx, uint32_t *tasks, int32_t *task_start, int32_t *n_tasks)
{
    if (!sd || !level_start || !itx || !tasks || !task_start || !n_tasks) return -22;
    int n_itx = 0, k = 0;
    std::vector<Dav1dCudaItxDesc> lv;
    for (int l = 0; l < n_levels; l++) {
        lv.clear();
        for (int i = level_start[l]; i < level_start[l + 1]; i++) {
            const Dav1dCudaIntraDesc &d = sd[i];
            if (d.eob < 0 || d.mode == DAV1D_CUDA_INTRA_PAL) continue;
            Dav1dCudaItxDesc t;
            memset(&t, 0, sizeof(t));
            t.coef_off = d.coef_off; t.x = (uint16_t)(d.x4 * 4); t.y = (uint16_t)(d.y4 * 4);
            t.eob = d.eob; t.plane = d.plane; t.tx = d.tx; t.txtp = d.txtp; t.cw4 = d.cw4; t.ch4 = d.ch4;
            lv.push_back(t);
        }
        std::stable_sort(lv.begin(), lv.end(), [](const Dav1dCudaItxDesc &x, const Dav1dCudaItxDesc &y) {
            const int kx = x.eob == 0 && x.txtp == 0 ? 0 : 1 + x.txtp, ky = y.eob == 0 && y.txtp == 0 ? 0 : 1 + y.txtp;
            return x.tx != y.tx ? x.tx < y.tx : kx < ky;
        });
        for (auto &t : lv) itx[n_itx++] = t;
        int ns = 0, nb = 0;
        const int made = itx_build_tasks(itx + n_itx - (int)lv.size(), (int)lv.size(), n_itx - (int)lv.size(),
                                         tasks + k, &ns, &nb);
        task_start[2 * l] = k;
        task_start[2 * l + 1] = k + ns;
        k += made;
    }
    task_start[2 * n_levels] = k;
    *n_tasks = k;
    return n_itx;
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
    return recon_submit_on(c, b, c->stream);
}

int dav1d_cuda_recon_submit_phases(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b, int phase_mask) {
    if (!c || !b || !b->dst) return -22;
    return recon_submit_on(c, b, c->stream, phase_mask);
}

int dav1d_cuda_recon_graph_build(Dav1dCudaContext *c, const Dav1dCudaReconBatch *b, Dav1dCudaReconGraph **out) {
    if (!c || !b || !out) return -22;
    *out = nullptr;
    cudaStream_t cap;
    D1_CHECK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    D1_CHECK(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    const int r = recon_submit_on(c, b, cap);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(cap, &graph);
    cudaStreamDestroy(cap);
    if (r) { if (graph) cudaGraphDestroy(graph); return r; }
    if (!cuda_ok(e, "cudaStreamEndCapture")) return -5;
    Dav1dCudaReconGraph *g = new Dav1dCudaReconGraph();
    g->graph = graph;
    g->tables = nullptr;
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

// Multi-frame graph: the frames of `n` independent streams in one graph.
int dav1d_cuda_recon_graph_build_multi(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n,
                                       Dav1dCudaReconGraph **out)
{
    return dav1d_cuda_recon_graph_build_multi_phases(c, bs, n, 31, out);
}

int dav1d_cuda_recon_graph_build_multi_phases(Dav1dCudaContext *c, const Dav1dCudaReconBatch *const *bs, int n,
                                              int phase_mask, Dav1dCudaReconGraph **out)
{
    if (!c || !bs || n < 1 || !out) return -22;
    *out = nullptr;
    if (n > 255) return -22;
    MultiTables t;
    if (build_multi_tables(bs, n, t)) return -22;
    const size_t fb = (t.frames.size() * sizeof(IntraFrameParams) + 255) & ~(size_t)255;
    const size_t sb = (t.items.size() * sizeof(Dav1dCudaIntraDesc) + 255) & ~(size_t)255;
    const size_t ib = (t.itx_frames.size() * sizeof(ItxFrameRef) + 255) & ~(size_t)255;
    const size_t rb = (t.rtasks.size() * sizeof(uint2) + 255) & ~(size_t)255;
    const size_t xb = (t.ritx.size() * sizeof(Dav1dCudaItxDesc) + 255) & ~(size_t)255;
    const size_t n_tail = t.tail_level >= 0 ? t.items.size() - (size_t)t.level_start[t.tail_level] : 0;
    const size_t tsb = (t.tail_dep_start.size() * sizeof(int32_t) + 255) & ~(size_t)255;
    const size_t tdb = (t.tail_deps.size() * sizeof(int32_t) + 255) & ~(size_t)255;
    const size_t tyb = ((n_tail + 1) * sizeof(unsigned) + 255) & ~(size_t)255;
    uint8_t *tab = nullptr;
    D1_CHECK(cudaMalloc(&tab, fb + sb + ib + rb + xb + tsb + tdb + tyb + 64));
    if (!t.tail_dep_start.empty())
        D1_CHECK(cudaMemcpy(tab + fb + sb + ib + rb + xb, t.tail_dep_start.data(),
                            t.tail_dep_start.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (!t.tail_deps.empty())
        D1_CHECK(cudaMemcpy(tab + fb + sb + ib + rb + xb + tsb, t.tail_deps.data(),
                            t.tail_deps.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    D1_CHECK(cudaMemcpy(tab, t.frames.data(), t.frames.size() * sizeof(IntraFrameParams), cudaMemcpyHostToDevice));
    if (!t.items.empty())
        D1_CHECK(cudaMemcpy(tab + fb, t.items.data(), t.items.size() * sizeof(Dav1dCudaIntraDesc),
                            cudaMemcpyHostToDevice));
    if (!t.itx_frames.empty())
        D1_CHECK(cudaMemcpy(tab + fb + sb, t.itx_frames.data(), t.itx_frames.size() * sizeof(ItxFrameRef),
                            cudaMemcpyHostToDevice));
    if (!t.rtasks.empty())
        D1_CHECK(cudaMemcpy(tab + fb + sb + ib, t.rtasks.data(), t.rtasks.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    if (!t.ritx.empty())
        D1_CHECK(cudaMemcpy(tab + fb + sb + ib + rb, t.ritx.data(), t.ritx.size() * sizeof(Dav1dCudaItxDesc),
                            cudaMemcpyHostToDevice));
    cudaStream_t cap;
    D1_CHECK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    D1_CHECK(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    const int r = recon_submit_multi_on(c, bs, n, (const IntraFrameParams *)tab, (const Dav1dCudaIntraDesc *)(tab + fb),
                                        (const ItxFrameRef *)(tab + fb + sb),
                                        (const Dav1dCudaItxDesc *)(tab + fb + sb + ib + rb),
                                        (const uint2 *)(tab + fb + sb + ib),
                                        (const int32_t *)(tab + fb + sb + ib + rb + xb),
                                        (const int32_t *)(tab + fb + sb + ib + rb + xb + tsb),
                                        (unsigned *)(tab + fb + sb + ib + rb + xb + tsb + tdb), t, cap, phase_mask);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(cap, &graph);
    cudaStreamDestroy(cap);
    if (r) { if (graph) cudaGraphDestroy(graph); cudaFree(tab); return r; }
    if (!cuda_ok(e, "cudaStreamEndCapture")) { cudaFree(tab); return -5; }
    Dav1dCudaReconGraph *g = new Dav1dCudaReconGraph();
    g->graph = graph;
    g->tables = tab;
    size_t nn = 0;
    cudaGraphGetNodes(graph, nullptr, &nn);
    g->n_nodes = (int)nn;
    if (!cuda_ok(cudaGraphInstantiate(&g->exec, graph, 0), "cudaGraphInstantiate")) {
        cudaGraphDestroy(graph);
        cudaFree(tab);
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
    if (g->tables) cudaFree(g->tables);
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
