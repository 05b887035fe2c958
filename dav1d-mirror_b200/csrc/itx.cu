// itxfm_add operator class: batched kernel (one launch per transform size)
// and the Dav1dInvTxfmDSPContext overrides built on the same kernel.
// Reference: src/itx_tmpl.c:40-284 (driver + init), src/itx_1d.c (1-D kernels).
#include <string.h>
#include "ctx.h"
#include "itx.cuh"

namespace d1 {

struct ItxArgs {
    PicView pic;
    void *cf;
    const Dav1dCudaItxDesc *descs;
    int n;
    int zero_coefs;
};

constexpr int ITX_WARPS = 4;

template <typename pixel, int W, int H>
__global__ void __launch_bounds__(ITX_WARPS * 32) itx_kernel(const ItxArgs a) {
    typedef ItxGeom<W, H> Geo;
    typedef typename PxTraits<pixel>::coef coef;
    constexpr int G = Geo::GMIN;
    constexpr int BPW = 32 / G;
    extern __shared__ int itx_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    const int blk = (blockIdx.x * ITX_WARPS + warp) * BPW + grp;
    const bool active = blk < a.n;
    Dav1dCudaItxDesc d;
    if (active) d = a.descs[blk];
    else { d.coef_off = 0; d.x = d.y = 0; d.eob = 0; d.plane = 0; d.tx = 0; d.txtp = 0; d.cw4 = d.ch4 = 0; }
    int *tile = itx_smem + (warp * BPW + grp) * Geo::TILE_INTS;
    const PlaneView pv = a.pic.p[d.plane];
    const int dstride = (int)(pv.stride / (int)sizeof(pixel));
    pixel *dst = (pixel *)pv.data + (int64_t)d.y * dstride + d.x;
    coef *cf = (coef *)a.cf + d.coef_off;
    itx_block<pixel, W, H, G>(active, gl, tile, cf, d.eob, d.txtp, dst, dstride,
                              a.pic.bdmax, a.zero_coefs != 0, d.cw4, d.ch4);
}

template <typename pixel, int W, int H>
static int launch_itx_class(const ItxArgs &a, cudaStream_t st) {
    typedef ItxGeom<W, H> Geo;
    constexpr int BPW = 32 / Geo::GMIN;
    const int per_cta = ITX_WARPS * BPW;
    const int grid = (a.n + per_cta - 1) / per_cta;
    const size_t smem = (size_t)ITX_WARPS * BPW * Geo::TILE_INTS * sizeof(int);
    itx_kernel<pixel, W, H><<<grid, ITX_WARPS * 32, smem, st>>>(a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "itx_kernel launch") ? 0 : -5;
}

template <typename pixel>
static int launch_itx_tx(int tx, const ItxArgs &a, cudaStream_t st) {
    switch (tx) {
    case 0: return launch_itx_class<pixel, 4, 4>(a, st);
    case 1: return launch_itx_class<pixel, 8, 8>(a, st);
    case 2: return launch_itx_class<pixel, 16, 16>(a, st);
    case 3: return launch_itx_class<pixel, 32, 32>(a, st);
    case 4: return launch_itx_class<pixel, 64, 64>(a, st);
    case 5: return launch_itx_class<pixel, 4, 8>(a, st);
    case 6: return launch_itx_class<pixel, 8, 4>(a, st);
    case 7: return launch_itx_class<pixel, 8, 16>(a, st);
    case 8: return launch_itx_class<pixel, 16, 8>(a, st);
    case 9: return launch_itx_class<pixel, 16, 32>(a, st);
    case 10: return launch_itx_class<pixel, 32, 16>(a, st);
    case 11: return launch_itx_class<pixel, 32, 64>(a, st);
    case 12: return launch_itx_class<pixel, 64, 32>(a, st);
    case 13: return launch_itx_class<pixel, 4, 16>(a, st);
    case 14: return launch_itx_class<pixel, 16, 4>(a, st);
    case 15: return launch_itx_class<pixel, 8, 32>(a, st);
    case 16: return launch_itx_class<pixel, 32, 8>(a, st);
    case 17: return launch_itx_class<pixel, 16, 64>(a, st);
    case 18: return launch_itx_class<pixel, 64, 16>(a, st);
    }
    return -22;
}

// (the task kernels - all transform sizes in one launch - live in itx_task.cuh, instantiated per
// pixel type and size group in itx_task_{s,b}{8,16}.cu, host side in itx_task.cu)
int itx_task_launch(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs, const uint32_t *tasks,
                    int n_small, int n_big, int zero_coefs, cudaStream_t st_small, cudaStream_t st_big);
int itx_build_tasks(const Dav1dCudaItxDesc *descs, int n, int index_base, uint32_t *tasks, int *n_small, int *n_big);

// `streams`/`n_streams`: the size classes touch disjoint pixels, so they are spread
// round-robin over the given streams (heaviest classes first on their own stream).
int itx_batch_launch_multi(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs,
                           const int32_t *class_count, int zero_coefs, cudaStream_t *streams, int n_streams)
{
    int off = 0, k = 0;
    for (int tx = 0; tx < DAV1D_CUDA_N_RECT_TX_SIZES; tx++) {
        const int n = class_count[tx];
        if (n > 0) {
            ItxArgs a;
            a.pic = pic;
            a.cf = cf;
            a.descs = descs + off;
            a.n = n;
            a.zero_coefs = zero_coefs;
            cudaStream_t st = streams[k++ % n_streams];
            const int r = pic.bdmax > 0xff ? launch_itx_tx<uint16_t>(tx, a, st) : launch_itx_tx<uint8_t>(tx, a, st);
            if (r) return r;
            off += n;
        }
    }
    return 0;
}

int itx_batch_launch(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs,
                     const int32_t *class_count, int zero_coefs, cudaStream_t st)
{
    int off = 0;
    for (int tx = 0; tx < DAV1D_CUDA_N_RECT_TX_SIZES; tx++) {
        const int n = class_count[tx];
        if (n <= 0) continue;
        ItxArgs a;
        a.pic = pic;
        a.cf = cf;
        a.descs = descs + off;
        a.n = n;
        a.zero_coefs = zero_coefs;
        const int r = pic.bdmax > 0xff ? launch_itx_tx<uint16_t>(tx, a, st)
                                       : launch_itx_tx<uint8_t>(tx, a, st);
        if (r) return r;
        off += n;
    }
    return 0;
}

// ----------------------------------------------------------- per-call surface
// itxfm_fn: void (pixel *dst, ptrdiff_t stride, coef *coeff, int eob [, int bitdepth_max])
template <typename pixel>
static void itx_single(const int tx, const int txtp, pixel *dst, const ptrdiff_t stride,
                       typename PxTraits<pixel>::coef *coeff, const int eob, const int bdmax)
{
    typedef typename PxTraits<pixel>::coef coef;
    const TxDim td = tx_dim(tx);
    const int w = td.w, h = td.h;
    const int sw = w < 32 ? w : 32, sh = h < 32 ? h : 32;
    const size_t cf_bytes = (size_t)sw * sh * sizeof(coef);
    const size_t row_bytes = (size_t)w * sizeof(pixel);
    const size_t tile_stride = (row_bytes + 63) & ~(size_t)63;
    const size_t cf_off = 0, px_off = (cf_bytes + 255) & ~(size_t)255;
    const size_t desc_off = px_off + tile_stride * h;
    const size_t total = desc_off + sizeof(Dav1dCudaItxDesc);

    Staging &s = staging();
    std::lock_guard<std::mutex> lk(s.mu);
    if (!s.ensure(total)) return;
    memcpy(s.host + cf_off, coeff, cf_bytes);
    const ptrdiff_t pxstride = stride / (ptrdiff_t)sizeof(pixel);
    for (int y = 0; y < h; y++)
        memcpy(s.host + px_off + y * tile_stride, dst + y * pxstride, row_bytes);
    Dav1dCudaItxDesc d;
    memset(&d, 0, sizeof(d));
    d.coef_off = 0;
    d.x = 0; d.y = 0;
    d.eob = (int16_t)eob;
    d.plane = 0;
    d.tx = (uint8_t)tx;
    d.txtp = (uint8_t)txtp;
    memcpy(s.host + desc_off, &d, sizeof(d));
    D1_CHECKV(cudaMemcpyAsync(s.dev, s.host, total, cudaMemcpyHostToDevice, s.stream));

    PicView pv;
    memset(&pv, 0, sizeof(pv));
    pv.p[0].data = s.dev + px_off;
    pv.p[0].stride = (int64_t)tile_stride;
    pv.p[0].w = w;
    pv.p[0].h = h;
    pv.bdmax = bdmax;
    int32_t cls[DAV1D_CUDA_N_RECT_TX_SIZES] = { 0 };
    cls[tx] = 1;
    if (itx_batch_launch(pv, s.dev + cf_off, (const Dav1dCudaItxDesc *)(s.dev + desc_off), cls, 1, s.stream))
        return;
    D1_CHECKV(cudaMemcpyAsync(s.host, s.dev, desc_off, cudaMemcpyDeviceToHost, s.stream));
    D1_CHECKV(cudaStreamSynchronize(s.stream));
    memcpy(coeff, s.host + cf_off, cf_bytes);
    for (int y = 0; y < h; y++)
        memcpy(dst + y * pxstride, s.host + px_off + y * tile_stride, row_bytes);
}

template <int TX, int TXTP>
static void itx_8bpc(uint8_t *dst, ptrdiff_t stride, int16_t *coeff, int eob) {
    itx_single<uint8_t>(TX, TXTP, dst, stride, coeff, eob, 0xff);
}
template <int TX, int TXTP>
static void itx_16bpc(uint16_t *dst, ptrdiff_t stride, int32_t *coeff, int eob, int bitdepth_max) {
    itx_single<uint16_t>(TX, TXTP, dst, stride, coeff, eob, bitdepth_max);
}

// Which (tx, txtp) slots the reference populates: itx_tmpl.c:248-268.
// class 84: all 16 types; class 16: 12 types (no 1-D identity x adst/flipadst);
// class 32: DCT_DCT + IDTX; class 64: DCT_DCT only.
template <bool HBD, int TX, int TXTP> struct Slot {
    static void *get() { return HBD ? (void *)itx_16bpc<TX, TXTP> : (void *)itx_8bpc<TX, TXTP>; }
};

template <bool HBD, int TX> static void fill64(Dav1dCudaInvTxfmDSPContext *c) {
    c->itxfm_add[TX][0] = Slot<HBD, TX, 0>::get();
}
template <bool HBD, int TX> static void fill32(Dav1dCudaInvTxfmDSPContext *c) {
    fill64<HBD, TX>(c);
    c->itxfm_add[TX][9] = Slot<HBD, TX, 9>::get();
}
template <bool HBD, int TX> static void fill16(Dav1dCudaInvTxfmDSPContext *c) {
    fill32<HBD, TX>(c);
    c->itxfm_add[TX][1] = Slot<HBD, TX, 1>::get();
    c->itxfm_add[TX][2] = Slot<HBD, TX, 2>::get();
    c->itxfm_add[TX][3] = Slot<HBD, TX, 3>::get();
    c->itxfm_add[TX][4] = Slot<HBD, TX, 4>::get();
    c->itxfm_add[TX][5] = Slot<HBD, TX, 5>::get();
    c->itxfm_add[TX][6] = Slot<HBD, TX, 6>::get();
    c->itxfm_add[TX][7] = Slot<HBD, TX, 7>::get();
    c->itxfm_add[TX][8] = Slot<HBD, TX, 8>::get();
    c->itxfm_add[TX][10] = Slot<HBD, TX, 10>::get();
    c->itxfm_add[TX][11] = Slot<HBD, TX, 11>::get();
}
template <bool HBD, int TX> static void fill84(Dav1dCudaInvTxfmDSPContext *c) {
    fill16<HBD, TX>(c);
    c->itxfm_add[TX][12] = Slot<HBD, TX, 12>::get();
    c->itxfm_add[TX][13] = Slot<HBD, TX, 13>::get();
    c->itxfm_add[TX][14] = Slot<HBD, TX, 14>::get();
    c->itxfm_add[TX][15] = Slot<HBD, TX, 15>::get();
}

template <bool HBD> static void fill_itx(Dav1dCudaInvTxfmDSPContext *c) {
    Staging &s = staging();
    {
        std::lock_guard<std::mutex> lk(s.mu);
        if (!s.ensure(1 << 20)) return;   // no device: leave the table untouched
    }
    c->itxfm_add[0][16] = Slot<HBD, 0, 16>::get();   // WHT_WHT 4x4
    fill84<HBD, 0>(c);    // 4x4
    fill84<HBD, 5>(c);    // 4x8
    fill84<HBD, 13>(c);   // 4x16
    fill84<HBD, 6>(c);    // 8x4
    fill84<HBD, 1>(c);    // 8x8
    fill84<HBD, 7>(c);    // 8x16
    fill32<HBD, 15>(c);   // 8x32
    fill84<HBD, 14>(c);   // 16x4
    fill84<HBD, 8>(c);    // 16x8
    fill16<HBD, 2>(c);    // 16x16
    fill32<HBD, 9>(c);    // 16x32
    fill64<HBD, 17>(c);   // 16x64
    fill32<HBD, 16>(c);   // 32x8
    fill32<HBD, 10>(c);   // 32x16
    fill32<HBD, 3>(c);    // 32x32
    fill64<HBD, 11>(c);   // 32x64
    fill64<HBD, 18>(c);   // 64x16
    fill64<HBD, 12>(c);   // 64x32
    fill64<HBD, 4>(c);    // 64x64
}

}  // namespace d1

using namespace d1;

extern "C" {

void dav1d_cuda_itx_dsp_init_8bpc(Dav1dCudaInvTxfmDSPContext *c, int bpc) {
    (void)bpc;
    fill_itx<false>(c);
}
void dav1d_cuda_itx_dsp_init_16bpc(Dav1dCudaInvTxfmDSPContext *c, int bpc) {
    (void)bpc;
    fill_itx<true>(c);
}

int dav1d_cuda_itx_tasks(const Dav1dCudaItxDesc *descs_host, int n, int index_base, uint32_t *tasks,
                         int32_t *n_small, int32_t *n_big)
{
    if (!descs_host || !tasks || !n_small || !n_big || n < 0) return -22;
    int a = 0, b = 0;
    const int k = itx_build_tasks(descs_host, n, index_base, tasks, &a, &b);
    *n_small = a; *n_big = b;
    return k;
}

int dav1d_cuda_itx_task_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, void *cf,
                              const Dav1dCudaItxDesc *descs, const uint32_t *tasks, int n_small, int n_big,
                              int zero_coefs)
{
    if (!c || !dst || !descs || !tasks) return -22;
    return itx_task_launch(pic_view(dst), cf, descs, tasks, n_small, n_big, zero_coefs, c->stream, c->stream);
}

int dav1d_cuda_itx_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, void *cf,
                         const Dav1dCudaItxDesc *descs,
                         const int32_t class_count[DAV1D_CUDA_N_RECT_TX_SIZES], int zero_coefs)
{
    if (!c || !dst || !descs || !class_count) return -22;
    return itx_batch_launch(pic_view(dst), cf, descs, class_count, zero_coefs, c->stream);
}

}  // extern "C"
