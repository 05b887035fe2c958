// itxfm_add operator class on the compact transforms of itx2.cuh: the task kernel (all transform
// sizes of a launch class in one launch, several small blocks per warp in lane groups) for one
// frame or for the merged frames of a group, and the Dav1dInvTxfmDSPContext overrides built on
// the same kernel.  Reference: src/itx_tmpl.c:40-284 (driver + init), src/itx_1d.c.
#include <string.h>
#include <mutex>
#include <vector>
#include "ctx.h"
#include "itx2.cuh"

namespace d1 {

constexpr int ITX2_WARPS = 4;
constexpr int ITX2_INTS_SMALL = 2 * 16 * 17;   // per warp: sizes up to 16x16 (two 16x16 per warp)
constexpr int ITX2_INTS_MID = 32 * 33;         // per warp: one block of up to 32x32
constexpr int ITX2_INTS_BIG = 64 * 65;         // per warp: one block of up to 64x64

// Three ways to name the work of a launch:
//   tasks  : task codes (first_index << 8 | tx << 3 | count - 1) over `descs` of one frame
//   mtasks : (code, frame) with per-frame planes / coefficient streams from `frames`; the codes
//            index the frames' descriptors concatenated in `descs`
//   neither: `descs` grouped by tx in increasing tx; cls_task[t] / cls_desc[t] = first task /
//            descriptor of size t (tasks are derived: consecutive runs of 32 / G blocks)
struct Itx2Args {
    PicView pic;
    PicView res;                 // to_res: int16 planes that receive the residual instead of dst += residual
    int to_res;
    void *cf;
    int cf16;                    // high bit depth: compact int16 stream + escape list (Dav1dCudaReconBatch.cf_int16)
    const Dav1dCudaCoefEsc *esc;
    int n_esc;
    const Dav1dCudaItxDesc *descs;
    const uint32_t *tasks;
    const ItxFrameRef *frames;
    const uint2 *mtasks;
    int n_tasks;
    int zero_coefs;
    int cls_task[DAV1D_CUDA_N_RECT_TX_SIZES + 1];
    int cls_desc[DAV1D_CUDA_N_RECT_TX_SIZES + 1];
};

HD int itx2_group(const int tx) {              // lanes per block
    const TxDim t = tx_dim(tx);
    const int sw = t.w < 32 ? t.w : 32, sh = t.h < 32 ? t.h : 32;
    return sh > sw ? sh : sw;
}

// CLS: 0 sizes up to 16x16 (several blocks per warp), 1 larger sizes without a 64-point side, 2 the
// sizes with a 64-point side.  The task list of the larger sizes is handed to both CLS 1 and CLS 2;
// each takes its own tasks (the 64-point class needs four times the shared memory and twice the
// registers of the 32-point one, and would halve its occupancy).
template <int CLS> struct Itx2Cls {
    static constexpr int MAXN = CLS == 0 ? 16 : CLS == 1 ? 32 : 64;
    static constexpr int INTS = CLS == 0 ? ITX2_INTS_SMALL : CLS == 1 ? ITX2_INTS_MID : ITX2_INTS_BIG;
    static constexpr int MIN_BLOCKS = CLS == 0 ? 8 : CLS == 1 ? 5 : 3;
};
template <typename pixel, int CLS>
__global__ void __launch_bounds__(ITX2_WARPS * 32, Itx2Cls<CLS>::MIN_BLOCKS) itx2_task_kernel(const __grid_constant__ Itx2Args a) {
    constexpr bool BIG = CLS > 0;
    extern __shared__ int itx2_smem[];
    typedef typename PxTraits<pixel>::coef coef;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * ITX2_WARPS + warp;
    if (t >= a.n_tasks) return;
    int *smem = itx2_smem + warp * Itx2Cls<CLS>::INTS;
    const PicView *pic = &a.pic, *rpic = &a.res;
    void *cf = a.cf;
    int first, tx, cnt;
    if (a.mtasks) {
        const uint2 tk = a.mtasks[t];
        const ItxFrameRef *fr = a.frames + tk.y;
        pic = &fr->pic; cf = fr->cf; rpic = &fr->res;
        first = (int)(tk.x >> 8); tx = (tk.x >> 3) & 31; cnt = (int)(tk.x & 7) + 1;
    } else if (a.tasks) {
        const uint32_t code = a.tasks[t];
        first = (int)(code >> 8); tx = (code >> 3) & 31; cnt = (int)(code & 7) + 1;
    } else {
        tx = 0;
        while (tx < DAV1D_CUDA_N_RECT_TX_SIZES - 1 && t >= a.cls_task[tx + 1]) tx++;
        const int bpw = 32 / itx2_group(tx);
        first = a.cls_desc[tx] + (t - a.cls_task[tx]) * bpw;
        cnt = imin(bpw, a.cls_desc[tx + 1] - first);
    }
    if (CLS > 0) {
        const TxDim td = tx_dim(tx);
        if ((td.w == 64 || td.h == 64) != (CLS == 2)) return;      // the other large class takes it
    }
    const int G = itx2_group(tx);
    const int grp = lane / G, gl = lane % G;
    const bool active = grp < cnt;
    Dav1dCudaItxDesc d;
    if (active) d = a.descs[first + grp];
    else { d.coef_off = 0; d.x = d.y = 0; d.eob = 0; d.plane = 0; d.tx = 0; d.txtp = 0; d.cw4 = d.ch4 = 0; }
    const PlaneView &pv = pic->p[d.plane];
    const int dstride = (int)(pv.stride / (int)sizeof(pixel));
    pixel *dst = (pixel *)pv.data + (int64_t)d.y * dstride + d.x;
    int16_t *res = nullptr;
    int rstride = 0;
    if (a.to_res) {
        const PlaneView &rv = rpic->p[d.plane];
        rstride = (int)(rv.stride / 2);
        res = (int16_t *)rv.data + (int64_t)d.y * rstride + d.x;
    }
    if constexpr (sizeof(coef) == 4) {
        // the frames of a multi-frame task list carry native coefficients
        Itx2Coef c;
        static_assert(sizeof(Itx2Esc) == sizeof(Dav1dCudaCoefEsc), "escape entry layout");
        c.s16 = a.mtasks ? 0 : a.cf16; c.esc = (const Itx2Esc *)a.esc; c.n_esc = a.n_esc; c.off = d.coef_off;
        c.p = c.s16 ? (void *)((int16_t *)cf + d.coef_off) : (void *)((int32_t *)cf + d.coef_off);
        itx2_block<pixel, Itx2Cls<CLS>::MAXN, Itx2Coef>(active, gl, G, smem + grp * itx2_tile_ints(tx), c, tx, d.txtp, d.eob,
                                                        d.cw4, d.ch4, dst, dstride, res, rstride, pic->bdmax,
                                                        a.zero_coefs != 0);
    } else {
        itx2_block<pixel, Itx2Cls<CLS>::MAXN>(active, gl, G, smem + grp * itx2_tile_ints(tx), (coef *)cf + d.coef_off, tx,
                                              d.txtp, d.eob, d.cw4, d.ch4, dst, dstride, res, rstride, pic->bdmax,
                                              a.zero_coefs != 0);
    }
}

template <typename pixel, int CLS>
static int itx2_launch_cls(Itx2Args a, int n, cudaStream_t st) {
    a.n_tasks = n;
    const int grid = (n + ITX2_WARPS - 1) / ITX2_WARPS;
    const size_t smem = (size_t)ITX2_WARPS * Itx2Cls<CLS>::INTS * sizeof(int);
    if (CLS == 2) {      // more than the default 48 KB (the per-call surface launches without a context)
        static std::once_flag once;
        std::call_once(once, [&] {
            cudaFuncSetAttribute(itx2_task_kernel<pixel, CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        });
    }
    itx2_task_kernel<pixel, CLS><<<grid, ITX2_WARPS * 32, smem, st>>>(a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "itx2_task_kernel") ? 0 : -5;
}
template <typename pixel, bool BIG>
static int itx2_launch_one(Itx2Args a, int n, cudaStream_t st) {
    if (n <= 0) return 0;
    if (!BIG) return itx2_launch_cls<pixel, 0>(a, n, st);
    const int r = itx2_launch_cls<pixel, 1>(a, n, st);
    return r ? r : itx2_launch_cls<pixel, 2>(a, n, st);
}

static int itx2_launch_both(Itx2Args a, int n_small, int n_big, bool hbd, cudaStream_t st_small, cudaStream_t st_big) {
    int r = hbd ? itx2_launch_one<uint16_t, false>(a, n_small, st_small) : itx2_launch_one<uint8_t, false>(a, n_small, st_small);
    if (r) return r;
    if (a.tasks) a.tasks += n_small;
    if (a.mtasks) a.mtasks += n_small;
    return hbd ? itx2_launch_one<uint16_t, true>(a, n_big, st_big) : itx2_launch_one<uint8_t, true>(a, n_big, st_big);
}

void itx_init_attrs() {}

// tasks[0 .. n_small) = sizes up to 16x16, tasks[n_small .. n_small + n_big) = larger
int itx_task_launch(const PicView &pic, const PicView *res, void *cf, const Dav1dCudaItxDesc *descs,
                    const uint32_t *tasks, int n_small, int n_big, int zero_coefs, cudaStream_t st_small,
                    cudaStream_t st_big, const CoefFmt *fmt)
{
    Itx2Args a;
    memset(&a, 0, sizeof(a));
    a.pic = pic; a.cf = cf; a.descs = descs; a.tasks = tasks; a.zero_coefs = zero_coefs;
    if (fmt) { a.cf16 = fmt->s16; a.esc = fmt->esc; a.n_esc = fmt->n_esc; }
    if (res) { a.res = *res; a.to_res = 1; }
    return itx2_launch_both(a, n_small, n_big, pic.bdmax > 0xff, st_small, st_big);
}

// the frames of a group: tasks = (code, frame), code indexes `descs` = the descriptors of all
// frames concatenated
int itx_multi_task_launch(const ItxFrameRef *frames, const Dav1dCudaItxDesc *descs, const uint2 *tasks, int n_small,
                          int n_big, bool hbd, bool to_res, cudaStream_t st_small, cudaStream_t st_big)
{
    Itx2Args a;
    memset(&a, 0, sizeof(a));
    a.frames = frames; a.mtasks = tasks; a.descs = descs; a.to_res = to_res;
    return itx2_launch_both(a, n_small, n_big, hbd, st_small, st_big);
}

static bool tx_is_big(int tx) { const TxDim t = tx_dim(tx); return t.w > 16 || t.h > 16; }

// descriptors grouped by tx (class_count[t] of size t, increasing t): one launch for the sizes up
// to 16x16, one for the larger ones; the tasks are implicit
int itx_batch_launch(const PicView &pic, const PicView *res, void *cf, const Dav1dCudaItxDesc *descs,
                     const int32_t *class_count, int zero_coefs, cudaStream_t st, const CoefFmt *fmt)
{
    const bool hbd = pic.bdmax > 0xff;
    for (int pass = 0; pass < 2; pass++) {
        Itx2Args a;
        memset(&a, 0, sizeof(a));
        a.pic = pic; a.cf = cf; a.descs = descs; a.zero_coefs = zero_coefs;
        if (fmt) { a.cf16 = fmt->s16; a.esc = fmt->esc; a.n_esc = fmt->n_esc; }
        if (res) { a.res = *res; a.to_res = 1; }
        int off = 0, nt = 0;
        for (int tx = 0; tx < DAV1D_CUDA_N_RECT_TX_SIZES; tx++) {
            const int n = class_count[tx] > 0 ? class_count[tx] : 0;
            a.cls_task[tx] = nt; a.cls_desc[tx] = off;
            if (tx_is_big(tx) == (pass == 1)) {        // sizes of the other pass own no tasks
                const int bpw = 32 / itx2_group(tx);
                nt += (n + bpw - 1) / bpw;
            }
            off += n;
        }
        a.cls_task[DAV1D_CUDA_N_RECT_TX_SIZES] = nt;
        a.cls_desc[DAV1D_CUDA_N_RECT_TX_SIZES] = off;
        int r;
        if (pass == 0) r = hbd ? itx2_launch_one<uint16_t, false>(a, nt, st) : itx2_launch_one<uint8_t, false>(a, nt, st);
        else r = hbd ? itx2_launch_one<uint16_t, true>(a, nt, st) : itx2_launch_one<uint8_t, true>(a, nt, st);
        if (r) return r;
    }
    return 0;
}

// blocks of one size a warp takes
static int itx_bpw(int tx) { return 32 / itx2_group(tx); }

// host: task codes for `n` descriptors (host copy) that are grouped by tx; small sizes first
int itx_build_tasks(const Dav1dCudaItxDesc *descs, int n, int index_base, uint32_t *tasks, int *n_small, int *n_big) {
    int k = 0;
    *n_small = *n_big = 0;
    for (int pass = 0; pass < 2; pass++) {
        int i = 0;
        while (i < n) {
            const int tx = descs[i].tx;
            int j = i;
            while (j < n && descs[j].tx == tx) j++;
            const bool big = tx_is_big(tx);
            if (big == (pass == 1)) {
                const int bpw = itx_bpw(tx);
                for (int f = i; f < j; f += bpw) {
                    const int cnt = (j - f) < bpw ? (j - f) : bpw;
                    tasks[k++] = ((uint32_t)(index_base + f) << 8) | ((uint32_t)tx << 3) | (uint32_t)(cnt - 1);
                    if (big) (*n_big)++; else (*n_small)++;
                }
            }
            i = j;
        }
    }
    return k;
}

// ----------------------------------------------------------- per-call surface
// itxfm_fn: void (pixel *dst, ptrdiff_t stride, coef *coeff, int eob [, int bitdepth_max]).
// The block's non-zero bounding box is found here, on the host, and shipped packed - the same
// format and the same kernel as the batched path; the caller's coefficient buffer is cleared as
// the contract requires (itx_tmpl.c:89).
template <typename pixel>
static void itx_single(const int tx, const int txtp, pixel *dst, const ptrdiff_t stride,
                       typename PxTraits<pixel>::coef *coeff, const int eob, const int bdmax)
{
    typedef typename PxTraits<pixel>::coef coef;
    const TxDim td = tx_dim(tx);
    const int w = td.w, h = td.h;
    const int sw = w < 32 ? w : 32, sh = h < 32 ? h : 32;
    const bool dc_only = eob == 0 && txtp == 0;
    int nzw = 1, nzh = 1;
    if (!dc_only)
        for (int x = 0; x < sw; x++)
            for (int y = 0; y < sh; y++)
                if (coeff[y + x * sh]) { if (x + 1 > nzw) nzw = x + 1; if (y + 1 > nzh) nzh = y + 1; }
    const int cw = (nzw + 3) & ~3, ch = (nzh + 3) & ~3;
    const size_t cf_bytes = (size_t)cw * ch * sizeof(coef);
    const size_t row_bytes = (size_t)w * sizeof(pixel);
    const size_t tile_stride = (row_bytes + 63) & ~(size_t)63;
    const size_t cf_off = 0, px_off = (cf_bytes + 255) & ~(size_t)255;
    const size_t desc_off = px_off + tile_stride * h;
    const size_t total = desc_off + sizeof(Dav1dCudaItxDesc);

    Staging &s = staging();
    std::lock_guard<std::mutex> lk(s.mu);
    if (!s.ensure(total)) return;
    coef *pk = (coef *)(s.host + cf_off);
    for (int x = 0; x < cw; x++)
        for (int y = 0; y < ch; y++) pk[y + x * ch] = coeff[y + x * sh];
    if (dc_only) coeff[0] = 0;
    else memset(coeff, 0, (size_t)sw * sh * sizeof(coef));
    const ptrdiff_t pxstride = stride / (ptrdiff_t)sizeof(pixel);
    for (int y = 0; y < h; y++)
        memcpy(s.host + px_off + y * tile_stride, dst + y * pxstride, row_bytes);
    Dav1dCudaItxDesc d;
    memset(&d, 0, sizeof(d));
    d.eob = (int16_t)eob;
    d.tx = (uint8_t)tx;
    d.txtp = (uint8_t)txtp;
    d.cw4 = (uint8_t)(cw / 4); d.ch4 = (uint8_t)(ch / 4);
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
    if (itx_batch_launch(pv, nullptr, s.dev + cf_off, (const Dav1dCudaItxDesc *)(s.dev + desc_off), cls, 0, s.stream, nullptr))
        return;
    D1_CHECKV(cudaMemcpyAsync(s.host + px_off, s.dev + px_off, tile_stride * h, cudaMemcpyDeviceToHost, s.stream));
    D1_CHECKV(cudaStreamSynchronize(s.stream));
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
    return itx_task_launch(pic_view(dst), nullptr, cf, descs, tasks, n_small, n_big, zero_coefs, c->stream, c->stream, nullptr);
}

int dav1d_cuda_itx_batch(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, void *cf,
                         const Dav1dCudaItxDesc *descs,
                         const int32_t class_count[DAV1D_CUDA_N_RECT_TX_SIZES], int zero_coefs)
{
    if (!c || !dst || !descs || !class_count) return -22;
    return itx_batch_launch(pic_view(dst), nullptr, cf, descs, class_count, zero_coefs, c->stream, nullptr);
}

}  // extern "C"
