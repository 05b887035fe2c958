// Intra-prediction operator classes: kernels and the
// Dav1dIntraPredDSPContext overrides (+ on-device dav1d_prepare_intra_edges).
// Reference: src/ipred_tmpl.c, src/ipred_prepare_tmpl.c.
#include <string.h>
#include <type_traits>
#include "ctx.h"
#include "stage.h"
#include "ipred.cuh"

namespace d1 {

constexpr int EDGE_BUF = 288;     // pixels; centre at EDGE_C
constexpr int EDGE_C = 144;

struct IpredArgs {
    int mode, w, h, angle, max_w, max_h, bdmax;
    void *dst; int dstride;
    const void *edge;             // device copy of topleft[-2h .. 2w], pointer to the centre
    const int16_t *ac; int alpha; // cfl_pred
};

// one block by one warp: the same set-up + pixel loop the batched executor runs (recon2.cu)
template <typename pixel, bool CFL> __global__ void ipred_kernel(const IpredArgs a) {
    __shared__ pixel s_edge[EDGE_BUF];
    __shared__ pixel s_scratch[IPRED_SCRATCH];
    __shared__ pixel s_tile[32 * 32];
    const int lane = threadIdx.x;
    const Grp g = grp_warp(lane);
    const pixel *ge = (const pixel *)a.edge;
    for (int i = -2 * a.h + lane; i <= 2 * a.w; i += 32) s_edge[EDGE_C + i] = ge[i];
    __syncwarp();
    PixParams<pixel> P;
    if (CFL) P = cfl_setup<pixel>(g, a.mode, s_edge + EDGE_C, a.w, a.h, a.ac, a.alpha, a.bdmax);
    else P = ipred_setup<pixel>(g, a.mode, a.angle, a.w, a.h, a.max_w, a.max_h, s_edge + EDGE_C, s_scratch, s_tile, a.bdmax);
    pixel *dst = (pixel *)a.dst;
    const int lw = 31 - __clz(a.w);
    for (int i = lane; i < a.w * a.h; i += 32) {
        const int y = i >> lw, x = i & (a.w - 1);
        dst[y * a.dstride + x] = (pixel)ipred_pixel<pixel>(P, x, y, i, a.bdmax);
    }
}

struct CflAcArgs { int16_t *ac; const void *y; int ystride, w_pad, h_pad, w, h, ss_hor, ss_ver; };
template <typename pixel> __global__ void cfl_ac_kernel(const CflAcArgs a) {
    cfl_ac_block<pixel>(grp_warp(threadIdx.x), a.ac, (const pixel *)a.y, a.ystride, a.w_pad, a.h_pad, a.w, a.h,
                        a.ss_hor, a.ss_ver);
}

// pal_pred (ipred_tmpl.c:717-730): two pixels per index byte
struct PalArgs { void *dst; int dstride; const void *pal; const uint8_t *idx; int w, h; };
template <typename pixel> __global__ void pal_pred_kernel(const PalArgs a) {
    __shared__ pixel s_pal[8];
    if (threadIdx.x < 8) s_pal[threadIdx.x] = ((const pixel *)a.pal)[threadIdx.x];
    __syncthreads();
    pixel *dst = (pixel *)a.dst;
    const int lw = 31 - __clz(a.w);
    for (int i = threadIdx.x; i < a.w * a.h; i += blockDim.x) {
        const int y = i >> lw, x = i & (a.w - 1);
        const int v = a.idx[i >> 1];
        dst[y * a.dstride + x] = s_pal[(x & 1) ? v >> 4 : v & 7];
    }
}

struct PrepArgs {
    int x, have_left, y, have_top, w, h, edge_flags;
    const void *dst; int stride; const void *top_sb_edge;
    int mode, angle, tw, th, filter_edge, bdmax;
    void *edge_out;               // device, centre pointer
    int *result;                  // [0] = mode, [1] = angle
};
template <typename pixel> __global__ void prepare_edges_kernel(const PrepArgs a) {
    __shared__ pixel s_edge[EDGE_BUF];
    const int lane = threadIdx.x;
    int angle = a.angle;
    const int m = prepare_edges<pixel>(grp_warp(lane), a.x, a.have_left, a.y, a.have_top, a.w, a.h, a.edge_flags,
                                       (const pixel *)a.dst, a.stride, (const pixel *)a.top_sb_edge, a.mode,
                                       &angle, a.tw, a.th, a.filter_edge, s_edge + EDGE_C, a.bdmax);
    // copy out only what the reference defines for this mode is not knowable by the
    // caller; the per-call wrapper merges by the same `needs` rules on the host.
    pixel *out = (pixel *)a.edge_out;
    for (int i = -2 * a.th * 4 + lane; i <= 2 * a.tw * 4; i += 32) out[i] = s_edge[EDGE_C + i];
    if (lane == 0) { a.result[0] = m; a.result[1] = angle; }
}

// ------------------------------------------------------------ per-call surface
// angular_ipred_fn (src/ipred.h:44-48)
template <typename pixel>
static void ipred_single(const int mode, pixel *dst, const ptrdiff_t stride, const pixel *topleft,
                         const int w, const int h, const int angle, const int max_w, const int max_h,
                         const int bdmax, const int16_t *ac, const int alpha, const bool cfl)
{
    const size_t ostride = ((size_t)w * sizeof(pixel) + 63) & ~(size_t)63;
    const int elo = 2 * h, ehi = 2 * w;
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t o_edge = st.reserve((size_t)(elo + ehi + 1) * sizeof(pixel));
    const size_t o_ac = st.reserve(cfl ? (size_t)w * h * 2 : 1);
    const size_t o_out = st.reserve(ostride * h);
    if (!st.commit()) return;
    memcpy(st.host(o_edge), topleft - elo, (size_t)(elo + ehi + 1) * sizeof(pixel));
    if (cfl) memcpy(st.host(o_ac), ac, (size_t)w * h * 2);
    if (!st.upload()) return;
    IpredArgs a;
    a.mode = mode; a.w = w; a.h = h; a.angle = angle; a.max_w = max_w; a.max_h = max_h; a.bdmax = bdmax;
    a.dst = st.dev(o_out); a.dstride = (int)(ostride / sizeof(pixel));
    a.edge = (const pixel *)st.dev(o_edge) + elo;
    a.ac = (const int16_t *)st.dev(o_ac); a.alpha = alpha;
    if (cfl) ipred_kernel<pixel, true><<<1, 32, 0, st.stream()>>>(a);
    else ipred_kernel<pixel, false><<<1, 32, 0, st.stream()>>>(a);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "ipred_kernel")) return;
    if (!st.download(o_out, ostride * h) || !st.sync()) return;
    st.get2d(o_out, ostride, dst, stride, (size_t)w * sizeof(pixel), h);
}

// cfl_ac_fn (src/ipred.h:56-59)
template <typename pixel>
static void cfl_ac_single(int16_t *ac, const pixel *ypx, const ptrdiff_t stride, const int w_pad, const int h_pad,
                          const int cw, const int ch, const int ss_hor, const int ss_ver)
{
    const int lw = (cw - 4 * w_pad) << ss_hor, lh = (ch - 4 * h_pad) << ss_ver;   // luma samples actually read
    const size_t ystride = ((size_t)lw * sizeof(pixel) + 63) & ~(size_t)63;
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t o_y = st.reserve(ystride * lh);
    const size_t o_ac = st.reserve((size_t)cw * ch * 2);
    if (!st.commit()) return;
    st.put2d(o_y, ystride, ypx, stride, (size_t)lw * sizeof(pixel), lh);
    if (!st.upload()) return;
    CflAcArgs a;
    a.ac = (int16_t *)st.dev(o_ac);
    a.y = st.dev(o_y); a.ystride = (int)(ystride / sizeof(pixel));
    a.w_pad = w_pad; a.h_pad = h_pad; a.w = cw; a.h = ch; a.ss_hor = ss_hor; a.ss_ver = ss_ver;
    cfl_ac_kernel<pixel><<<1, 32, 0, st.stream()>>>(a);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "cfl_ac_kernel")) return;
    if (!st.download(o_ac, (size_t)cw * ch * 2) || !st.sync()) return;
    memcpy(ac, st.host(o_ac), (size_t)cw * ch * 2);
}

// pal_pred_fn (src/ipred.h:76-79)
template <typename pixel>
static void pal_pred_single(pixel *dst, const ptrdiff_t stride, const pixel *pal, const uint8_t *idx,
                            const int w, const int h)
{
    const size_t ostride = ((size_t)w * sizeof(pixel) + 63) & ~(size_t)63;
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t o_pal = st.reserve(8 * sizeof(pixel));
    const size_t o_idx = st.reserve((size_t)w * h / 2);
    const size_t o_out = st.reserve(ostride * h);
    if (!st.commit()) return;
    memcpy(st.host(o_pal), pal, 8 * sizeof(pixel));
    memcpy(st.host(o_idx), idx, (size_t)w * h / 2);
    if (!st.upload()) return;
    PalArgs a;
    a.dst = st.dev(o_out); a.dstride = (int)(ostride / sizeof(pixel));
    a.pal = st.dev(o_pal); a.idx = st.dev(o_idx); a.w = w; a.h = h;
    pal_pred_kernel<pixel><<<1, 256, 0, st.stream()>>>(a);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "pal_pred_kernel")) return;
    if (!st.download(o_out, ostride * h) || !st.sync()) return;
    st.get2d(o_out, ostride, dst, stride, (size_t)w * sizeof(pixel), h);
}

// dav1d_prepare_intra_edges (src/ipred_prepare.h:77-84)
template <typename pixel>
static int prepare_single(const int x, const int have_left, const int y, const int have_top, const int w, const int h,
                          const int edge_flags, const pixel *dst, const ptrdiff_t stride, const pixel *top_sb_edge,
                          const int mode, int *angle, const int tw, const int th, const int filter_edge,
                          pixel *topleft_out, const int bdmax)
{
    // Stage exactly the neighbourhood the reference may touch: the row above
    // (cols -1 .. T-1) and the column to the left (rows 0 .. L-1).
    const int T = have_top ? imin(2 * tw * 4, (w - x) * 4) : 0;
    const int L = have_left ? imin(2 * th * 4, (h - y) * 4) : 0;
    const ptrdiff_t spx = stride / (ptrdiff_t)sizeof(pixel);
    const int ww = T + 1, wh = L + 1;
    const size_t wstride = (size_t)ww * sizeof(pixel);
    const int span = 2 * th * 4 + 2 * tw * 4 + 1;
    Staging &sg = staging();
    std::lock_guard<std::mutex> lk(sg.mu);
    Stage st(sg);
    const size_t o_win = st.reserve(wstride * wh);
    const size_t o_top = st.reserve((size_t)(T + 1) * sizeof(pixel));
    const size_t o_edge = st.reserve((size_t)span * sizeof(pixel));
    const size_t o_res = st.reserve(8);
    if (!st.commit()) return mode;
    pixel *win = (pixel *)st.host(o_win);
    memset(win, 0, wstride * wh);
    if (have_top && !top_sb_edge)
        for (int i = have_left ? -1 : 0; i < T; i++) win[1 + i] = dst[-spx + i];
    for (int i = 0; i < L; i++) win[(size_t)(1 + i) * ww] = dst[spx * i - 1];
    pixel *tp = (pixel *)st.host(o_top);
    memset(tp, 0, (size_t)(T + 1) * sizeof(pixel));
    if (have_top && top_sb_edge)
        for (int i = have_left ? -1 : 0; i < T; i++) tp[1 + i] = top_sb_edge[x * 4 + i];
    if (!st.upload()) return mode;
    PrepArgs a;
    a.x = x; a.have_left = have_left; a.y = y; a.have_top = have_top; a.w = w; a.h = h; a.edge_flags = edge_flags;
    a.dst = (const pixel *)st.dev(o_win) + ww + 1; a.stride = ww;
    a.top_sb_edge = top_sb_edge ? (const void *)((const pixel *)st.dev(o_top) + 1 - x * 4) : nullptr;
    a.mode = mode; a.angle = *angle; a.tw = tw; a.th = th; a.filter_edge = filter_edge; a.bdmax = bdmax;
    a.edge_out = (pixel *)st.dev(o_edge) + 2 * th * 4;
    a.result = (int *)st.dev(o_res);
    prepare_edges_kernel<pixel><<<1, 32, 0, st.stream()>>>(a);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "prepare_edges_kernel")) return mode;
    if (!st.download(o_edge, (o_res - o_edge) + 8) || !st.sync()) return mode;
    const int *res = (const int *)st.host(o_res);
    const int m = res[0];
    *angle = res[1];
    // copy back only the parts of the edge array the returned mode defines
    // (ipred_prepare_tmpl.c:50-74 needs_* table), the rest of the caller's buffer is untouched
    static const uint8_t needs_tbl[14] = { 3, 2, 1, 1, 2, 0, 2 | 8 | 4, 1 | 2 | 4, 1 | 16 | 4, 3, 3, 3, 7, 7 };
    const int needs = needs_tbl[m];
    const pixel *e = (const pixel *)st.host(o_edge) + 2 * th * 4;
    const int szl = th * 4, szt = tw * 4;
    if (needs & 1) memcpy(topleft_out - szl, e - szl, (size_t)szl * sizeof(pixel));
    if (needs & 16) memcpy(topleft_out - 2 * szl, e - 2 * szl, (size_t)szl * sizeof(pixel));
    if (needs & 2) memcpy(topleft_out + 1, e + 1, (size_t)szt * sizeof(pixel));
    if (needs & 8) memcpy(topleft_out + 1 + szt, e + 1 + szt, (size_t)szt * sizeof(pixel));
    if (needs & 4) topleft_out[0] = e[0];
    return m;
}

#define HBD_ARGS , int bitdepth_max
template <int M> static void ip8(uint8_t *d, ptrdiff_t s, const uint8_t *tl, int w, int h, int a, int mw, int mh)
{ ipred_single<uint8_t>(M, d, s, tl, w, h, a, mw, mh, 0xff, nullptr, 0, false); }
template <int M> static void ip16(uint16_t *d, ptrdiff_t s, const uint16_t *tl, int w, int h, int a, int mw, int mh HBD_ARGS)
{ ipred_single<uint16_t>(M, d, s, tl, w, h, a, mw, mh, bitdepth_max, nullptr, 0, false); }
template <int M> static void cfl8(uint8_t *d, ptrdiff_t s, const uint8_t *tl, int w, int h, const int16_t *ac, int alpha)
{ ipred_single<uint8_t>(M, d, s, tl, w, h, 0, 0, 0, 0xff, ac, alpha, true); }
template <int M> static void cfl16(uint16_t *d, ptrdiff_t s, const uint16_t *tl, int w, int h, const int16_t *ac, int alpha HBD_ARGS)
{ ipred_single<uint16_t>(M, d, s, tl, w, h, 0, 0, 0, bitdepth_max, ac, alpha, true); }
template <typename pixel, int SH, int SV>
static void ac_p(int16_t *ac, const pixel *y, ptrdiff_t stride, int w_pad, int h_pad, int cw, int ch)
{ cfl_ac_single<pixel>(ac, y, stride, w_pad, h_pad, cw, ch, SH, SV); }
template <typename pixel>
static void pal_p(pixel *d, ptrdiff_t s, const pixel *pal, const uint8_t *idx, int w, int h)
{ pal_pred_single<pixel>(d, s, pal, idx, w, h); }

template <bool HBD, int M> static void fill_mode(Dav1dCudaIntraPredDSPContext *c) {
    c->intra_pred[M] = HBD ? (void *)ip16<M> : (void *)ip8<M>;
}

template <bool HBD> static void fill_ipred(Dav1dCudaIntraPredDSPContext *c) {
    Staging &s = staging();
    {
        std::lock_guard<std::mutex> lk(s.mu);
        if (!s.ensure(1 << 20)) return;
    }
    typedef typename std::conditional<HBD, uint16_t, uint8_t>::type pixel;
    fill_mode<HBD, 0>(c); fill_mode<HBD, 1>(c); fill_mode<HBD, 2>(c); fill_mode<HBD, 3>(c);
    fill_mode<HBD, 4>(c); fill_mode<HBD, 5>(c); fill_mode<HBD, 6>(c); fill_mode<HBD, 7>(c);
    fill_mode<HBD, 8>(c); fill_mode<HBD, 9>(c); fill_mode<HBD, 10>(c); fill_mode<HBD, 11>(c);
    fill_mode<HBD, 12>(c); fill_mode<HBD, 13>(c);
    c->cfl_ac[0] = (void *)ac_p<pixel, 1, 1>;   // 420
    c->cfl_ac[1] = (void *)ac_p<pixel, 1, 0>;   // 422
    c->cfl_ac[2] = (void *)ac_p<pixel, 0, 0>;   // 444
    c->cfl_pred[M_DC] = HBD ? (void *)cfl16<M_DC> : (void *)cfl8<M_DC>;
    c->cfl_pred[M_LEFT_DC] = HBD ? (void *)cfl16<M_LEFT_DC> : (void *)cfl8<M_LEFT_DC>;
    c->cfl_pred[M_TOP_DC] = HBD ? (void *)cfl16<M_TOP_DC> : (void *)cfl8<M_TOP_DC>;
    c->cfl_pred[M_DC_128] = HBD ? (void *)cfl16<M_DC_128> : (void *)cfl8<M_DC_128>;
    c->pal_pred = (void *)pal_p<pixel>;
}

}  // namespace d1

using namespace d1;

extern "C" {

void dav1d_cuda_intra_pred_dsp_init_8bpc(Dav1dCudaIntraPredDSPContext *c) { fill_ipred<false>(c); }
void dav1d_cuda_intra_pred_dsp_init_16bpc(Dav1dCudaIntraPredDSPContext *c) { fill_ipred<true>(c); }

int dav1d_cuda_prepare_intra_edges_8bpc(int x, int have_left, int y, int have_top, int w, int h, int edge_flags,
                                        const uint8_t *dst, ptrdiff_t stride, const uint8_t *top_sb_edge, int mode,
                                        int *angle, int tw, int th, int filter_edge, uint8_t *topleft_out)
{
    return prepare_single<uint8_t>(x, have_left, y, have_top, w, h, edge_flags, dst, stride, top_sb_edge, mode,
                                   angle, tw, th, filter_edge, topleft_out, 0xff);
}
int dav1d_cuda_prepare_intra_edges_16bpc(int x, int have_left, int y, int have_top, int w, int h, int edge_flags,
                                         const uint16_t *dst, ptrdiff_t stride, const uint16_t *top_sb_edge, int mode,
                                         int *angle, int tw, int th, int filter_edge, uint16_t *topleft_out,
                                         int bitdepth_max)
{
    return prepare_single<uint16_t>(x, have_left, y, have_top, w, h, edge_flags, dst, stride, top_sb_edge, mode,
                                    angle, tw, th, filter_edge, topleft_out, bitdepth_max);
}

}  // extern "C"
