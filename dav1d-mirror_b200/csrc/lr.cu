// Loop restoration (Wiener / self-guided) of a whole frame on the device, third of the in-loop post-filters
// (SURVEY 8f-2).
//
// Reference: dav1d_lr_sbrow (src/lr_apply_tmpl.c:165-202) -> lr_sbrow (:107-163: the restoration unit of a
// position, units of unit_size with a last one of up to 1.5x) -> lr_stripe (:36-97: 64-row stripes shifted up
// by 8 luma rows, filter set-up) -> dsp->lr.wiener[] / .sgr[] (src/looprestoration_tmpl.c: padding() :41-125,
// wiener_c :131-189, selfguided_filter :344-446, sgr_5x5 / 3x3 / mix :448-519).
//
// What the reference's in-place filter reads, per stripe: inside the stripe the CDEF output (pixels of
// neighbouring units before THEY are restored: its `left` backups), and for the two rows above / below the
// stripe the DEBLOCKED, pre-CDEF lines that dav1d_copy_lpf saved (f->lf.lr_lpf_line), the outer one repeated
// for the third row; at the frame's borders rows / columns are replicated.  With the CDEF output and the
// deblocked picture both in HBM this is an out-of-place filter of independent tiles: one CTA per 32-pixel-wide
// column of a stripe of a plane.
#include "ctx.h"
#include "common.cuh"

namespace d1 {

constexpr int AV1RESTORATION_BYTES = 108;     // sizeof(Av1Restoration): lr[3][4] of 9-byte units (src/lf_mask.h:40-46, 61-63)
constexpr int LR_TW = 32, LR_MAXH = 64;
constexpr int LR_TS = LR_TW + 6;              // tile stride
constexpr int LR_AS = LR_TW + 2;              // A / B stride

struct LrArgs {
    PlaneView src[3], pre[3], dst[3];
    int bdmax, ss_hor, ss_ver;
    int w, h, sb128w;
    int unit_log2[2];
    int restore_planes;
    const uint8_t *lr_mask;
};

// Sgr_Params of the specification (dav1d_sgr_params, src/tables.c:415-420): s0 (5x5), s1 (3x3) per set
__device__ const uint16_t g_sgr_params[16][2] = {
    { 140, 3236 }, { 112, 2158 }, { 93, 1618 }, { 80, 1438 }, { 70, 1295 }, { 58, 1177 }, { 47, 1079 }, { 37, 996 },
    { 30, 925 }, { 25, 863 }, { 0, 2589 }, { 0, 1618 }, { 0, 1177 }, { 0, 925 }, { 56, 0 }, { 22, 0 },
};

// dav1d_sgr_x_by_x[z] (src/tables.c:422-441) = round(256 / (z + 1)), at most 255, 0 for z = 255
DEV unsigned sgr_x_by_x(const unsigned z) {
    if (z >= 255) return 0;
    const unsigned v = (512 + z + 1) / (2 * z + 2);
    return v > 255 ? 255 : v;
}

// A and B of selfguided_filter() (:363-381) at the positions the filter reads: box sums of radius 1 (n = 9) or
// 2 (n = 25) around (i, j), i in [-1, tw], j in [-1, th] (n = 25: every second row from -1)
template <int n>
DEV void sgr_ab(const uint16_t *tile, int *A, int *B, const int tw, const int th, const unsigned s,
                const int bdm8, const int tid, const int nthr, const uint8_t *x_by_x)
{
    constexpr int r = n == 25 ? 2 : 1, step = n == 25 ? 2 : 1;
    constexpr unsigned one_by_x = n == 25 ? 164 : 455;
    const int rows = (th + 2 + step - 1) / step, cols = tw + 2;
    // (row, column) of k = tid + m * nthr without a division per step
    const int dq = nthr / cols, dr = nthr % cols;
    int jr = tid / cols, i0 = tid % cols;
    for (; jr < rows; jr += dq, i0 += dr) {
        if (i0 >= cols) { i0 -= cols; jr++; if (jr >= rows) break; }
        const int j = jr * step - 1, i = i0 - 1;
        int sum = 0, sumsq = 0;
#pragma unroll
        for (int dy = -r; dy <= r; dy++)
#pragma unroll
            for (int dx = -r; dx <= r; dx++) {
                const int v = tile[(j + 3 + dy) * LR_TS + i + 3 + dx];
                sum += v; sumsq += v * v;
            }
        const int a = (sumsq + ((1 << (2 * bdm8)) >> 1)) >> (2 * bdm8);
        const int b = (sum + ((1 << bdm8) >> 1)) >> bdm8;
        const unsigned p = (unsigned)imax(a * n - b * b, 0);
        const unsigned z = (p * s + (1u << 19)) >> 20;
        const unsigned x = x_by_x[z > 255 ? 255 : z];
        A[(j + 1) * LR_AS + i + 1] = (int)((x * (unsigned)sum * one_by_x + (1u << 11)) >> 12);
        B[(j + 1) * LR_AS + i + 1] = (int)x;
    }
}

// the filter output of one pixel (:383-445); A / B indexed with the +1 offsets of sgr_ab
DEV int sgr_px(const int *A, const int *B, const int i, const int j, const int n, const int src) {
    const int *a0 = A + (j + 1) * LR_AS + i + 1, *b0 = B + (j + 1) * LR_AS + i + 1;
    if (n == 25) {
        if (!(j & 1)) {
            const int a = (b0[-LR_AS] + b0[LR_AS]) * 6 + (b0[-1 - LR_AS] + b0[-1 + LR_AS] + b0[1 - LR_AS] + b0[1 + LR_AS]) * 5;
            const int b = (a0[-LR_AS] + a0[LR_AS]) * 6 + (a0[-1 - LR_AS] + a0[-1 + LR_AS] + a0[1 - LR_AS] + a0[1 + LR_AS]) * 5;
            return (b - a * src + (1 << 8)) >> 9;
        }
        const int a = b0[0] * 6 + (b0[-1] + b0[1]) * 5;
        const int b = a0[0] * 6 + (a0[-1] + a0[1]) * 5;
        return (b - a * src + (1 << 7)) >> 8;
    }
    const int a = (b0[0] + b0[-1] + b0[1] + b0[-LR_AS] + b0[LR_AS]) * 4 +
                  (b0[-1 - LR_AS] + b0[-1 + LR_AS] + b0[1 - LR_AS] + b0[1 + LR_AS]) * 3;
    const int b = (a0[0] + a0[-1] + a0[1] + a0[-LR_AS] + a0[LR_AS]) * 4 +
                  (a0[-1 - LR_AS] + a0[-1 + LR_AS] + a0[1 - LR_AS] + a0[1 + LR_AS]) * 3;
    return (b - a * src + (1 << 8)) >> 9;
}

template <typename pixel>
__global__ void __launch_bounds__(256) lr_kernel(const __grid_constant__ LrArgs a) {
    __shared__ uint16_t tile[(LR_MAXH + 6) * LR_TS];
    __shared__ int AB[2][(LR_MAXH + 2) * LR_AS];       // A / B of the self-guided filter; `hor` of the Wiener filter
    __shared__ uint8_t x_by_x[256];
    const int pl = blockIdx.z, tid = threadIdx.x;
    const int ssh = pl ? a.ss_hor : 0, ssv = pl ? a.ss_ver : 0;
    const int pw = (a.w + ssh) >> ssh, ph = (a.h + ssv) >> ssv;
    const int x0 = blockIdx.x * LR_TW;
    const int sh = 64 >> ssv, off = 8 >> ssv;
    const int s = blockIdx.y;
    const int y0 = s ? s * sh - off : 0;
    const int y1 = imin((s + 1) * sh - off, ph);
    if (x0 >= pw || y0 >= ph) return;
    const int tw = imin(LR_TW, pw - x0), th = y1 - y0;
    const int64_t sstride = a.src[pl].stride / (int64_t)sizeof(pixel), pstride = a.pre[pl].stride / (int64_t)sizeof(pixel);
    const int64_t dstride = a.dst[pl].stride / (int64_t)sizeof(pixel);
    const pixel *src = (const pixel *)a.src[pl].data, *pre = (const pixel *)a.pre[pl].data;
    pixel *dst = (pixel *)a.dst[pl].data;

    // the restoration unit of this tile (lr_sbrow, lr_apply_tmpl.c:121-163)
    int type = 0;
    const uint8_t *u = nullptr;
    if ((a.restore_planes >> pl) & 1) {
        const int unit = 1 << a.unit_log2[!!pl], half = unit >> 1;
        const int row_y = s * sh;
        int aligned = row_y & ~(unit - 1);
        if (aligned && aligned + half > ph) aligned -= unit;
        aligned <<= ssv;
        const int sb_idx = (aligned >> 7) * a.sb128w, unit_idx = ((aligned >> 6) & 1) << 1;
        const int n_full = pw >= half ? (pw - half) >> a.unit_log2[!!pl] : 0;     // units before the last, which takes up to 1.5 units
        const int xu = imin(x0 / unit, n_full) * unit;
        const int shift_hor = 7 - ssh;
        u = a.lr_mask + (size_t)(sb_idx + (xu >> shift_hor)) * AV1RESTORATION_BYTES +
            (pl * 4 + unit_idx + ((xu >> (shift_hor - 1)) & 1)) * 9;
        type = u[0];
    }
    if (!type) {                                       // DAV1D_RESTORATION_NONE: the CDEF output as it is
        for (int k = tid; k < tw * th; k += 256) {
            const int y = k / tw, x = k % tw;
            dst[(int64_t)(y0 + y) * dstride + x0 + x] = src[(int64_t)(y0 + y) * sstride + x0 + x];
        }
        return;
    }
    // padding() (looprestoration_tmpl.c:41-125): rows -3 .. th + 2, columns -3 .. tw + 2
    const bool have_top = y0 > 0, have_bottom = y1 < ph;
    const int tcols = tw + 6, tdq = 256 / tcols, tdr = 256 % tcols;
    for (int rr = tid / tcols, cc = tid % tcols; rr < th + 6; rr += tdq, cc += tdr) {
        if (cc >= tcols) { cc -= tcols; rr++; if (rr >= th + 6) break; }
        const int r = rr - 3, c = cc - 3;
        const int xx = iclip(x0 + c, 0, pw - 1);
        int v;
        if (r < 0) v = have_top ? pre[(int64_t)(y0 - 2 + (r == -1)) * pstride + xx] : src[(int64_t)y0 * sstride + xx];
        else if (r >= th) v = have_bottom ? pre[(int64_t)imin(y1 + (r > th), ph - 1) * pstride + xx] : src[(int64_t)(y1 - 1) * sstride + xx];
        else v = src[(int64_t)(y0 + r) * sstride + xx];
        tile[(r + 3) * LR_TS + c + 3] = (uint16_t)v;
    }
    __syncthreads();
    const int bitdepth = PxTraits<pixel>::bitdepth(a.bdmax);
    if (type > 2) x_by_x[tid] = (uint8_t)sgr_x_by_x(tid);      // visible after the barrier before sgr_ab
    if (type == 2) {
        // DAV1D_RESTORATION_WIENER (enum Dav1dRestorationType: NONE 0, SWITCHABLE 1, WIENER 2, SGRPROJ 3) (lr_stripe :54-72 + wiener_c :131-189)
        int fh[7], fv[7];
        for (int k = 0; k < 3; k++) {
            fh[k] = fh[6 - k] = (int8_t)u[1 + k];
            fv[k] = fv[6 - k] = (int8_t)u[4 + k];
        }
        fh[3] = -(fh[0] + fh[1] + fh[2]) * 2 + 128;    // the centre tap's 128 (added apart at 8 bit, :158-160)
        fv[3] = 128 - (fv[0] + fv[1] + fv[2]) * 2;
        const int rbh = 3 + (bitdepth == 12) * 2, clip_limit = 1 << (bitdepth + 1 + 7 - rbh);
        uint16_t *hor = (uint16_t *)AB;                // (th + 6) x tw, stride LR_TW
        for (int k = tid; k < (th + 6) * tw; k += 256) {
            const int j = tw == LR_TW ? k >> 5 : k / tw, i = tw == LR_TW ? k & 31 : k % tw;
            int sum = 1 << (bitdepth + 6);
#pragma unroll
            for (int t = 0; t < 7; t++) sum += tile[j * LR_TS + i + t] * fh[t];
            hor[j * LR_TW + i] = (uint16_t)iclip((sum + (1 << (rbh - 1))) >> rbh, 0, clip_limit - 1);
        }
        __syncthreads();
        const int rbv = 11 - (bitdepth == 12) * 2, round_offset = 1 << (bitdepth + rbv - 1);
        for (int k = tid; k < th * tw; k += 256) {
            const int j = tw == LR_TW ? k >> 5 : k / tw, i = tw == LR_TW ? k & 31 : k % tw;
            int sum = -round_offset;
#pragma unroll
            for (int t = 0; t < 7; t++) sum += hor[(j + t) * LR_TW + i] * fv[t];
            dst[(int64_t)(y0 + j) * dstride + x0 + i] = (pixel)clip_px<pixel>((sum + (1 << (rbv - 1))) >> rbv, a.bdmax);
        }
        return;
    }
    // DAV1D_RESTORATION_SGRPROJ + set (lr_stripe :73-83 + sgr_5x5 / 3x3 / mix :448-519)
    const int set = (type - 3) & 15;
    const unsigned s0 = g_sgr_params[set][0], s1 = g_sgr_params[set][1];
    const int w0 = (int8_t)u[7], w1 = 128 - ((int8_t)u[7] + (int8_t)u[8]);
    const int bdm8 = bitdepth - 8;
    int acc[8];                                        // w0 * dst0 + w1 * dst1 of this thread's pixels
#pragma unroll
    for (int q = 0; q < 8; q++) acc[q] = 0;
    for (int pass = 0; pass < 2; pass++) {
        const unsigned sp = pass ? s1 : s0;
        if (!sp) continue;
        const int n = pass ? 9 : 25, wt = pass ? w1 : w0;
        __syncthreads();
        if (pass) sgr_ab<9>(tile, AB[0], AB[1], tw, th, sp, bdm8, tid, 256, x_by_x);
        else sgr_ab<25>(tile, AB[0], AB[1], tw, th, sp, bdm8, tid, 256, x_by_x);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int k = tid + q * 256;
            if (k < tw * th) {
                const int j = tw == LR_TW ? k >> 5 : k / tw, i = tw == LR_TW ? k & 31 : k % tw;
                acc[q] += wt * sgr_px(AB[0], AB[1], i, j, n, tile[(j + 3) * LR_TS + i + 3]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int k = tid + q * 256;
        if (k < tw * th) {
            const int j = tw == LR_TW ? k >> 5 : k / tw, i = tw == LR_TW ? k & 31 : k % tw;
            const int px = tile[(j + 3) * LR_TS + i + 3];
            dst[(int64_t)(y0 + j) * dstride + x0 + i] = (pixel)clip_px<pixel>(px + ((acc[q] + (1 << 10)) >> 11), a.bdmax);
        }
    }
}

}  // namespace d1

using namespace d1;

extern "C" int dav1d_cuda_lr_frame(Dav1dCudaContext *c, const Dav1dCudaPicture *dst, const Dav1dCudaPicture *src,
                                   const Dav1dCudaPicture *deblocked, const Dav1dCudaLrFrame *p)
{
    if (!c || !dst || !src || !deblocked || !p || !p->lr_mask || !dst->p[0].data || !src->p[0].data ||
        !deblocked->p[0].data || dst->p[0].data == src->p[0].data || dst->p[0].data == deblocked->p[0].data ||
        p->w <= 0 || p->h <= 0 || p->sb128w < (p->w + 127) / 128) return -22;
    for (int k = 0; k < 2; k++)
        if (p->unit_size_log2[k] < 5 || p->unit_size_log2[k] > 8) return -22;
    // 128x128 superblocks: lr_sbrow (lr_apply_tmpl.c:137-143) looks the unit of a whole 128-row superblock row up
    // at its first row.  The frame header gives such streams units of at least 128 luma pixels
    // (obu.c:944-954: unit_size[0] = 6 + sb128 ..., unit_size[1] one less only when both directions are
    // subsampled), and then every 64-row stripe finds the same unit at its own first row: the kernel's
    // lookup is the reference's.  Smaller units with sb128 are not a stream the reference can be handed.
    if (p->sb128 && (p->unit_size_log2[0] < 7 || p->unit_size_log2[1] < 7 - (src->ss_ver ? 1 : 0))) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    LrArgs a;
    const PicView s = pic_view(src), q = pic_view(deblocked), d = pic_view(dst);
    for (int i = 0; i < 3; i++) { a.src[i] = s.p[i]; a.pre[i] = q.p[i]; a.dst[i] = d.p[i]; }
    a.bdmax = s.bdmax; a.ss_hor = s.ss_hor; a.ss_ver = s.ss_ver;
    a.w = p->w; a.h = p->h; a.sb128w = p->sb128w;
    a.unit_log2[0] = p->unit_size_log2[0]; a.unit_log2[1] = p->unit_size_log2[1];
    a.restore_planes = p->restore_planes;
    a.lr_mask = (const uint8_t *)p->lr_mask;
    const int n_planes = (src->p[1].data && src->p[2].data && dst->p[1].data && dst->p[2].data) ? 3 : 1;
    const dim3 grid((unsigned)((p->w + LR_TW - 1) / LR_TW), (unsigned)((p->h + 8 + 63) / 64), (unsigned)n_planes);
    if (s.bdmax > 0xff) lr_kernel<uint16_t><<<grid, 256, 0, c->stream>>>(a);
    else lr_kernel<uint8_t><<<grid, 256, 0, c->stream>>>(a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "lr_kernel") ? 0 : -5;
}
