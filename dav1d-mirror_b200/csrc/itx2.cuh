// Compact 2-D inverse transform + add ("itx2"): every 1-D transform exists ONCE per pixel type as an
// out-of-line function that works on a shared-memory tile with run-time strides, and is used by
// the row pass and the column pass of all 19 transform sizes, by the inter-residual task kernel
// and by the intra executor alike.  The straight-line per-(size, type, box) expansions of
// itx.cuh (1.1 MB of SASS for the large sizes) made instruction fetch the top stall of every
// transform kernel; here the whole transform code is a few tens of KB and the hot part
// (4/8/16-point) stays in the instruction caches whatever mix of blocks an SM is running.
//
// Arithmetic: inv_txfm_add_c (reference src/itx_tmpl.c:40-100) step for step - dc-only shortcut
// (:53-65), rect2 pre-scaling (:80-82), per-size inter-pass shift (:142-160), bit-depth dependent
// clamps (:68-76), WHT 4x4 (:166-185); the 1-D butterflies are those of itx_1d.cuh
// (src/itx_1d.c).  Coefficients come as the non-zero bounding box (cw columns x ch rows,
// multiples of 4, column-major with stride ch; the dense reference layout is the box
// min(w,32) x min(h,32)); inputs beyond the box are literal zeros, which selects the reduced
// 8- / 16-input variants of the long DCTs (exact: every stage maps 0 -> 0).
//
// Everything here is __host__ __device__: tests/host/itx2_check.cpp runs the same
// code lane by lane on the CPU against the reference's itxfm_add for all 156 slots.
#pragma once
#include "itx_geom.cuh"

namespace d1 {

#if defined(__CUDA_ARCH__)
#define D1_ITX2_FN __device__ __noinline__
#else
#define D1_ITX2_FN inline
#endif

// One N-point transform of kind KIND (K_DCT / K_ADST / K_WHT) whose inputs beyond NZ are zero:
// c[k] = p[k * istride], transform, q[k * ostride] = cl2((c[k] + rnd) >> sh).  Works in place
// (q == p) and, with a reversed output (negative ostride), gives FLIPADST.
template <int N, int KIND, int NZ>
D1_ITX2_FN void itx2_vec(const int *p, const int istride, int *q, const int ostride, const Clamp cl, const int rnd,
                         const int sh, const Clamp cl2)
{
    int c[N];
#pragma unroll
    for (int k = 0; k < NZ; k++) c[k] = p[k * istride];
#pragma unroll
    for (int k = NZ; k < N; k++) c[k] = 0;
    if (KIND == K_DCT) {
        if (N == 4) idct4<1, false>(c, cl);
        else if (N == 8) idct8<1, false>(c, cl);
        else if (N == 16) idct16<1, false>(c, cl);
        else if (N == 32) idct32<1, false>(c, cl);
        else idct64<1>(c, cl);
    } else if (KIND == K_ADST) {
        if (N == 4) iadst4<1, false>(c, cl);
        else if (N == 8) iadst8<1, false>(c, cl);
        else iadst16<1, false>(c, cl);
    } else {
        iwht4<1>(c);
    }
#pragma unroll
    for (int k = 0; k < N; k++) q[k * ostride] = cl2((c[k] + rnd) >> sh);
}

// identity: element-wise, run-time length (itx_1d.c:983-1017):
//   4: v + (v*1697 + 2048 >> 12)   8: 2v   16: 2v + (v*1697 + 1024 >> 11)   32: 4v
HD void itx2_identity(const int n, const int nz, const int *p, const int istride, int *q, const int ostride,
                      const int rnd, const int sh, const Clamp cl2)
{
    const int m = n == 4 ? 1 : n == 32 ? 4 : 2;
    const int fm = (n == 4 || n == 16) ? 1697 : 0, fr = n == 4 ? 2048 : 1024, fs = n == 4 ? 12 : 11;
#pragma unroll 2
    for (int k = 0; k < n; k++) {
        int v = k < nz ? p[k * istride] : 0;
        v = v * m + ((v * fm + fr) >> fs);
        q[k * ostride] = cl2((v + rnd) >> sh);
    }
}

// inputs the selected variant reads (zero-filled by the caller up to here)
HD int itx2_nzb(const int n, const int kind, const int nz) {
    if (kind == K_IDENTITY) return nz;
    const int nin = n == 64 ? 32 : n;
    if (kind != K_DCT || nin <= 8) return nin;
    if (nz <= 8) return 8;
    if (nin > 16 && nz <= 16) return 16;
    return nin;
}

// One vector: n-point transform `kind` (enum Itx1d; FLIPADST = ADST with the output order reversed).
// MAXN: longest transform the calling kernel can meet (prunes the long DCTs, and with them their
// registers, from the kernel that only runs sizes up to 16x16).
template <int MAXN = 64>
HD void itx2_run(const int n, const int kind, const int nz, const int *p, const int istride, int *q, int ostride,
                 const Clamp cl, const int rnd, const int sh, const Clamp cl2)
{
    if (kind == K_FLIPADST) { q += (n - 1) * ostride; ostride = -ostride; }
#define D1_V(N, K, NZ) itx2_vec<N, K, NZ>(p, istride, q, ostride, cl, rnd, sh, cl2)
    if (kind == K_IDENTITY) {
        itx2_identity(n, nz, p, istride, q, ostride, rnd, sh, cl2);
    } else if (kind == K_DCT) {
        const int b = itx2_nzb(n, K_DCT, nz);
        switch (n) {
        case 4: D1_V(4, K_DCT, 4); break;
        case 8: D1_V(8, K_DCT, 8); break;
        case 16: if (b == 8) D1_V(16, K_DCT, 8); else D1_V(16, K_DCT, 16); break;
        case 32:
            if (MAXN >= 32) { if (b == 8) D1_V(32, K_DCT, 8); else if (b == 16) D1_V(32, K_DCT, 16); else D1_V(32, K_DCT, 32); }
            break;
        default:
            if (MAXN >= 64) { if (b == 8) D1_V(64, K_DCT, 8); else if (b == 16) D1_V(64, K_DCT, 16); else D1_V(64, K_DCT, 32); }
            break;
        }
    } else if (kind == K_WHT) {
        D1_V(4, K_WHT, 4);
    } else {
        switch (n) {
        case 4: D1_V(4, K_ADST, 4); break;
        case 8: D1_V(8, K_ADST, 8); break;
        default: D1_V(16, K_ADST, 16); break;
        }
    }
#undef D1_V
}

// Geometry + clamps of one block, computed once per block by every lane of its group.
struct Itx2Blk {
    int w, h, sw, sh, ts, shift, lw;   // lw = log2(w)
    bool rect2, wht, dc_only;
    int rk, ck;                 // row / column 1-D kind
    int cw, ch;                 // stored coefficient box
    int rin, cin;               // inputs the row / column variant reads
    Clamp rowcl, colcl;
};

template <typename pixel>
HD Itx2Blk itx2_setup(const int tx, const int txtp, const int eob, const int cw4, const int ch4, const int bdmax) {
    Itx2Blk b;
    const TxDim t = tx_dim(tx);
    b.w = t.w; b.h = t.h; b.shift = t.shift;
    b.sw = b.w < 32 ? b.w : 32; b.sh = b.h < 32 ? b.h : 32;
    b.ts = b.w + 1;
    b.lw = b.w == 4 ? 2 : b.w == 8 ? 3 : b.w == 16 ? 4 : b.w == 32 ? 5 : 6;
    b.rect2 = b.w * 2 == b.h || b.h * 2 == b.w;
    b.wht = txtp == 16;
    b.dc_only = eob == 0 && txtp == 0;
    b.rk = txtp_row_kind(txtp); b.ck = txtp_col_kind(txtp);
    b.cw = cw4 ? cw4 * 4 : b.sw; b.ch = ch4 ? ch4 * 4 : b.sh;
    b.rin = itx2_nzb(b.w, b.rk, b.cw);
    b.cin = itx2_nzb(b.h, b.ck, b.ch);
    if (PxTraits<pixel>::hbd) {
        b.rowcl.lo = (int)((unsigned)~bdmax << 7);
        b.colcl.lo = (int)((unsigned)~bdmax << 5);
    } else {
        b.rowcl.lo = b.colcl.lo = -32768;
    }
    b.rowcl.hi = ~b.rowcl.lo; b.colcl.hi = ~b.colcl.lo;
    return b;
}

HD int itx2_tile_ints(const int tx) {          // shared-memory ints one block of size tx needs
    const TxDim t = tx_dim(tx);
    return t.h * (t.w + 1);
}

// ---- the phases of one block; lane gl of a group of G lanes (G >= max(sw, sh), power of two).
// Between two phases the group synchronises (itx2_block below; the host check runs each phase
// for all lanes in turn).

// ---- where a block's coefficients come from: a plain pointer into the stream (int16 at 8 bit, int32 at
// high bit depth: what the reference stores), or - Dav1dCudaReconBatch.cf_int16 - the compact high-bit-depth
// stream: int16 storage, the value -32768 standing for "look the coefficient up in the escape list"
struct Itx2Esc { uint32_t off; int32_t value; };     // == Dav1dCudaCoefEsc (this header is also compiled host-only)
struct Itx2Coef {
    void *p;                            // the block's first coefficient
    const Itx2Esc *esc;                 // escapes of the frame, sorted by stream offset
    int n_esc;
    uint32_t off;                       // stream offset of the block (key into `esc`)
    int s16;                            // 0: int32 storage
};
HD int cf_get(const int16_t *p, const int i) { return p[i]; }
HD int cf_get(const int32_t *p, const int i) { return p[i]; }
HD void cf_zero(int16_t *p, const int i) { p[i] = 0; }
HD void cf_zero(int32_t *p, const int i) { p[i] = 0; }
HD int cf_escape(const Itx2Coef &c, const int i) {
    const uint32_t key = c.off + (uint32_t)i;
    int lo = 0, hi = c.n_esc - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t o = c.esc[mid].off;
        if (o == key) return c.esc[mid].value;
        if (o < key) lo = mid + 1; else hi = mid - 1;
    }
    return -32768;                      // not listed: the value itself
}
HD int cf_get(const Itx2Coef &c, const int i) {
    if (!c.s16) return ((const int32_t *)c.p)[i];
    const int v = ((const int16_t *)c.p)[i];
    return v == -32768 ? cf_escape(c, i) : v;
}
HD void cf_zero(const Itx2Coef &c, const int i) {
    if (c.s16) ((int16_t *)c.p)[i] = 0; else ((int32_t *)c.p)[i] = 0;
}

// dc-only (itx_tmpl.c:53-65): every pixel gets the same offset
template <typename pixel, typename CF = const typename PxTraits<pixel>::coef *>
HD int itx2_dc_value(const Itx2Blk &b, const CF cf) {
    int dc = cf_get(cf, 0);
    if (b.rect2) dc = (dc * 181 + 128) >> 8;
    dc = (dc * 181 + 128) >> 8;
    dc = (dc + ((1 << b.shift) >> 1)) >> b.shift;
    return (dc * 181 + 128 + 2048) >> 12;
}

// Last phase: residual r(x, y) = dc (dc-only blocks) or the column pass's output in the tile ->
//   res == nullptr: dst = clip_px(dst + r)          (inter residuals: read-modify-write)
//   res != nullptr: res = saturate_int16(r)         (intra residuals, added by the intra executor
//                   to the prediction: clip_px(pred + r) is the same for the saturated r)
// Four pixels per lane and step, four steps' loads in flight.
template <typename pixel>
HD void itx2_phase_out(const Itx2Blk &b, const int gl, const int G, const int *tile, const int dc, pixel *dst,
                       const int dstride, int16_t *res, const int rstride, const int bdmax)
{
    const int ls = b.lw - 2, nv = b.h << ls;          // 4-pixel segments per row: w / 4
    for (int i0 = gl; i0 < nv; i0 += 4 * G) {
        int v[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * G;
            if (i < nv && !res) {
                const int y = i >> ls, x = (i - (y << ls)) * 4;
#if defined(__CUDA_ARCH__)
                load_px<pixel, 4>(dst + y * dstride + x, v[u]);
#else
                for (int k = 0; k < 4; k++) v[u][k] = dst[y * dstride + x + k];
#endif
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * G;
            if (i < nv) {
                const int y = i >> ls, x = (i - (y << ls)) * 4;
                int r[4];
#pragma unroll
                for (int k = 0; k < 4; k++) r[k] = b.dc_only ? dc : tile[y * b.ts + x + k];
                if (res) {
                    int16_t *rp = res + y * rstride + x;
#pragma unroll
                    for (int k = 0; k < 4; k++) rp[k] = (int16_t)(r[k] < -32768 ? -32768 : r[k] > 32767 ? 32767 : r[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) v[u][k] = clip_px<pixel>(v[u][k] + r[k], bdmax);
#if defined(__CUDA_ARCH__)
                    store_px<pixel, 4>(dst + y * dstride + x, v[u]);
#else
                    for (int k = 0; k < 4; k++) dst[y * dstride + x + k] = (pixel)v[u][k];
#endif
                }
            }
        }
    }
}

// coefficient box -> tile (row-major, stride ts), zeros where the passes read beyond the box
template <typename pixel, typename CF = typename PxTraits<pixel>::coef *>
HD void itx2_phase_stage(const Itx2Blk &b, const int gl, const int G, const CF cf,
                         int *tile, const bool zero_coefs)
{
    // G >= ch: a lane owns one row of the box and walks along the columns (consecutive lanes read
    // consecutive coefficients of a column)
    if (gl < b.ch) {
        // eight columns per step: their loads are in flight together (cw is a multiple of 4)
        for (int x0 = 0; x0 < b.cw; x0 += 8) {
            int v[8];
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (k < 4 || x0 + k < b.cw) v[k] = cf_get(cf, gl + (x0 + k) * b.ch);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k < 4 || x0 + k < b.cw) {
                    int t = v[k];
                    if (zero_coefs) cf_zero(cf, gl + (x0 + k) * b.ch);
                    if (b.wht) t >>= 2;
                    else if (b.rect2) t = (t * 181 + 128) >> 8;
                    tile[gl * b.ts + x0 + k] = t;
                }
            }
        }
        // row-pass inputs right of the box
        for (int x = b.cw; x < b.rin; x++) tile[gl * b.ts + x] = 0;
    }
    // column-pass inputs below the box (all w columns)
    for (int y = b.ch; y < b.cin; y++)
        for (int x = gl; x < b.w; x += G) tile[y * b.ts + x] = 0;
}

template <int MAXN = 64>
HD void itx2_phase_rows(const Itx2Blk &b, const int gl, int *tile) {
    if (gl >= b.ch) return;
    int *row = tile + gl * b.ts;
    const int rnd = b.wht ? 0 : (1 << b.shift) >> 1, sh = b.wht ? 0 : b.shift;
    Clamp c2 = b.colcl;
    if (b.wht) { c2.lo = (int)0x80000000; c2.hi = 0x7fffffff; }
    itx2_run<MAXN>(b.w, b.rk, b.cw, row, 1, row, 1, b.rowcl, rnd, sh, c2);
}

// column pass in place: column x of the tile (rows < cin in, rows < h out), (t + 8) >> 4 (WHT: t)
template <int MAXN = 64>
HD void itx2_phase_cols(const Itx2Blk &b, const int gl, const int G, int *tile) {
    const int rnd = b.wht ? 0 : 8, sh = b.wht ? 0 : 4;
    Clamp c2;
    c2.lo = (int)0x80000000; c2.hi = 0x7fffffff;
    for (int x = gl; x < b.w; x += G)
        itx2_run<MAXN>(b.h, b.ck, b.ch, tile + x, b.ts, tile + x, b.ts, b.colcl, rnd, sh, c2);
}

#if defined(__CUDACC__)
// One block by the group of G lanes that contains this lane (all lanes of the warp call this;
// lanes of a group without a block pass active = false).  res: see itx2_phase_out.
template <typename pixel, int MAXN = 64, typename CF = typename PxTraits<pixel>::coef *>
DEV void itx2_block(const bool active, const int gl, const int G, int *tile, const CF cf,
                    const int tx, const int txtp, const int eob, const int cw4, const int ch4, pixel *dst,
                    const int dstride, int16_t *res, const int rstride, const int bdmax, const bool zero_coefs)
{
    const Itx2Blk b = itx2_setup<pixel>(tx, txtp, eob, cw4, ch4, bdmax);
    int dc = 0;
    if (active && b.dc_only) dc = itx2_dc_value<pixel, CF>(b, cf);
    __syncwarp();                                   // every lane has read cf[0] before it is cleared
    if (active && b.dc_only && gl == 0 && zero_coefs) cf_zero(cf, 0);
    const bool full = active && !b.dc_only;
    if (full) itx2_phase_stage<pixel, CF>(b, gl, G, cf, tile, zero_coefs);
    __syncwarp();
    if (full) itx2_phase_rows<MAXN>(b, gl, tile);
    __syncwarp();
    if (full) itx2_phase_cols<MAXN>(b, gl, G, tile);
    __syncwarp();
    if (active) itx2_phase_out<pixel>(b, gl, G, tile, dc, dst, dstride, res, rstride, bdmax);
    __syncwarp();
}
#endif

}  // namespace d1
