// Pieces shared by the transform code paths (itx.cuh, itx2.cuh): pixel vector
// loads / stores, transform geometry (enum RectTxfmSize -> w, h, shift) and the
// TxfmType -> (row, column) 1-D kind mapping.  Host-compilable.
#pragma once
#include "itx_1d.cuh"

#if !defined(__CUDACC__)
#define __ldcg(p) (*(p))
#endif

namespace d1 {

#if defined(__CUDACC__)
// 64/128-bit pixel vectors (L2-coherent loads); scalar fallback when misaligned.
template <typename pixel, int VW> DEV void load_px(const pixel *p, int *v) {
    constexpr int BYTES = VW * (int)sizeof(pixel);
    if (((uintptr_t)p & (BYTES - 1)) == 0) {
        uint32_t q[BYTES / 4];
        if (BYTES == 16) { const uint4 t = __ldcg((const uint4 *)p); q[0] = t.x; q[1] = t.y; q[BYTES / 4 - 2] = t.z; q[BYTES / 4 - 1] = t.w; }
        else if (BYTES == 8) { const uint2 t = __ldcg((const uint2 *)p); q[0] = t.x; q[BYTES / 4 - 1] = t.y; }
        else q[0] = __ldcg((const uint32_t *)p);
#pragma unroll
        for (int k = 0; k < VW; k++)
            v[k] = sizeof(pixel) == 2 ? (int)((q[k / 2] >> (16 * (k & 1))) & 0xffff) : (int)((q[k / 4] >> (8 * (k & 3))) & 0xff);
    } else {
#pragma unroll
        for (int k = 0; k < VW; k++) v[k] = __ldcg(p + k);
    }
}
template <typename pixel, int VW> DEV void store_px(pixel *p, const int *v) {
    constexpr int BYTES = VW * (int)sizeof(pixel);
    if (((uintptr_t)p & (BYTES - 1)) == 0) {
        if (sizeof(pixel) == 2) {
            uint32_t q[VW / 2];
#pragma unroll
            for (int k = 0; k < VW / 2; k++) q[k] = (uint32_t)(v[2 * k] & 0xffff) | ((uint32_t)v[2 * k + 1] << 16);
            if (VW == 8) *(uint4 *)p = make_uint4(q[0], q[1], q[VW / 2 - 2], q[VW / 2 - 1]);
            else *(uint2 *)p = make_uint2(q[0], q[1]);
        } else {
            uint32_t q[VW / 4];
#pragma unroll
            for (int k = 0; k < VW / 4; k++)
                q[k] = (uint32_t)(v[4 * k] & 0xff) | ((uint32_t)(v[4 * k + 1] & 0xff) << 8) |
                       ((uint32_t)(v[4 * k + 2] & 0xff) << 16) | ((uint32_t)v[4 * k + 3] << 24);
            if (VW == 8) *(uint2 *)p = make_uint2(q[0], q[VW / 4 - 1]);
            else *(uint32_t *)p = q[0];
        }
    } else {
#pragma unroll
        for (int k = 0; k < VW; k++) p[k] = (pixel)v[k];
    }
}
#endif

struct TxDim { uint8_t w, h, shift; };

// enum RectTxfmSize -> (w, h, inter-pass shift), itx_tmpl.c:142-160 + levels.h:44-78.
// Packed constants (log2(w/4), log2(h/4): 3 bits per size; shift: 2 bits) instead of a table in
// local memory.
HD TxDim tx_dim(int tx) {
    const unsigned long long LW = 0x1132846d2244688ull, LH = 0xa161389940c688ull, SH = 0x2a955542a4ull;
    TxDim t;
    t.w = (uint8_t)(4 << ((LW >> (3 * tx)) & 7));
    t.h = (uint8_t)(4 << ((LH >> (3 * tx)) & 7));
    t.shift = (uint8_t)((SH >> (2 * tx)) & 3);
    return t;
}

// enum TxfmType is named VERT_HORZ (levels.h:80-100); the row (first) pass is
// the horizontal transform (itx_tmpl.c:212-245).
HD int txtp_row_kind(int txtp) {
    // D=0 A=1 F=2 I=3 W=4, 3 bits each
    const uint64_t rows = 0ull | (0ull << 0) | (0ull << 3) | (1ull << 6) | (1ull << 9) | (0ull << 12) |
                          (2ull << 15) | (2ull << 18) | (2ull << 21) | (1ull << 24) | (3ull << 27) |
                          (3ull << 30) | (0ull << 33) | (3ull << 36) | (1ull << 39) | (3ull << 42) |
                          (2ull << 45) | (4ull << 48);
    return (int)((rows >> (3 * txtp)) & 7);
}
HD int txtp_col_kind(int txtp) {
    const uint64_t cols = 0ull | (0ull << 0) | (1ull << 3) | (0ull << 6) | (1ull << 9) | (2ull << 12) |
                          (0ull << 15) | (2ull << 18) | (1ull << 21) | (2ull << 24) | (3ull << 27) |
                          (0ull << 30) | (3ull << 33) | (1ull << 36) | (3ull << 39) | (2ull << 42) |
                          (3ull << 45) | (4ull << 48);
    return (int)((cols >> (3 * txtp)) & 7);
}

}  // namespace d1
