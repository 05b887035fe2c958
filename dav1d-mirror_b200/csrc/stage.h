// Helper for the per-call (DSP-table) surface: lays host operands out in the
// pinned staging mirror, ships them to the device arena in one copy, and
// brings selected regions back after the kernel.
#pragma once
#include <string.h>
#include <vector>
#include "ctx.h"

namespace d1 {

class Stage {
public:
    explicit Stage(Staging &s) : s_(s), off_(0) {}

    // Reserve `bytes` (256-byte aligned). Returns the arena offset.
    size_t reserve(size_t bytes) {
        const size_t o = off_;
        off_ = (off_ + bytes + 255) & ~(size_t)255;
        return o;
    }
    // Must be called once after all reserve() calls and before put()/host().
    bool commit() { return s_.ensure(off_ ? off_ : 256); }

    uint8_t *host(size_t off) { return s_.host + off; }
    uint8_t *dev(size_t off) { return s_.dev + off; }

    // Copy `rows` rows of `row_bytes` from a strided host buffer into the arena (dense rows of `dst_stride`).
    void put2d(size_t off, size_t dst_stride, const void *src, ptrdiff_t src_stride, size_t row_bytes, int rows) {
        for (int y = 0; y < rows; y++)
            memcpy(s_.host + off + (size_t)y * dst_stride, (const uint8_t *)src + (ptrdiff_t)y * src_stride, row_bytes);
    }
    void get2d(size_t off, size_t src_stride_, void *dst, ptrdiff_t dst_stride, size_t row_bytes, int rows) {
        for (int y = 0; y < rows; y++)
            memcpy((uint8_t *)dst + (ptrdiff_t)y * dst_stride, s_.host + off + (size_t)y * src_stride_, row_bytes);
    }
    bool upload() { return cuda_ok(cudaMemcpyAsync(s_.dev, s_.host, off_, cudaMemcpyHostToDevice, s_.stream), "H2D staging"); }
    bool download(size_t off, size_t bytes) {
        return cuda_ok(cudaMemcpyAsync(s_.host + off, s_.dev + off, bytes, cudaMemcpyDeviceToHost, s_.stream), "D2H staging");
    }
    bool sync() {
        return cuda_ok(cudaStreamSynchronize(s_.stream), "staging sync") &&
               cuda_ok(cudaGetLastError(), "kernel");
    }
    cudaStream_t stream() { return s_.stream; }

private:
    Staging &s_;
    size_t off_;
};

}  // namespace d1
