// Inverse transforms, all sizes in one launch ("task" kernels).
//
// The recorder groups residual descriptors by transform size (and, inside a size, by
// transform type); a task is up to 32/G consecutive descriptors of one size handled by one
// warp, G lanes per block, encoded as first_index << 8 | tx << 3 | (count - 1).  Two
// instantiations per pixel type: sizes up to 16x16 (few registers, 2 KB of shared memory per
// warp) and the larger ones.  The same kernel serves one frame (planes / coefficients /
// descriptors in the launch parameters) and the merged dependency level of several frames
// (task = (code, frame), per-frame parameters from a device table).
//
// Included by itx_task_{s,b}{8,16}.cu - one translation unit per pixel type and size group, so
// that the pieces of this large code build in parallel and can choose between inlined (small
// sizes: fewer registers) and shared out-of-line row / column passes (D1_ITX_PASS_NOINLINE).
#pragma once
#include "ctx.h"
#include "itx.cuh"

namespace d1 {

constexpr int ITX_TASK_WARPS = 4;
constexpr int ITX_TASK_SMEM_SMALL = 2 * 16 * 17 * 4;     // 16x16: two blocks per warp
constexpr int ITX_TASK_SMEM_BIG = 32 * 65 * 4;           // 64-wide: one block per warp

struct ItxTaskArgs {
    PicView pic;
    void *cf;
    const Dav1dCudaItxDesc *descs;
    const uint32_t *tasks;       // single frame: task codes
    const ItxFrameRef *frames;   // several frames: per-frame parameters (else nullptr) ...
    const uint2 *mtasks;         // ... and (code, frame) tasks
    int n_tasks;
    int zero_coefs;
};

template <typename pixel, int W, int H>
DEV void itx_task_body(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs, const bool zero_coefs,
                       const int first, const int cnt, int *smem, const int lane) {
    typedef ItxGeom<W, H> Geo;
    typedef typename PxTraits<pixel>::coef coef;
    constexpr int G = Geo::GMIN;
    const int grp = lane / G, gl = lane % G;
    const bool active = grp < cnt;
    Dav1dCudaItxDesc d;
    if (active) d = descs[first + grp];
    else { d.coef_off = 0; d.x = d.y = 0; d.eob = 0; d.plane = 0; d.tx = 0; d.txtp = 0; d.cw4 = d.ch4 = 0; }
    int *tile = smem + grp * Geo::TILE_INTS;
    const PlaneView &pv = pic.p[d.plane];
    const int dstride = (int)(pv.stride / (int)sizeof(pixel));
    pixel *dst = (pixel *)pv.data + (int64_t)d.y * dstride + d.x;
    itx_block<pixel, W, H, G>(active, gl, tile, (coef *)cf + d.coef_off, d.eob, d.txtp, dst, dstride,
                              pic.bdmax, zero_coefs, d.cw4, d.ch4);
}

template <typename pixel, bool BIG>
__global__ void __launch_bounds__(ITX_TASK_WARPS * 32, BIG ? 4 : 8) itx_task_kernel(const __grid_constant__ ItxTaskArgs a) {
    extern __shared__ int itx_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * ITX_TASK_WARPS + warp;
    pdl_launch_dependents();
    if (t >= a.n_tasks) return;
    int *smem = itx_smem + warp * ((BIG ? ITX_TASK_SMEM_BIG : ITX_TASK_SMEM_SMALL) / 4);
    uint32_t code;
    const PicView *pic = &a.pic;
    void *cf = a.cf;
    const Dav1dCudaItxDesc *descs = a.descs;
    if (a.frames) {
        const uint2 tk = a.mtasks[t];
        const ItxFrameRef *fr = a.frames + tk.y;
        code = tk.x; pic = &fr->pic; cf = fr->cf;      // descs: the frames' descriptors concatenated (a.descs)
    } else {
        code = a.tasks[t];
    }
    pdl_wait();
    const bool zero = a.zero_coefs != 0;
    const int first = (int)(code >> 8), tx = (code >> 3) & 31, cnt = (int)(code & 7) + 1;
#define D1_TASKCASE(T, W, H) \
    case T: \
        if constexpr (BIG == (W > 16 || H > 16)) \
            itx_task_body<pixel, W, H>(*pic, cf, descs, zero, first, cnt, smem, lane); \
        break;
    switch (tx) {
    D1_TASKCASE(0, 4, 4) D1_TASKCASE(1, 8, 8) D1_TASKCASE(2, 16, 16) D1_TASKCASE(3, 32, 32) D1_TASKCASE(4, 64, 64)
    D1_TASKCASE(5, 4, 8) D1_TASKCASE(6, 8, 4) D1_TASKCASE(7, 8, 16) D1_TASKCASE(8, 16, 8) D1_TASKCASE(9, 16, 32)
    D1_TASKCASE(10, 32, 16) D1_TASKCASE(11, 32, 64) D1_TASKCASE(12, 64, 32) D1_TASKCASE(13, 4, 16)
    D1_TASKCASE(14, 16, 4) D1_TASKCASE(15, 8, 32) D1_TASKCASE(16, 32, 8) D1_TASKCASE(17, 16, 64)
    D1_TASKCASE(18, 64, 16)
    default: break;
    }
#undef D1_TASKCASE
}

// one launch: n tasks of the small (<= 16x16) or the large sizes
template <typename pixel, bool BIG>
int itx_task_launch_one(ItxTaskArgs a, int n, cudaStream_t st) {
    static bool attr = false;
    if (!attr && BIG) {
        cudaFuncSetAttribute(itx_task_kernel<pixel, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             ITX_TASK_WARPS * ITX_TASK_SMEM_BIG);
        attr = true;
    }
    if (n <= 0) return 0;
    a.n_tasks = n;
    const int grid = (n + ITX_TASK_WARPS - 1) / ITX_TASK_WARPS;
    launch_pdl(itx_task_kernel<pixel, BIG>, grid, ITX_TASK_WARPS * 32,
               ITX_TASK_WARPS * (BIG ? ITX_TASK_SMEM_BIG : ITX_TASK_SMEM_SMALL), st, a);
    count_launch();
    return cuda_ok(cudaGetLastError(), "itx_task_kernel") ? 0 : -5;
}

// one translation unit per (pixel type, size group): itx_task_{s,b}{8,16}.cu
int itx_task_small_8bpc(const ItxTaskArgs &a, int n, cudaStream_t st);
int itx_task_big_8bpc(const ItxTaskArgs &a, int n, cudaStream_t st);
int itx_task_small_16bpc(const ItxTaskArgs &a, int n, cudaStream_t st);
int itx_task_big_16bpc(const ItxTaskArgs &a, int n, cudaStream_t st);

}  // namespace d1
