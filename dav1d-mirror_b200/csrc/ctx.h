// Host-side plumbing shared by the translation units of libdav1d_cuda.so:
// sticky error, launch counter, the batched context and the per-call staging
// arena used by the DSP-table overrides.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <mutex>
#include <utility>
#include <stdlib.h>
#include "../../include/dav1d_cuda.h"

namespace d1 {

void set_error(int code, const char *what, const char *detail);
bool cuda_ok(cudaError_t e, const char *what);      // false + sticky error on failure
void count_launch(int n = 1);

#define D1_CHECK(call) do { if (!d1::cuda_ok((call), #call)) return -5; } while (0)
#define D1_CHECKV(call) do { if (!d1::cuda_ok((call), #call)) return; } while (0)

// Programmatic dependent launch (PDL): a kernel launched with this helper may start while
// its stream predecessor is still running (which hides the launch gap and the prologue's
// loads in chains of small launches, e.g. one per wavefront level).  Such a kernel issues
// griddepcontrol.launch_dependents first and griddepcontrol.wait before it touches anything
// the predecessor writes (d1::pdl_* in common.cuh); only launch-invariant inputs
// (descriptors, task lists) are read before the wait.  Measured on B200 inside captured graphs
// it did NOT pay (4K frame, 93 nodes: 836 -> 866 us alone, 272 -> 287 us/frame with 8 streams),
// so the attribute is off unless D1_PDL=1; the device-side instructions are no-ops then.
inline bool pdl_enabled() {
    static const bool on = getenv("D1_PDL") && atoi(getenv("D1_PDL")) != 0;
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                              cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Device views handed to kernels by value.
struct PlaneView {
    void *data;
    int64_t stride;   // bytes
    int w, h;
};
struct PicView {
    PlaneView p[3];
    int bdmax, ss_hor, ss_ver;
    const void *tma;  // Dav1dCudaPicture.tma: the planes' tensor maps in device memory, or null
};
inline PicView pic_view(const Dav1dCudaPicture *pic) {
    PicView v;
    for (int i = 0; i < 3; i++) {
        v.p[i].data = pic->p[i].data;
        v.p[i].stride = (int64_t)pic->p[i].stride;
        v.p[i].w = pic->p[i].w;
        v.p[i].h = pic->p[i].h;
    }
    v.bdmax = pic->bitdepth_max;
    v.ss_hor = pic->ss_hor;
    v.ss_ver = pic->ss_ver;
    v.tma = pic->tma;
    return v;
}

// Multi-frame transform tasks (itx.cu / recon.cu): per-frame planes, coefficients and the
// frame's level-sorted residual descriptors.
// storage of a frame's coefficient stream (Dav1dCudaReconBatch.cf_int16 / cf_esc)
struct CoefFmt {
    int s16;
    const Dav1dCudaCoefEsc *esc;
    int n_esc;
};

struct ItxFrameRef {
    PicView pic;
    PicView res;          // int16 residual planes (intra residual pre-pass)
    void *cf;
    const Dav1dCudaItxDesc *descs;
};

// Per-call staging: one arena per process, serialised by a mutex.  The
// reference calls DSP functions concurrently from many threads
// (SURVEY 8b "Threading"); correctness is kept by the lock, speed is not a
// goal of this surface.
struct Staging {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    uint8_t *dev = nullptr;        // device arena
    size_t dev_size = 0;
    uint8_t *host = nullptr;       // pinned mirror (same size)
    bool ok = false;
    bool ensure(size_t bytes);     // (re)allocate, returns false on failure
};
Staging &staging();

}  // namespace d1

struct Dav1dCudaContext {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int num_sms;
    void *tmp_pool;        // int16 scratch for unfused prep/compound
    size_t tmp_pool_bytes;
    // fork/join inside one frame: independent launch classes run on auxiliary
    // streams (and become parallel branches when the frame is captured as a graph)
    static constexpr int N_AUX = 15;
    cudaStream_t aux[N_AUX];
    cudaEvent_t ev_fork, ev_join[N_AUX];
    bool aux_ready;
    // intra executor: the status word the kernel raises (bit0: operations that wait for each
    // other) and the workspace of its rounds (pending / ready lists, counters); device memory
    unsigned *status;
    void *rounds_ws;
    size_t rounds_ws_bytes;
};
