// Host-side plumbing shared by the translation units of libdav1d_cuda.so:
// sticky error, launch counter, the batched context and the per-call staging
// arena used by the DSP-table overrides.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <mutex>
#include "../../include/dav1d_cuda.h"

namespace d1 {

void set_error(int code, const char *what, const char *detail);
bool cuda_ok(cudaError_t e, const char *what);      // false + sticky error on failure
void count_launch(int n = 1);

#define D1_CHECK(call) do { if (!d1::cuda_ok((call), #call)) return -5; } while (0)
#define D1_CHECKV(call) do { if (!d1::cuda_ok((call), #call)) return; } while (0)

// Device views handed to kernels by value.
struct PlaneView {
    void *data;
    int64_t stride;   // bytes
    int w, h;
};
struct PicView {
    PlaneView p[3];
    int bdmax, ss_hor, ss_ver;
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
    return v;
}

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
    static constexpr int N_AUX = 3;
    cudaStream_t aux[N_AUX];
    cudaEvent_t ev_fork, ev_join[N_AUX];
    bool aux_ready;
};
