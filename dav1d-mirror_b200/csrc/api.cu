// Context, error channel, HBM picture allocation and the per-call staging
// arena of libdav1d_cuda.so.
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <cuda.h>            // CUtensorMap and its enums (types only: no libcuda symbol is linked)
#include "ctx.h"

namespace d1 {

void mc_init_attrs();
void mc_set_tma(int mode);
int mc_get_tma();
void recon_init_attrs();

static std::atomic<int> g_err{0};
static char g_err_msg[512] = "";
static std::mutex g_err_mu;
static std::atomic<uint64_t> g_launches{0};

void set_error(int code, const char *what, const char *detail) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    if (g_err.load() == 0) {   // sticky: first error wins
        g_err.store(code);
        snprintf(g_err_msg, sizeof(g_err_msg), "%s: %s", what, detail ? detail : "");
        fprintf(stderr, "[dav1d_cuda] error %d: %s\n", code, g_err_msg);
    }
}

bool cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    set_error((int)e, what, cudaGetErrorString(e));
    return false;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n); }

Staging &staging() {
    static Staging s;
    return s;
}

bool Staging::ensure(size_t bytes) {
    if (!ok) {
        int n = 0;
        if (!cuda_ok(cudaGetDeviceCount(&n), "cudaGetDeviceCount") || n < 1) {
            set_error(-19, "no CUDA device", "the DSP-table overrides have no CPU fallback");
            return false;
        }
        if (!cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate"))
            return false;
        ok = true;
    }
    if (bytes <= dev_size) return true;
    size_t want = dev_size ? dev_size : (size_t)4 << 20;
    while (want < bytes) want <<= 1;
    if (dev) cudaFree(dev);
    if (host) cudaFreeHost(host);
    dev = nullptr; host = nullptr; dev_size = 0;
    if (!cuda_ok(cudaMalloc(&dev, want), "cudaMalloc(staging)")) return false;
    if (!cuda_ok(cudaMallocHost(&host, want), "cudaMallocHost(staging)")) return false;
    dev_size = want;
    return true;
}

}  // namespace d1

using namespace d1;

extern "C" {

int dav1d_cuda_last_error(void) { return g_err.load(); }
const char *dav1d_cuda_last_error_string(void) { return g_err_msg; }
void dav1d_cuda_clear_error(void) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err.store(0);
    g_err_msg[0] = 0;
    cudaGetLastError();
}
uint64_t dav1d_cuda_launch_count(void) { return g_launches.load(); }

int dav1d_cuda_available(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        cudaGetLastError();
        return 0;
    }
    return 1;
}

int dav1d_cuda_open(Dav1dCudaContext **out, int device, void *stream) {
    if (!out) return -22;
    *out = nullptr;
    int n = 0;
    D1_CHECK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        set_error(-22, "dav1d_cuda_open", "no such device");
        return -22;
    }
    // One GPU per process: the per-call staging arena, the occupancy figures of the persistent
    // kernels and the kernel attributes are process-wide and belong to the first device used.
    static std::atomic<int> g_device{-1};
    int expect = -1;
    if (!g_device.compare_exchange_strong(expect, device) && expect != device) {
        set_error(-22, "dav1d_cuda_open", "this process already uses another device (one GPU per process)");
        return -22;
    }
    D1_CHECK(cudaSetDevice(device));
    Dav1dCudaContext *c = new Dav1dCudaContext();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->stream = (cudaStream_t)stream;
    bool ok = true;
    if (!stream) {   // no caller stream: give the context its own so that contexts run concurrently
        ok = cuda_ok(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate");
        c->own_stream = ok;
    }
    int sms = 0;
    ok = ok && cuda_ok(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device), "cudaDeviceGetAttribute");
    c->num_sms = sms;
    ok = ok && cuda_ok(cudaMalloc(&c->status, 64), "cudaMalloc(status)");
    ok = ok && cuda_ok(cudaMemset(c->status, 0, 64), "cudaMemset(status)");
    if (!ok) {
        if (c->status) cudaFree(c->status);
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        return -5;
    }
    mc_init_attrs();
    recon_init_attrs();
    cudaGetLastError();
    *out = c;
    return 0;
}

void dav1d_cuda_close(Dav1dCudaContext *c) {
    if (!c) return;
    cudaStreamSynchronize(c->stream);
    if (c->tmp_pool) cudaFree(c->tmp_pool);
    if (c->aux_ready) {
        for (int i = 0; i < Dav1dCudaContext::N_AUX; i++) {
            cudaStreamDestroy(c->aux[i]);
            cudaEventDestroy(c->ev_join[i]);
        }
        cudaEventDestroy(c->ev_fork);
    }
    if (c->own_stream) cudaStreamDestroy(c->stream);
    if (c->status) cudaFree(c->status);
    if (c->rounds_ws) cudaFree(c->rounds_ws);
    delete c;
}

int dav1d_cuda_synchronize(Dav1dCudaContext *c) {
    if (!c) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    D1_CHECK(cudaStreamSynchronize(c->stream));
    D1_CHECK(cudaGetLastError());
    // status word of the intra executor: a dependency that never became final
    unsigned st = 0;
    D1_CHECK(cudaMemcpy(&st, c->status, sizeof(st), cudaMemcpyDeviceToHost));
    if (st) {
        cudaMemset(c->status, 0, sizeof(st));
        set_error(-5, "intra executor", "intra-class operations wait for each other: the frame is incomplete "
                                        "(inconsistent descriptors)");
        return -5;
    }
    return 0;
}

// Tensor maps of a picture's planes for the TMA staging of the MC kernels (csrc/mc.cuh): per plane
// MC_TMA_CLASSES maps whose boxes are one staged window row wide (48 16-bit / 64 8-bit pixels) and
// 15 / 23 / 39 rows high.  cuTensorMapEncodeTiled comes through the runtime's entry-point query: the
// library does not link libcuda.
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}
constexpr int TMA_CLASSES = 3;
const int tma_rows[TMA_CLASSES] = { 15, 23, 39 };
// false: the picture goes without maps (cp.async staging) - not an error
bool picture_tensor_maps(Dav1dCudaPicture *pic, void *dev_maps) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const int hbd = pic->bitdepth_max > 0xff;
    CUtensorMap maps[3 * TMA_CLASSES];
    memset(maps, 0, sizeof(maps));
    for (int pl = 0; pl < 3; pl++) {
        const Dav1dCudaPlane &p = pic->p[pl];
        if (!p.data || p.w <= 0 || p.h <= 0) continue;
        if (p.stride <= 0 || (p.stride & 15) || ((uintptr_t)p.data & 15)) return false;
        for (int k = 0; k < TMA_CLASSES; k++) {
            const cuuint64_t dims[2] = { (cuuint64_t)(p.stride >> hbd), (cuuint64_t)p.h };
            const cuuint64_t strides[1] = { (cuuint64_t)p.stride };
            const cuuint32_t box[2] = { hbd ? 48u : 64u, (cuuint32_t)tma_rows[k] };
            const cuuint32_t estr[2] = { 1, 1 };
            if (enc(&maps[pl * TMA_CLASSES + k], hbd ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                    p.data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return false;
        }
    }
    return cudaMemcpy(dev_maps, maps, sizeof(maps), cudaMemcpyHostToDevice) == cudaSuccess;
}
}  // namespace

// Geometry of the reference's default allocator, src/picture.c:46-84.
int dav1d_cuda_picture_alloc(Dav1dCudaContext *c, Dav1dCudaPicture *pic,
                             int w, int h, int ss_hor, int ss_ver, int bitdepth_max)
{
    if (!c || !pic || w <= 0 || h <= 0) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    memset(pic, 0, sizeof(*pic));
    const int hbd = bitdepth_max > 0xff;
    const int aligned_w = (w + 127) & ~127;
    const int aligned_h = (h + 127) & ~127;
    ptrdiff_t y_stride = (ptrdiff_t)aligned_w << hbd;
    ptrdiff_t uv_stride = y_stride >> ss_hor;
    if (!(y_stride & 1023)) y_stride += 64;
    if (!(uv_stride & 1023)) uv_stride += 64;
    const size_t y_sz = (size_t)y_stride * aligned_h;
    const size_t uv_sz = (size_t)uv_stride * (aligned_h >> ss_ver);
    uint8_t *buf = nullptr;
    // the planes, 64 bytes of slack as the reference has them, then the planes' tensor maps
    const size_t maps_off = (y_sz + 2 * uv_sz + 64 + 127) & ~(size_t)127;
    D1_CHECK(cudaMalloc(&buf, maps_off + 3 * TMA_CLASSES * sizeof(CUtensorMap)));
    D1_CHECK(cudaMemset(buf, 0, maps_off + 3 * TMA_CLASSES * sizeof(CUtensorMap)));
    pic->p[0].data = buf;
    pic->p[0].stride = y_stride;
    pic->p[0].w = w;
    pic->p[0].h = h;
    for (int i = 1; i < 3; i++) {
        pic->p[i].data = buf + y_sz + (i - 1) * uv_sz;
        pic->p[i].stride = uv_stride;
        pic->p[i].w = (w + ss_hor) >> ss_hor;
        pic->p[i].h = (h + ss_ver) >> ss_ver;
    }
    pic->bitdepth_max = bitdepth_max;
    pic->ss_hor = ss_hor;
    pic->ss_ver = ss_ver;
    if (picture_tensor_maps(pic, buf + maps_off)) pic->tma = buf + maps_off;
    return 0;
}

void dav1d_cuda_set_mc_tma(int mode) { mc_set_tma(mode); }
int dav1d_cuda_get_mc_tma(void) { return mc_get_tma(); }

// ---- Dav1dPicAllocator seam: pictures with a twin in HBM and one in pinned host memory
namespace {
struct TwinPicture {
    uint32_t magic;
    Dav1dCudaContext *ctx;
    Dav1dCudaPicture dev;
    uint8_t *host;           // pinned; data[] of the Dav1dPicture point into it
    size_t bytes;            // of either copy (planes + padding)
};
constexpr uint32_t TWIN_MAGIC = 0x44315450u;     // "D1TP"

int twin_alloc(Dav1dCudaDav1dPicture *p, void *cookie) {
    Dav1dCudaContext *c = (Dav1dCudaContext *)cookie;
    if (!p || !c || p->p.w <= 0 || p->p.h <= 0 || (p->p.bpc != 8 && p->p.bpc != 10 && p->p.bpc != 12)) return -22;
    const int layout = p->p.layout;                    // I400 0, I420 1, I422 2, I444 3
    const int has_chroma = layout != 0;
    const int ss_ver = layout == 1, ss_hor = layout != 3;
    TwinPicture *t = new TwinPicture();
    t->magic = TWIN_MAGIC; t->ctx = c;
    // the device copy: same rules as src/picture.c:46-84 (dav1d_cuda_picture_alloc)
    const int r = dav1d_cuda_picture_alloc(c, &t->dev, p->p.w, p->p.h, ss_hor, ss_ver, (1 << p->p.bpc) - 1);
    if (r) { delete t; return -12; }
    const int aligned_h = (p->p.h + 127) & ~127;
    const ptrdiff_t y_stride = t->dev.p[0].stride, uv_stride = has_chroma ? t->dev.p[1].stride : 0;
    const size_t y_sz = (size_t)y_stride * aligned_h, uv_sz = (size_t)t->dev.p[1].stride * (aligned_h >> ss_ver);
    t->bytes = y_sz + 2 * uv_sz;
    if (!cuda_ok(cudaMallocHost((void **)&t->host, t->bytes + 64), "cudaMallocHost(picture)")) {
        dav1d_cuda_picture_free(c, &t->dev);
        delete t;
        return -12;
    }
    memset(t->host, 0, t->bytes + 64);
    p->stride[0] = y_stride;
    p->stride[1] = uv_stride;
    p->data[0] = t->host;                               // cudaMallocHost memory is page-aligned (>= DAV1D_PICTURE_ALIGNMENT)
    p->data[1] = has_chroma ? t->host + y_sz : nullptr;
    p->data[2] = has_chroma ? t->host + y_sz + uv_sz : nullptr;
    p->allocator_data = t;
    return 0;
}
void twin_release(Dav1dCudaDav1dPicture *p, void *) {
    TwinPicture *t = p ? (TwinPicture *)p->allocator_data : nullptr;
    if (!t || t->magic != TWIN_MAGIC) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);              // a copy or a batch may still use the picture
    dav1d_cuda_picture_free(t->ctx, &t->dev);
    cudaFreeHost(t->host);
    t->magic = 0;
    delete t;
    p->allocator_data = nullptr;
}
TwinPicture *twin_of(const Dav1dCudaDav1dPicture *p) {
    TwinPicture *t = p ? (TwinPicture *)p->allocator_data : nullptr;
    return t && t->magic == TWIN_MAGIC ? t : nullptr;
}
}  // namespace

int dav1d_cuda_pic_allocator_init(Dav1dCudaContext *c, Dav1dCudaPicAllocator *a) {
    if (!c || !a) return -22;
    a->cookie = c;
    a->alloc_picture_callback = twin_alloc;
    a->release_picture_callback = twin_release;
    return 0;
}
const Dav1dCudaPicture *dav1d_cuda_picture_of(const Dav1dCudaDav1dPicture *pic) {
    TwinPicture *t = twin_of(pic);
    return t ? &t->dev : nullptr;
}
int dav1d_cuda_picture_to_host(Dav1dCudaContext *c, const Dav1dCudaDav1dPicture *pic) {
    TwinPicture *t = twin_of(pic);
    if (!c || !t) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    D1_CHECK(cudaMemcpyAsync(t->host, t->dev.p[0].data, t->bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}
int dav1d_cuda_picture_to_device(Dav1dCudaContext *c, const Dav1dCudaDav1dPicture *pic) {
    TwinPicture *t = twin_of(pic);
    if (!c || !t) return -22;
    D1_CHECK(cudaSetDevice(c->device));
    D1_CHECK(cudaMemcpyAsync(t->dev.p[0].data, t->host, t->bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

void dav1d_cuda_picture_free(Dav1dCudaContext *c, Dav1dCudaPicture *pic) {
    (void)c;
    if (pic && pic->p[0].data) cudaFree(pic->p[0].data);
    if (pic) memset(pic, 0, sizeof(*pic));
}

int dav1d_cuda_picture_upload(Dav1dCudaContext *c, const Dav1dCudaPicture *pic, int plane,
                              const void *host, ptrdiff_t host_stride)
{
    const int hbd = pic->bitdepth_max > 0xff;
    const Dav1dCudaPlane *p = &pic->p[plane];
    // equal strides: one linear copy (measured on B200: pitched 2-D copies do not overlap with
    // copies in the opposite direction, linear ones do - tools/exp_copy.py)
    if (host_stride == p->stride)
        D1_CHECK(cudaMemcpyAsync(p->data, host, (size_t)(p->h - 1) * p->stride + ((size_t)p->w << hbd),
                                 cudaMemcpyHostToDevice, c->stream));
    else
        D1_CHECK(cudaMemcpy2DAsync(p->data, p->stride, host, host_stride, (size_t)p->w << hbd, p->h,
                                   cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int dav1d_cuda_picture_download(Dav1dCudaContext *c, const Dav1dCudaPicture *pic, int plane,
                                void *host, ptrdiff_t host_stride)
{
    const int hbd = pic->bitdepth_max > 0xff;
    const Dav1dCudaPlane *p = &pic->p[plane];
    if (host_stride == p->stride)
        D1_CHECK(cudaMemcpyAsync(host, p->data, (size_t)(p->h - 1) * p->stride + ((size_t)p->w << hbd),
                                 cudaMemcpyDeviceToHost, c->stream));
    else
        D1_CHECK(cudaMemcpy2DAsync(host, host_stride, p->data, p->stride, (size_t)p->w << hbd, p->h,
                                   cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

}  // extern "C"
