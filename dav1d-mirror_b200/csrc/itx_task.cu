// Host side of the transform task kernels (itx_task.cuh): launch entry points shared by both
// pixel types and the task builder.
#include <string.h>
#include "itx_task.cuh"
namespace d1 {

static int itx_task_launch_both(ItxTaskArgs a, int n_small, int n_big, bool hbd, cudaStream_t st_small,
                                cudaStream_t st_big)
{
    int r = hbd ? itx_task_small_16bpc(a, n_small, st_small) : itx_task_small_8bpc(a, n_small, st_small);
    if (r) return r;
    if (a.tasks) a.tasks += n_small;
    if (a.mtasks) a.mtasks += n_small;
    return hbd ? itx_task_big_16bpc(a, n_big, st_big) : itx_task_big_8bpc(a, n_big, st_big);
}

// tasks[0 .. n_small) = sizes up to 16x16, tasks[n_small .. n_small + n_big) = larger
int itx_task_launch(const PicView &pic, void *cf, const Dav1dCudaItxDesc *descs, const uint32_t *tasks,
                    int n_small, int n_big, int zero_coefs, cudaStream_t st_small, cudaStream_t st_big)
{
    ItxTaskArgs a;
    a.pic = pic; a.cf = cf; a.descs = descs; a.tasks = tasks; a.frames = nullptr; a.mtasks = nullptr;
    a.n_tasks = 0; a.zero_coefs = zero_coefs;
    return itx_task_launch_both(a, n_small, n_big, pic.bdmax > 0xff, st_small, st_big);
}

// one dependency level of several frames: tasks = (code, frame), code indexes `descs` = the
// residual descriptors of all frames concatenated
int itx_multi_task_launch(const ItxFrameRef *frames, const Dav1dCudaItxDesc *descs, const uint2 *tasks, int n_small,
                          int n_big, bool hbd, cudaStream_t st_small, cudaStream_t st_big)
{
    ItxTaskArgs a;
    memset(&a, 0, sizeof(a));
    a.frames = frames; a.mtasks = tasks; a.descs = descs;
    return itx_task_launch_both(a, n_small, n_big, hbd, st_small, st_big);
}

// blocks of one size a warp takes: 32 / G
static int itx_bpw(int tx) {
    const TxDim t = tx_dim(tx);
    const int sw = t.w < 32 ? t.w : 32, sh = t.h < 32 ? t.h : 32;
    return 32 / (sh > sw ? sh : sw);
}

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
            const TxDim t = tx_dim(tx);
            const bool big = t.w > 16 || t.h > 16;
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

void itx_init_attrs() {}

}  // namespace d1
