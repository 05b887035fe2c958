// 8-bit instantiation of the transform task kernels (itx_task.cuh).
#include "itx_task.cuh"
namespace d1 {
int itx_task_launch_8bpc(const ItxTaskArgs &a, int n_small, int n_big, cudaStream_t st_small, cudaStream_t st_big) {
    return itx_task_launch_px<uint8_t>(a, n_small, n_big, st_small, st_big);
}
}  // namespace d1
