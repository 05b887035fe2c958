// transform task kernel, sizes up to 16x16, 10/12-bit pixels (itx_task.cuh)
#include "itx_task.cuh"
namespace d1 {
int itx_task_small_16bpc(const ItxTaskArgs &a, int n, cudaStream_t st) {
    return itx_task_launch_one<uint16_t, false>(a, n, st);
}
}  // namespace d1
