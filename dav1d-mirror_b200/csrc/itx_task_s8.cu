// transform task kernel, sizes up to 16x16, 8-bit pixels (itx_task.cuh)
#include "itx_task.cuh"
namespace d1 {
int itx_task_small_8bpc(const ItxTaskArgs &a, int n, cudaStream_t st) {
    return itx_task_launch_one<uint8_t, false>(a, n, st);
}
}  // namespace d1
