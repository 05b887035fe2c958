// transform task kernel, sizes above 16x16, 8-bit pixels (itx_task.cuh)
#ifdef D1_ITX_BIG_NOINLINE
#define D1_ITX_PASS_NOINLINE
#endif
#include "itx_task.cuh"
namespace d1 {
int itx_task_big_8bpc(const ItxTaskArgs &a, int n, cudaStream_t st) {
    return itx_task_launch_one<uint8_t, true>(a, n, st);
}
}  // namespace d1
