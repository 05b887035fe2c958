// Shared helpers for the sm_100a kernels of the dav1d pixel-reconstruction path.
//
// Everything in this tree is integer arithmetic that has to match the
// reference's C templates bit for bit (include/common/intops.h,
// include/common/bitdepth.h in the reference).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#else
#define HD inline
#define DEV inline
#endif

namespace d1 {

HD int imin(int a, int b) { return a < b ? a : b; }
HD int imax(int a, int b) { return a > b ? a : b; }
HD int iclip(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
HD int iabs(int v) { return v < 0 ? -v : v; }

// pixel traits: u8 planes for 8 bpc, u16 planes for 10/12 bpc
// (reference: include/common/bitdepth.h:54-83)
template <typename pixel> struct PxTraits;
template <> struct PxTraits<uint8_t> {
    typedef int16_t coef;
    static constexpr bool hbd = false;
    static constexpr int prep_bias = 0;                 // mc_tmpl.c:39-42
    HD static int bitdepth(int) { return 8; }
    HD static int inter_bits(int) { return 4; }
};
template <> struct PxTraits<uint16_t> {
    typedef int32_t coef;
    static constexpr bool hbd = true;
    static constexpr int prep_bias = 8192;              // mc_tmpl.c:44-48
    HD static int bitdepth(int bdmax) { return bdmax > 1023 ? 12 : 10; }
    HD static int inter_bits(int bdmax) { return bdmax > 1023 ? 2 : 4; }
};

// programmatic dependent launch, see d1::launch_pdl (ctx.h); no-ops in a plain launch
#ifdef __CUDACC__
DEV void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
DEV void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

template <typename pixel> HD int clip_px(int v, int bdmax) {
#ifdef __CUDA_ARCH__
    return min(max(v, 0), bdmax);
#else
    return v < 0 ? 0 : v > bdmax ? bdmax : v;
#endif
}

}  // namespace d1
