// Shared helpers for the sm_100a kernels of libmpb_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>

#include "../../include/mpb.h"

#define MPB_FULL_MASK 0xffffffffu

namespace mpb {

void set_error(const char* fmt, ...);
int check_launch(const char* what);     // cudaGetLastError() -> MPB_OK / MPB_ECUDA
int sm_count();                         // SMs of the current device (cached per device)
// Two zero-initialised device words for a self-resetting dynamic work counter (next index, finished CTAs).  Slots come
// from a small per-device pool owned by the library (the only memory it ever allocates: 64 x 8 bytes per device,
// created by mpb_init() or lazily by the first launch) and are handed out round-robin, so up to 64 launches may be in
// flight concurrently on different streams of one device.  The last CTA of a launch re-arms its slot.
unsigned* sched_slot();
int init_current_device();              // allocate + zero the pool of the current device (mpb_init)

// Programmatic dependent launch (the three kernels of a Stoch-GPMP iteration run back to back on one stream): a kernel
// launched with launch_pdl may start while its predecessor drains -- its CTAs become resident as the predecessor's exit --
// and runs everything that does not depend on the predecessor (staging of the robot / field tables, barrier and
// tensor-memory set-up) before pdl_wait(), which returns once the predecessor grid has completed and its writes are visible.
// pdl_trigger() at the top of a kernel lets ITS successor be scheduled as early as resources allow.  Without the launch
// attribute (or after a kernel that never triggers) both instructions are no-ops / ordinary stream order.  MPB_PDL=0 disables it.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

#define MPB_REQUIRE(cond, ...)                       \
    do {                                             \
        if (!(cond)) {                               \
            mpb::set_error(__VA_ARGS__);             \
            return MPB_EINVAL;                       \
        }                                            \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MPB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MPB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(MPB_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ int warp_and(int v) { return __all_sync(MPB_FULL_MASK, v); }

}  // namespace mpb
