// Measured FP32 peak of the device this library runs on: a dependent-chain-free FFMA loop (8 independent accumulators
// per thread, 4 CTAs of 256 threads per SM) timed by the caller with CUDA events.  MEASURED_PEAKS.json carries only HBM
// and bf16 tensor peaks; the dominant kernel of this path (cost_eval) is FP32-issue bound, so bench.py measures its
// roofline denominator here instead of quoting 148 x 128 x 2 x clock.  No counterpart in the reference.
#include "mpb_common.cuh"

namespace mpb {

__global__ void __launch_bounds__(256) ffma_peak_kernel(float* __restrict__ out, int iters, float a, float b) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = (float)(threadIdx.x + k);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(acc[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true: keeps the loop alive
}

}  // namespace mpb

// Launches the loop; flops executed = 2 * 64 * iters * threads, threads returned through *threads_out (host).
extern "C" int mpb_bench_fp32_peak(float* scratch, int iters, long long* threads_out, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(scratch && iters >= 1 && threads_out, "mpb_bench_fp32_peak: bad arguments");
    const int grid = sm_count() * 8;
    *threads_out = (long long)grid * 256;
    ffma_peak_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(scratch, iters, 0.999f, 0.001f);
    return check_launch("mpb_bench_fp32_peak");
}
