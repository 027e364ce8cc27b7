// Stand-alone micro-benchmark (not part of libmpb_b200): FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+) issue rate, and
// FFMA2 with interleaved ALU-pipe work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma2_bench
// profiles/tools/ffma2_bench.cu ; run on the GPU box.  Decides whether packing two waypoints per lane pays in K2.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float s = 0.f;
    if (MODE == 0) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
#pragma unroll 1
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
    } else {
        unsigned long long acc[8], pa, pb;
        float2 fa = make_float2(a, a), fb = make_float2(b, b);
        pa = *reinterpret_cast<unsigned long long*>(&fa);
        pb = *reinterpret_cast<unsigned long long*>(&fb);
        unsigned u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { float2 v = make_float2(threadIdx.x + i, threadIdx.x - i); acc[i] = *reinterpret_cast<unsigned long long*>(&v); u[i] = threadIdx.x * 7 + i; }
#pragma unroll 1
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i] = ffma2(acc[i], pa, pb);
                    if (MODE == 2) u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e3779b9u;   // LOP3/SHF + IADD on the ALU pipe
                    if (MODE == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(iters), "r"(it));   // one LOP3 per FFMA2
                }
#pragma unroll
        for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&acc[i]); s += v.x + v.y + (float)u[i]; }
    }
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, double flop_per_iter_thread) {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096; float best = 1e9f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double thr = 148.0 * 8 * 256;
    printf("%-28s %.3f ms  %.1f TFLOP/s  (%.2f warp-instr/clk/SMSP at 1.965 GHz)\n", name, best, flop_per_iter_thread * iters * thr / (best * 1e-3) / 1e12,
           (MODE == 0 ? 64.0 : 32.0) * iters * thr / 32 / (best * 1e-3) / 1.965e9 / (148 * 4));
    cudaFree(d);
}
int main() {
    run<0>("FFMA (16 chains)", 128);
    run<1>("FFMA2 (8 packed chains)", 128);
    run<2>("FFMA2 + 3 ALU interleaved", 128);
    run<3>("FFMA2 + 1 LOP3 interleaved", 128);
    return 0;
}

// ---- FP64 rate (is a DFMA-accumulated dot cheaper than the double-float FP32 one?) ----
__global__ void __launch_bounds__(256) kd(double* out, int iters, double a, double b, const float* xin) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) kcvt(double* out, int iters, float x0) {
    float x[8]; double acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = x0 + threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc = fma((double)x[i], (double)x[(i + 1) & 7], acc); x[i] += 1.f; }
    if (acc == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
struct RunD { RunD() {
    double* d; cudaMalloc(&d, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1024; const double thr = 148.0 * 8 * 256;
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            if (mode == 0) kd<<<148 * 8, 256>>>(d, iters, 0.999, 0.001, nullptr); else kcvt<<<148 * 8, 256>>>(d, iters, 1.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-28s %.3f ms  %.2f cycles per warp-DFMA per SMSP\n", mode == 0 ? "DFMA (8 chains)" : "2x F2F.F64 + DFMA + FADD", best,
               best * 1e-3 * 1.965e9 * (148 * 4) / (32.0 * iters * thr / 32));
    }
} } run_d;
