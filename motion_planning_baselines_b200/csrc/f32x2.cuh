// Packed FP32 pairs for sm_100a: fma/add/sub/mul.rn.f32x2 compile to FFMA2 / FADD2 / FMUL2, which do two IEEE
// single-precision operations per issue slot (measured on B200, profiles/tools/ffma2_bench.cu: the same 70 TFLOP/s as
// FFMA at half the issued instructions).  Each half rounds exactly like the scalar instruction, so a packed
// computation is bit-identical to the scalar one done twice.  ptxas folds a pair built from one scalar ({a,a}) into a
// broadcast operand (`R.F32`) and a negated scalar into the operand modifier, so constants need no duplication.
// Used by the issue-bound cost kernel to evaluate two waypoints per lane.  No counterpart in the reference.
#pragma once
#include <cuda_runtime.h>

namespace mpb {

__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

#define MPB_F32X2_BINARY(name, op)                                                        \
    __device__ __forceinline__ float2 name(float2 a, float2 b) {                          \
        float2 d;                                                                         \
        asm("{.reg .b64 ra, rb, rd;\n\t"                                                  \
            "mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5};\n\t" op " rd, ra, rb;\n\t"         \
            "mov.b64 {%0,%1}, rd;}"                                                       \
            : "=f"(d.x), "=f"(d.y)                                                        \
            : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                                    \
        return d;                                                                         \
    }
MPB_F32X2_BINARY(add2, "add.rn.f32x2")
MPB_F32X2_BINARY(sub2, "sub.rn.f32x2")
MPB_F32X2_BINARY(mul2, "mul.rn.f32x2")
#undef MPB_F32X2_BINARY

__device__ __forceinline__ float2 fma2(float2 a, float b, float2 c) { return fma2(a, bc2(b), c); }
__device__ __forceinline__ float2 fma2(float a, float2 b, float2 c) { return fma2(bc2(a), b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float b) { return mul2(a, bc2(b)); }
__device__ __forceinline__ float2 sub2(float2 a, float b) { return sub2(a, bc2(b)); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

}  // namespace mpb
