// In-kernel Gaussian noise: Philox4x32-10 (Salmon et al., SC'11) keyed on the GLOBAL element index + Box-Muller.
//
// Replaces the noise the reference draws inside MultivariateNormal.rsample (torch.randn) for
//   MultiMPPrior.sample              mp_baselines/planners/costs/factors/mp_priors_multi.py:253-256  (virtual tensor [S,P,M])
//   STOMP.sample                     mp_baselines/planners/stomp.py:97-108                         (virtual tensor [S,D,P,H])
//   ControlTrajectoryGaussian.sample mp_baselines/planners/priors/gaussian.py:276-298              (virtual tensor [C,N,T])
// Element e of the virtual GLOBAL noise tensor (flat index over the whole job: all particles of all GPUs, all samples
// of all ranks) is output e % 4 of  Philox(counter = (e / 4, offset), key = seed)  pushed through Box-Muller, so a
// result does not depend on how particles or samples are sharded over GPUs, CTAs or threads.  mpb_philox_normal dumps
// exactly these values (same device function) so that the oracle can replay a run on identical noise; oracle/philox.py
// restates the generator in numpy (integer part pinned bit-exactly by the Random123 known-answer vectors).
#pragma once
#include <stdint.h>

#include "../../include/mpb.h"

namespace mpb {

struct NoiseArgs {          // by-value kernel argument (from mpb_noise_desc)
    uint32_t k0, k1;        // seed
    uint32_t o0, o1;        // offset (draw counter)
    long long s_off, p_off, P_glob;
};

// mpb_noise_desc -> NoiseArgs; `n_local` = particles (MPPI: samples) of this call.  Returns nullptr or a message.
inline const char* noise_args(const mpb_noise_desc& d, long long n_local, NoiseArgs& out) {
    if (d.s_offset < 0 || d.p_offset < 0) return "negative noise offsets";
    if (d.P_global < 1) return "noise P_global must be >= 1";
    (void)n_local;
    out.k0 = (uint32_t)d.seed; out.k1 = (uint32_t)(d.seed >> 32);
    out.o0 = (uint32_t)d.offset; out.o1 = (uint32_t)(d.offset >> 32);
    out.s_off = d.s_offset; out.p_off = d.p_offset; out.P_glob = d.P_global;
    return nullptr;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// (u, v) uniform uint32 -> two independent standard normals.  a in (0,1], angle in [-pi, pi): the fast MUFU paths are
// used on purpose (lg2 / sin / cos approximations: absolute error ~5e-7, far below the sampler's own rounding) --
// the dump entry runs the very same code, so replays are bit-identical.
__device__ __forceinline__ float2 box_muller(uint32_t u, uint32_t v) {
    const float a = fmaf(__uint2float_rn(u), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float t = fmaf(__uint2float_rn(v), 1.4629180792671596e-09f, -3.1415925803542134f);   // 2 pi 2^-32 v + (pi 2^-32 - pi)
    // a >= 2^-33 and -2 ln a is 0 or >= 2^-24: never denormal, so the .ftz forms return the same bits as the plain ones and
    // spare the denormal pre-scaling (compare, two multiplies, one add per MUFU) the compiler wraps around those
    float l2, r;                                                                               // sqrt(-2 ln a)
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(a));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));              // MUFU.SQRT alone: 1e-7 relative
    float sn, cs;
    __sincosf(t, &sn, &cs);
    return make_float2(r * cs, r * sn);
}

// The four normals of group `grp` (elements 4*grp .. 4*grp+3 of the virtual global noise tensor).
__device__ __forceinline__ float4 philox_normal4(unsigned long long grp, const NoiseArgs& n) {
    const uint4 u = philox4x32_10((uint32_t)grp, (uint32_t)(grp >> 32), n.o0, n.o1, n.k0, n.k1);
    const float2 a = box_muller(u.x, u.y), b = box_muller(u.z, u.w);
    return make_float4(a.x, a.y, b.x, b.y);
}

// One element (used where a thread needs a single value and groups straddle its work; costs a full Philox call).
__device__ __forceinline__ float philox_normal1(unsigned long long e, const NoiseArgs& n) {
    const float4 q = philox_normal4(e >> 2, n);
    const int l = (int)(e & 3ull);
    return l == 0 ? q.x : (l == 1 ? q.y : (l == 2 ? q.z : q.w));
}

}  // namespace mpb
