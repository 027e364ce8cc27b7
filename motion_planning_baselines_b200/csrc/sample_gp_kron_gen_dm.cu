// K1, dof-major variant of the Blackwell sampler (sample_gp_kron_gen.cu) for the FUSED Stoch-GPMP iteration:
//     x[p,s,:] = mu[p,:] + L @ eps[s,p,:],   eps ~ N(0, I) generated on the fly
//
// Replaces MultiMPPrior.sample (mp_baselines/planners/costs/factors/mp_priors_multi.py:253-256) including the noise draw.
// Same arithmetic as sample_gp_kron_gen_kernel (two-term fp16 split, three kind::f16 MMAs per k-step in the same order,
// fp32 accumulation in tensor memory, x = fma(acc, 1 / scale, mu)) and the same noise (layout MPB_NOISE_SPMD): the
// samples are bit-identical.  What differs is the LAYOUT of the sample rows it writes and, with it, the shape of the work:
//
//   natural row (the reference's, [H][2 dof]):  column 14 h + 7 pv + j         (pv = 0 position, 1 velocity; j = dof)
//   dof-major row (this kernel, [dof][2H]):     column 128 j + 2 h + pv  =  128 j + n,  n = accumulator row of dof j
//
// In the natural layout a finished row needs all seven accumulators of a sample at once; they only fit tensor memory
// for 32-sample tiles (two sets), and an M128 x N32 x K16 MMA costs the same 55 cycles as an N64 one because each MMA
// re-reads its 4 KiB factor tile from shared memory -- the MMAs themselves bounded that kernel (0.047 of 0.062 ms).  In the
// dof-major layout the dofs are independent all the way to global memory: a work unit is (64-sample tile, ONE dof), its
// accumulator is 64 columns (eight units fit tensor memory: an 8-deep ring between the MMAs and the epilogue), the
// epilogue thread of accumulator row n stores x[s][128 j + n] for its 64 samples -- a warp writes 128 contiguous bytes
// per sample, no staging in shared memory, no bulk stores -- and the factor of a dof (64 KiB, hi + lo) is loaded into
// shared memory ONCE per CTA and dof (units are dealt in dof-major order, so a CTA sees one or two dofs) instead of being
// streamed once per tile: half the MMA time per sample, 1 / 50 of the factor traffic.  The kernel is then bound by the noise
// generation (Philox4x32-10 + Box-Muller + fp16 split on the CUDA cores), which gets as many producer warps as a CTA can hold.
//
// Consumers: the cost kernel and the update kernel read dof-major rows directly (mpb_cost_eval_dm, mpb_softmax_update_dm);
// mpb_traj_from_dof_major converts for everything else (state_samples accessors, tests).
//
// Two kernels share this pipeline (bit-identical results, tests/test_gpu_dof_major.py runs both):
//   sample_gp_kron_gen_dm3_kernel (further down, the DEFAULT): 8 epilogue + 24 producer warps, the loader / MMA / mat-vec
//                roles folded into them -- 55.0 us per launch at C4;
//   sample_gp_kron_gen_dm_kernel (next, MPB_DM_VARIANT=2): dedicated loader / MMA / mat-vec warps, 16 producer warps, with the
//                stage-disable and clock64 trace hooks the measurements in profiles/r02_k1_dm.txt were taken with -- 57.1 us.
// sample_gp_kron_gen_dm_kernel: one persistent CTA per SM, warp-specialised (864 threads):
//   warps 0-7    epilogue (TMEM lane quadrant x half of the columns): tcgen05.ld 32 columns, fma with mu, 32 coalesced stores
//   warp 8       factor loader (one lane: eight 8 KiB bulk-async copies per dof, double buffered, mbarrier complete_tx)
//   warp 9       MMA issuer (warp-uniform loop, one elected lane): 24 tcgen05.mma.kind::f16 (M128 x N64 x K16) per unit
//   warps 10-25  noise producers, two groups of eight filling alternate units: eight Philox calls per thread and unit,
//                8-byte conflict-free stores into the canonical K-major (no swizzle) operand tiles of a 3-stage ring (a
//                stage = the whole K = 128 of a unit, 32 KiB)
//   warp 26      (optional) y_p = Sigma^-1 mu_p and the copy of mu, as in sample_gp_kron_gen_kernel
#include <cuda_fp16.h>

#include <cstdlib>

#include "mpb_common.cuh"
#include "philox.cuh"
#include "tcgen05.cuh"

namespace mpb {

template <int DOF>
struct GenDmCfg {
    static constexpr int NOUT = 128;                         // 2H = rows of a per-dof block (UMMA M); H = 64
    static constexpr int TS = 64;                            // samples per tile (UMMA N)
    static constexpr int KC = 16, NKC = NOUT / KC;
    static constexpr int M = NOUT * DOF;
    static constexpr uint32_t B_TILE = TS * KC * 2;          // noise tile of one (chunk, part): 2 KiB
    static constexpr uint32_t A_TILE = NOUT * KC * 2;        // factor tile of one (chunk, part): 4 KiB
    static constexpr uint32_t B_STAGE = NKC * 2 * B_TILE;    // one unit: 32 KiB
    static constexpr uint32_t A_BUF = NKC * 2 * A_TILE;      // one dof: 64 KiB
    static constexpr uint32_t A_IMG_STAGE = DOF * 2 * A_TILE;    // stride between k-chunks in the host image (sample_gp_kron_gen.cu)
#ifndef MPB_DM_GROUPS
#define MPB_DM_GROUPS 2
#endif
    // producer groups filling alternate units; 2 groups: 3-stage noise ring + double-buffered factor, 4 groups: 5-stage ring +
    // one factor buffer (a dof switch then waits for the 64 KiB load once; the deeper ring keeps the producers busy meanwhile).
    // Measured at C4 (profiles/r02_k1_dm.txt): 4 groups pace the units 8 % faster (3 340 vs 3 640 cycles) but take twice as
    // long to fill the pipeline (first MMA after 16 500 instead of 7 000 cycles): 56.3 vs 55.2 us per launch -- 2 it is.
    // 3 groups of 8 (24 producer warps; the 1 024-thread CTA then leaves room for only 4 epilogue warps): the producers pace
    // the units 21 % faster (2 870 cycles) but four epilogue warps need 3 500 per unit and become the bound: 68.2 us.
    static constexpr int NGROUPS = MPB_DM_GROUPS;
    static constexpr int B_STAGES = NGROUPS >= 3 ? 5 : 3, A_BUFS = NGROUPS >= 3 ? 1 : 2, NSETS = 8;
    static constexpr int GW = NGROUPS == 4 ? 4 : 8;          // producer warps per group (3 groups: 24 producer warps, 4 epilogue warps)
    static constexpr int EPI_WARPS = NGROUPS == 3 ? 4 : 8, LOAD_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1, FIRST_PROD_WARP = EPI_WARPS + 2;
    static constexpr int PROD_WARPS = NGROUPS * GW;
    static constexpr int MV_WARP = FIRST_PROD_WARP + PROD_WARPS;
    static constexpr int THREADS = (MV_WARP + 1) * 32;
    static constexpr uint32_t OFF_A = 0;
    static constexpr uint32_t OFF_B = OFF_A + A_BUFS * A_BUF;
    static constexpr uint32_t OFF_BAR = OFF_B + B_STAGES * B_STAGE;
    static constexpr uint32_t SMEM = OFF_BAR + 256 + 128 /*alignment slack*/;
    static constexpr uint32_t TMEM_COLS = 512;
    static_assert(NSETS * TS <= 512, "accumulator ring must fit tensor memory");
    static_assert(SMEM <= 227 * 1024, "shared-memory budget");
};

struct GenDmArgs {
    const unsigned char* Limg;      // the image mpb_sample_gp_kron_gen_prepare builds: [NKC][DOF][hi|lo][A_TILE] + inverse scales
    const float* mu;                // [P, M] natural layout
    float* x;                       // [P, S, M] DOF-MAJOR rows
    int P, S;
    long long Ntot;                 // P * S rows
    int ntiles;
    const float* Sinv;              // optional, as in GenArgs
    float* y;
    float* mu_copy;
    long long* trace;               // optional [256] clock stamps of CTA 0 (MPB_KRON_DM_TRACE = device pointer; timing experiments)
    int dbg;                        // MPB_KRON_DM_DBG bit mask (timing experiments only): 1 no Philox (zeros), 2 no MMAs, 4 no stores, 8 no operand stores, 16 no TMEM loads, 32 no mat-vec warp
};

__device__ __forceinline__ void split4_f16_dm(const float4 v, uint2& hi, uint2& lo) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(__fsub_rn(v.x, f01.x), __fsub_rn(v.y, f01.y));
    const __half2 l23 = __floats2half2_rn(__fsub_rn(v.z, f23.x), __fsub_rn(v.w, f23.y));
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

__device__ __forceinline__ void umma_f16_dm(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ bool elect_one_dm() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void st_shared_v2_dm(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

template <int DOF>
__global__ void __launch_bounds__(GenDmCfg<DOF>::THREADS, 1)
sample_gp_kron_gen_dm_kernel(const GenDmArgs a, const NoiseArgs noise) {
    using C = GenDmCfg<DOF>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* a_full = bars;                           // [A_BUFS] factor of a dof landed
    uint64_t* a_empty = a_full + C::A_BUFS;            // [A_BUFS] the MMAs of the last unit that read it completed
    uint64_t* b_full = a_empty + C::A_BUFS;            // [B_STAGES] noise of a unit written (PROD_WARPS / 2 arrivals: one group)
    uint64_t* b_empty = b_full + C::B_STAGES;          // [B_STAGES]
    uint64_t* acc_full = b_empty + C::B_STAGES;        // [NSETS] accumulator of a unit complete
    uint64_t* acc_empty = acc_full + C::NSETS;         // [NSETS] accumulator read by the eight epilogue warps
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(acc_empty + C::NSETS);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool spin = (a.dbg & 64) != 0;
    auto WAIT = [&](uint64_t* bar, uint32_t parity) { if (spin) mbar_wait_spin(bar, parity); else mbar_wait(bar, parity); };
    auto stamp = [&](int ordinal, int ev) {
        if (a.trace && blockIdx.x == 0 && ordinal < 24) a.trace[8 * ordinal + ev] = clock64();   // ev >= 128: second bank (slot + 24 * 8)
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::A_BUFS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < C::B_STAGES; ++s) { mbar_init(&b_full[s], C::PROD_WARPS / C::NGROUPS); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < C::NSETS; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], C::EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == C::MMA_WARP) tmem_alloc(tmem_base_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    pdl_wait();                                        // the means come from the previous iteration's update kernel

    // Units u = j * ntiles + t (dof-major) are dealt in contiguous, balanced ranges: CTA c runs [u0, u1).
    const long long U = (long long)DOF * a.ntiles;
    const int u0 = (int)(U * blockIdx.x / gridDim.x), u1 = (int)(U * (blockIdx.x + 1) / gridDim.x);

    if (warp == C::LOAD_WARP) {
        // ================================ factor loader ================================
        if (lane == 0) {
            int buf = 0, jprev = -1;
            uint32_t phase = 0;
            for (int u = u0; u < u1; ++u) {
                const int j = u / a.ntiles;
                if (j == jprev) continue;
                jprev = j;
                WAIT(&a_empty[buf], phase ^ 1);
                mbar_expect_tx(&a_full[buf], C::A_BUF);
                for (int kc = 0; kc < C::NKC; ++kc)
                    bulk_load(sm + C::OFF_A + buf * C::A_BUF + kc * (2 * C::A_TILE),
                              a.Limg + (size_t)kc * C::A_IMG_STAGE + (size_t)j * (2 * C::A_TILE), 2 * C::A_TILE, &a_full[buf]);
                if (++buf == C::A_BUFS) { buf = 0; phase ^= 1; }
            }
        }
    } else if (warp == C::MMA_WARP) {
        // ================================ MMA issuer ===================================
        const bool leader = elect_one_dm();
        const uint32_t idesc = make_idesc_f16(C::NOUT, C::TS);
        const uint64_t adesc0 = make_nosw_desc(smem_u32(sm + C::OFF_A), 128u, 256u);
        const uint64_t bdesc0 = make_nosw_desc(smem_u32(sm + C::OFF_B), 128u, 256u);
        const uint32_t adesc_lo = (uint32_t)adesc0, bdesc_lo = (uint32_t)bdesc0, desc_hi = (uint32_t)(adesc0 >> 32);
        int abuf = -1, bs = 0, jprev = -1;
        uint32_t aph = 0, bph = 0;
        int ord = 0;
        for (int u = u0; u < u1; ++u, ++ord) {
            const int j = u / a.ntiles;
            const int set = ord % C::NSETS;
            if (j != jprev) {
                if (jprev >= 0 && leader) umma_commit(&a_empty[abuf]);      // every MMA on the previous factor has been issued
                __syncwarp();
                jprev = j;
                if (++abuf == C::A_BUFS) { abuf = 0; aph ^= 1; }
                WAIT(&a_full[abuf], aph);
            }
            WAIT(&acc_empty[set], (((uint32_t)(ord / C::NSETS)) & 1u) ^ 1u);
            WAIT(&b_full[bs], bph);
            tc_fence_after();
            if (lane == 0) stamp(ord, 2);
            const uint32_t d = tmem_base + (uint32_t)(set * C::TS);
            const uint32_t a_w = adesc_lo + (uint32_t)((abuf * C::A_BUF) >> 4);
            const uint32_t b_w = bdesc_lo + (uint32_t)((bs * C::B_STAGE) >> 4);
            if (leader) {
#pragma unroll
                for (int kc = 0; kc < C::NKC && !(a.dbg & 2); ++kc) {
                    const uint32_t ahi = a_w + (uint32_t)((kc * 2 * C::A_TILE) >> 4), alo = ahi + (C::A_TILE >> 4);
                    const uint32_t bhi = b_w + (uint32_t)((kc * 2 * C::B_TILE) >> 4), blo = bhi + (C::B_TILE >> 4);
                    umma_f16_dm(d, alo, bhi, desc_hi, idesc, kc == 0 ? 0u : 1u);          // small terms first
                    umma_f16_dm(d, ahi, blo, desc_hi, idesc, 1u);
                    umma_f16_dm(d, ahi, bhi, desc_hi, idesc, 1u);
                }
                umma_commit(&b_empty[bs]);
                umma_commit(&acc_full[set]);
            }
            __syncwarp();
            if (lane == 0) stamp(ord, 3);
            if (++bs == C::B_STAGES) { bs = 0; bph ^= 1; }
        }
    } else if (warp < C::EPI_WARPS) {
        // ================================ epilogue =====================================
        // warp (q4, half): TMEM lane quadrant q4, columns (= samples) 32 half .. 32 half + 31 of the unit's accumulator
        constexpr int NH = 8 / C::EPI_WARPS;                    // halves (32 columns) a warp handles per unit: 1 or 2
        const int q4 = warp & 3;
        const int n_out = 32 * q4 + lane;                       // TMEM lane = accumulator row = column inside the dof block
        const float* inv_scale_g = reinterpret_cast<const float*>(a.Limg + (size_t)C::NKC * C::A_IMG_STAGE);
        const bool one_particle = (a.S % C::TS) == 0;           // a tile never straddles two particles
        const bool st_on = !(a.dbg & 4);
        int ord = 0;
        for (int u = u0; u < u1; ++u, ++ord) {
            const int j = u / a.ntiles, t = u - j * a.ntiles;
            const int set = ord % C::NSETS;
            const float inv_scale = __ldg(inv_scale_g + j);
            const int mcol = DOF * n_out + j;                   // natural column of (row n, dof j)
            WAIT(&acc_full[set], ((uint32_t)(ord / C::NSETS)) & 1u);
            tc_fence_after();
            if (threadIdx.x == 0) stamp(ord, 4);
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {
            const int half = (warp >> 2) + hh;
            const long long row0 = (long long)t * C::TS + 32 * half;        // first row of this half
            long long rows_ll = a.Ntot - row0;
            const int rows = rows_ll > 32 ? 32 : (int)rows_ll;  // may be <= 0 in the last tile
            // particle of the first row; it advances where a row index crosses a multiple of S (no division per row: the
            // 64-bit divide is a 150-cycle subroutine)
            const long long rowc = rows > 0 ? row0 : a.Ntot - 1;
            int p = (int)(rowc / a.S), rem = (int)(rowc - (long long)p * a.S);
            float m = __ldg(a.mu + (size_t)p * C::M + mcol);
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(set * C::TS + 32 * half);
            float v[32];
            if (!(a.dbg & 16)) {
                tmem_ld32(taddr, v);
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = 0.f;
            }
            if (hh == NH - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[set]);    // the accumulator is in registers: hand it back before the stores
            }
            float* xo = a.x + (size_t)row0 * C::M + (size_t)j * C::NOUT + n_out;
            if (one_particle && rows == 32) {
                if (st_on) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) xo[(size_t)c * C::M] = fmaf(v[c], inv_scale, m);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (rem == a.S) {
                        rem = 0;
                        ++p;
                        if (p < a.P) m = __ldg(a.mu + (size_t)p * C::M + mcol);
                    }
                    ++rem;
                    if (c < rows && st_on) xo[(size_t)c * C::M] = fmaf(v[c], inv_scale, m);
                }
            }
            }
            if (threadIdx.x == 0) stamp(ord, 5);
        }
    } else if (warp >= C::FIRST_PROD_WARP && warp < C::MV_WARP) {
        // ================================ noise producers ==============================
        // Two GROUPS of eight producer warps fill alternate units (group g: units u0 + g, u0 + g + 2, ...).  With one group
        // of sixteen, all producers ran in lockstep through the per-unit overhead (barrier wait, index arithmetic, operand
        // stores, arrive: ~1 400 of 3 700 cycles per unit, measured with the generation switched off) and every scheduler
        // idled through it; two groups are half a unit out of phase, so one computes while the other is in its overhead,
        // and the overhead is paid once per eight Philox calls of a thread instead of once per four.
        // thread = one sample row (8-sample group sgrp, row r8) x four k-chunks (kh) x the 4-k groups of parity par
        const int pw = warp - C::FIRST_PROD_WARP;
        constexpr int GW = C::GW;                                 // warps per group: 8 or 4
        constexpr int KH = GW / 4;                                // k-halves a group splits a unit into: 2 or 1
        constexpr int NB = 4 / KH;                                // batches of four Philox calls per thread and unit: 2 or 4
        const int grp = pw / GW, gw = pw % GW;
        const int r8 = lane & 7, par = (lane >> 3) & 1, g1 = lane >> 4;
        const int sgrp = (gw & 3) * 2 + g1, kh = gw >> 2;
        for (int u = u0 + grp; u < u1; u += C::NGROUPS) {
            const int ord = u - u0;
            const int bs = ord % C::B_STAGES;
            const uint32_t bph = (uint32_t)(ord / C::B_STAGES) & 1u;
            const int j = u / a.ntiles, t = u - j * a.ntiles;
            long long n = (long long)t * C::TS + sgrp * 8 + r8;
            if (n >= a.Ntot) n = a.Ntot - 1;                      // rows past the end are never stored
            const long long p = n / a.S, s = n - p * a.S;
            const unsigned long long grow = (unsigned long long)((((noise.s_off + s) * noise.P_glob + noise.p_off + p) * DOF + j) * (C::NOUT / 4));
            const uint32_t stage = smem_u32(sm + C::OFF_B + bs * C::B_STAGE);
            bool waited = false;
#pragma unroll 1
            for (int b = 0; b < NB; ++b) {                        // batches of four Philox calls (interleaved chains)
                uint2 hi[4], lo[4];
                if (!(a.dbg & 1)) {
                    float4 e[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int kc = (8 / KH) * kh + 2 * b + (i >> 1), q = 2 * (i & 1) + par;
                        e[i] = philox_normal4(grow + (unsigned)(kc * 4 + q), noise);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) split4_f16_dm(e[i], hi[i], lo[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) hi[i] = lo[i] = make_uint2(0u, 0u);
                }
                if (!waited) {
                    if (gw == 0 && lane == 0) stamp(ord, 0);
                    WAIT(&b_empty[bs], bph ^ 1);
                    if (gw == 0 && lane == 0) stamp(ord, 6);
                    waited = true;
                }
#pragma unroll
                for (int i = 0; i < 4 && !(a.dbg & 8); ++i) {
                    const int kc = (8 / KH) * kh + 2 * b + (i >> 1), q = 2 * (i & 1) + par;
                    // (sample row, k) -> core matrix (row / 8, k / 8): 16-byte rows, K groups 128 B apart, row groups 256 B
                    const uint32_t off = stage + (uint32_t)(kc * (2 * C::B_TILE)) + (uint32_t)(sgrp * 256 + (q >> 1) * 128 + r8 * 16 + (q & 1) * 8);
                    st_shared_v2_dm(off, hi[i]);
                    st_shared_v2_dm(off + C::B_TILE, lo[i]);
                }
            }
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_full[bs]);
            if (gw == 0 && lane == 0) stamp(ord, 1);
        }
    } else if (warp == C::MV_WARP && a.y && !(a.dbg & 32)) {
        // ================================ Sigma^-1 mu ===================================
        // as in sample_gp_kron_gen_kernel (same arithmetic as prior_matvec_dof_kernel: y is bit-identical), natural layout
        constexpr int MAXP = 4;
        for (int pb = blockIdx.x; pb < a.P; pb += MAXP * gridDim.x) {
            int np = 0;
            const float* mrow[MAXP];
#pragma unroll
            for (int q = 0; q < MAXP; ++q) {
                const int p = pb + q * gridDim.x;
                mrow[q] = a.mu + (size_t)(p < a.P ? p : pb) * C::M;
                if (p < a.P) np = q + 1;
            }
#pragma unroll 1
            for (int i = lane; i < C::M; i += 32) {
                float sv[7];
#pragma unroll
                for (int m = 0; m < 7; ++m) {
                    const int jj = i + (m - 3) * DOF;
                    sv[m] = (jj >= 0 && jj < C::M) ? __ldg(a.Sinv + (size_t)jj * C::M + i) : 0.f;
                }
                float mv[MAXP][7];
#pragma unroll
                for (int q = 0; q < MAXP; ++q)
#pragma unroll
                    for (int m = 0; m < 7; ++m) {
                        const int jj = i + (m - 3) * DOF;
                        mv[q][m] = (jj >= 0 && jj < C::M) ? __ldg(mrow[q] + jj) : 0.f;
                    }
#pragma unroll
                for (int q = 0; q < MAXP; ++q) {
                    float hi = 0.f, lo = 0.f;
#pragma unroll
                    for (int m = 0; m < 7; ++m) {
                        const float pr = __fmul_rn(sv[m], mv[q][m]);
                        const float e = fmaf(sv[m], mv[q][m], -pr);
                        const float t = __fadd_rn(hi, pr);
                        const float z = __fsub_rn(t, hi);
                        lo = __fadd_rn(lo, __fadd_rn(__fadd_rn(__fsub_rn(hi, __fsub_rn(t, z)), __fsub_rn(pr, z)), e));
                        hi = t;
                    }
                    if (q < np) {
                        a.y[(size_t)(pb + q * gridDim.x) * C::M + i] = __fadd_rn(hi, lo);
                        if (a.mu_copy) a.mu_copy[(size_t)(pb + q * gridDim.x) * C::M + i] = mv[q][3];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C::MMA_WARP) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---- variant with 24 producer warps (the default; MPB_DM_VARIANT=2 selects the kernel above) ---------------------------------
// Measured at C4: 55.0 us per launch against 57.1, step 0.2378 against 0.2401 ms (profiles/r02_k1_dm.txt).
// The generation is bound by thread-level parallelism (three groups of eight producer warps pace the units 21 % faster than
// two), but a CTA holds 32 warps: with a loader, an MMA and a mat-vec warp next to eight epilogue warps only 21 are left.
// Here the dedicated warps are folded into the others: warps 0-7 are the epilogue (and share the Sigma^-1 mu rows at the
// start), warps 8-31 are three producer groups; group g owns operand stage g; warp 0 of a group issues the MMAs of the
// unit its group has just written (it waits for its group's arrivals, the accumulator and the factor, then one elected lane
// issues the 24 MMAs); the factors of the CTA's (at most two) dofs are bulk-loaded at the start by the first producer warp.
template <int DOF>
struct GenDm3Cfg {
    static constexpr int NOUT = 128, TS = 64, KC = 16, NKC = NOUT / KC, M = NOUT * DOF;
    static constexpr uint32_t B_TILE = TS * KC * 2, A_TILE = NOUT * KC * 2;
    static constexpr uint32_t B_STAGE = NKC * 2 * B_TILE, A_BUF = NKC * 2 * A_TILE, A_IMG_STAGE = DOF * 2 * A_TILE;
    static constexpr int NGROUPS = 3, GW = 8, B_STAGES = NGROUPS, A_BUFS = 2, NSETS = 8;
    static constexpr int EPI_WARPS = 8, FIRST_PROD_WARP = EPI_WARPS, PROD_WARPS = NGROUPS * GW;
    static constexpr int THREADS = (EPI_WARPS + PROD_WARPS) * 32;
    static constexpr uint32_t OFF_A = 0, OFF_B = OFF_A + A_BUFS * A_BUF, OFF_BAR = OFF_B + B_STAGES * B_STAGE;
    static constexpr uint32_t SMEM = OFF_BAR + 256 + 128;
    static constexpr uint32_t TMEM_COLS = 512;
    static_assert(THREADS == 1024 && SMEM <= 227 * 1024, "one full CTA per SM");
};

template <int DOF>
__global__ void __launch_bounds__(GenDm3Cfg<DOF>::THREADS, 1)
sample_gp_kron_gen_dm3_kernel(const GenDmArgs a, const NoiseArgs noise) {
    using C = GenDm3Cfg<DOF>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* a_full = bars;                           // [A_BUFS] factor of the CTA's first / second dof landed (loaded once)
    uint64_t* b_full = a_full + C::A_BUFS;             // [B_STAGES] noise of a unit written (GW arrivals: the stage's group)
    uint64_t* b_empty = b_full + C::B_STAGES;          // [B_STAGES] the MMAs that read the stage completed
    uint64_t* acc_full = b_empty + C::B_STAGES;        // [NSETS]
    uint64_t* acc_empty = acc_full + C::NSETS;         // [NSETS] EPI_WARPS arrivals
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(acc_empty + C::NSETS);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::A_BUFS; ++s) mbar_init(&a_full[s], 1);
        for (int s = 0; s < C::B_STAGES; ++s) { mbar_init(&b_full[s], C::GW); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < C::NSETS; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], C::EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_base_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    pdl_wait();

    const long long U = (long long)DOF * a.ntiles;
    const int u0 = (int)(U * blockIdx.x / gridDim.x), u1 = (int)(U * (blockIdx.x + 1) / gridDim.x);
    const int j0 = u0 / a.ntiles;                      // the range [u0, u1) holds dof j0 and possibly j0 + 1 (the host checks)

    if (warp < C::EPI_WARPS) {
        // ================================ Sigma^-1 mu, then the epilogue ================
        if (a.y) {      // same arithmetic as the mat-vec warp of the other variants; the rows are shared by the eight warps
            constexpr int MAXP = 4;
            for (int pb = blockIdx.x; pb < a.P; pb += MAXP * gridDim.x) {
                int np = 0;
                const float* mrow[MAXP];
#pragma unroll
                for (int q = 0; q < MAXP; ++q) {
                    const int p = pb + q * gridDim.x;
                    mrow[q] = a.mu + (size_t)(p < a.P ? p : pb) * C::M;
                    if (p < a.P) np = q + 1;
                }
#pragma unroll 1
                for (int i = lane + 32 * warp; i < C::M; i += 32 * C::EPI_WARPS) {
                    float sv[7];
#pragma unroll
                    for (int m = 0; m < 7; ++m) {
                        const int jj = i + (m - 3) * DOF;
                        sv[m] = (jj >= 0 && jj < C::M) ? __ldg(a.Sinv + (size_t)jj * C::M + i) : 0.f;
                    }
#pragma unroll
                    for (int q = 0; q < MAXP; ++q) {
                        float mv[7];
#pragma unroll
                        for (int m = 0; m < 7; ++m) {
                            const int jj = i + (m - 3) * DOF;
                            mv[m] = (jj >= 0 && jj < C::M) ? __ldg(mrow[q] + jj) : 0.f;
                        }
                        float hi = 0.f, lo = 0.f;
#pragma unroll
                        for (int m = 0; m < 7; ++m) {
                            const float pr = __fmul_rn(sv[m], mv[m]);
                            const float e = fmaf(sv[m], mv[m], -pr);
                            const float t = __fadd_rn(hi, pr);
                            const float z = __fsub_rn(t, hi);
                            lo = __fadd_rn(lo, __fadd_rn(__fadd_rn(__fsub_rn(hi, __fsub_rn(t, z)), __fsub_rn(pr, z)), e));
                            hi = t;
                        }
                        if (q < np) {
                            a.y[(size_t)(pb + q * gridDim.x) * C::M + i] = __fadd_rn(hi, lo);
                            if (a.mu_copy) a.mu_copy[(size_t)(pb + q * gridDim.x) * C::M + i] = mv[3];
                        }
                    }
                }
            }
        }
        const int q4 = warp & 3, half = warp >> 2;
        const int n_out = 32 * q4 + lane;
        const float* inv_scale_g = reinterpret_cast<const float*>(a.Limg + (size_t)C::NKC * C::A_IMG_STAGE);
        const bool one_particle = (a.S % C::TS) == 0;
        int ord = 0;
        for (int u = u0; u < u1; ++u, ++ord) {
            const int j = u / a.ntiles, t = u - j * a.ntiles;
            const int set = ord % C::NSETS;
            const long long row0 = (long long)t * C::TS + 32 * half;
            const float inv_scale = __ldg(inv_scale_g + j);
            const int mcol = DOF * n_out + j;
            long long rows_ll = a.Ntot - row0;
            const int rows = rows_ll > 32 ? 32 : (int)rows_ll;
            const long long rowc = rows > 0 ? row0 : a.Ntot - 1;
            int p = (int)(rowc / a.S), rem = (int)(rowc - (long long)p * a.S);
            float m = __ldg(a.mu + (size_t)p * C::M + mcol);
            mbar_wait(&acc_full[set], ((uint32_t)(ord / C::NSETS)) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(set * C::TS + 32 * half);
            float v[32];
            tmem_ld32(taddr, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[set]);
            float* xo = a.x + (size_t)row0 * C::M + (size_t)j * C::NOUT + n_out;
            if (one_particle && rows == 32) {
#pragma unroll
                for (int c = 0; c < 32; ++c) xo[(size_t)c * C::M] = fmaf(v[c], inv_scale, m);
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (rem == a.S) {
                        rem = 0;
                        ++p;
                        if (p < a.P) m = __ldg(a.mu + (size_t)p * C::M + mcol);
                    }
                    ++rem;
                    if (c < rows) xo[(size_t)c * C::M] = fmaf(v[c], inv_scale, m);
                }
            }
        }
    } else {
        // ================================ noise producers (+ factor loads, MMA issue) ==
        const int pw = warp - C::FIRST_PROD_WARP;
        const int grp = pw / C::GW, gw = pw % C::GW;
        if (pw == 0 && lane == 0 && u0 < u1) {          // the factors of the CTA's dofs, once
            const int jl = (u1 - 1) / a.ntiles;
            for (int b = 0; b <= jl - j0 && b < C::A_BUFS; ++b) {
                mbar_expect_tx(&a_full[b], C::A_BUF);
                for (int kc = 0; kc < C::NKC; ++kc)
                    bulk_load(sm + C::OFF_A + b * C::A_BUF + kc * (2 * C::A_TILE),
                              a.Limg + (size_t)kc * C::A_IMG_STAGE + (size_t)(j0 + b) * (2 * C::A_TILE), 2 * C::A_TILE, &a_full[b]);
            }
        }
        const bool leader = elect_one_dm();
        const uint32_t idesc = make_idesc_f16(C::NOUT, C::TS);
        const uint64_t adesc0 = make_nosw_desc(smem_u32(sm + C::OFF_A), 128u, 256u);
        const uint64_t bdesc0 = make_nosw_desc(smem_u32(sm + C::OFF_B), 128u, 256u);
        const uint32_t adesc_lo = (uint32_t)adesc0, bdesc_lo = (uint32_t)bdesc0, desc_hi = (uint32_t)(adesc0 >> 32);
        const int r8 = lane & 7, par = (lane >> 3) & 1, g1 = lane >> 4;
        const int sgrp = (gw & 3) * 2 + g1, kh = gw >> 2;
        const int bs = grp;                                       // a group owns its operand stage
        const uint32_t stage = smem_u32(sm + C::OFF_B + bs * C::B_STAGE);
        uint32_t it = 0;                                          // uses of the stage so far
        for (int u = u0 + grp; u < u1; u += C::NGROUPS, ++it) {
            const int ord = u - u0;
            const int j = u / a.ntiles, t = u - j * a.ntiles;
            long long n = (long long)t * C::TS + sgrp * 8 + r8;
            if (n >= a.Ntot) n = a.Ntot - 1;
            const long long p = n / a.S, s = n - p * a.S;
            const unsigned long long grow = (unsigned long long)((((noise.s_off + s) * noise.P_glob + noise.p_off + p) * DOF + j) * (C::NOUT / 4));
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
                uint2 hi[4], lo[4];
                float4 e[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int kc = 4 * kh + 2 * b + (i >> 1), q = 2 * (i & 1) + par;
                    e[i] = philox_normal4(grow + (unsigned)(kc * 4 + q), noise);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) split4_f16_dm(e[i], hi[i], lo[i]);
                if (b == 0) mbar_wait(&b_empty[bs], (it & 1u) ^ 1u);      // the MMAs of this group's previous unit are done
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int kc = 4 * kh + 2 * b + (i >> 1), q = 2 * (i & 1) + par;
                    const uint32_t off = stage + (uint32_t)(kc * (2 * C::B_TILE)) + (uint32_t)(sgrp * 256 + (q >> 1) * 128 + r8 * 16 + (q & 1) * 8);
                    st_shared_v2_dm(off, hi[i]);
                    st_shared_v2_dm(off + C::B_TILE, lo[i]);
                }
            }
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_full[bs]);
            if (gw == 0) {
                // this warp issues the unit's MMAs once its group has written the stage
                const int set = ord % C::NSETS;
                const int abuf = j - j0;
                mbar_wait(&b_full[bs], it & 1u);
                mbar_wait(&acc_empty[set], (((uint32_t)(ord / C::NSETS)) & 1u) ^ 1u);
                mbar_wait(&a_full[abuf], 0u);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(set * C::TS);
                const uint32_t a_w = adesc_lo + (uint32_t)((abuf * C::A_BUF) >> 4);
                const uint32_t b_w = bdesc_lo + (uint32_t)((bs * C::B_STAGE) >> 4);
                if (leader) {
#pragma unroll
                    for (int kc = 0; kc < C::NKC; ++kc) {
                        const uint32_t ahi = a_w + (uint32_t)((kc * 2 * C::A_TILE) >> 4), alo = ahi + (C::A_TILE >> 4);
                        const uint32_t bhi = b_w + (uint32_t)((kc * 2 * C::B_TILE) >> 4), blo = bhi + (C::B_TILE >> 4);
                        umma_f16_dm(d, alo, bhi, desc_hi, idesc, kc == 0 ? 0u : 1u);
                        umma_f16_dm(d, ahi, blo, desc_hi, idesc, 1u);
                        umma_f16_dm(d, ahi, bhi, desc_hi, idesc, 1u);
                    }
                    umma_commit(&b_empty[bs]);
                    umma_commit(&acc_full[set]);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// dof-major <-> natural rows (state_samples accessors, tests): one thread per element, coalesced on the output side
__global__ void traj_from_dof_major_kernel(const float* __restrict__ xdm, float* __restrict__ x, long long B, int H, int dof, int to_dm) {
    const int M = 2 * H * dof;
    const long long total = B * M;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / M;
        const int c = (int)(i - b * M);
        if (to_dm) {            // c is a dof-major column: j, n = 2 h + pv
            const int j = c / (2 * H), n = c - j * 2 * H;
            x[i] = xdm[b * M + (n >> 1) * 2 * dof + (n & 1) * dof + j];
        } else {                // c is a natural column: h, pv, j
            const int h = c / (2 * dof), r = c - h * 2 * dof, pv = r / dof, j = r - pv * dof;
            x[i] = xdm[b * M + j * 2 * H + 2 * h + pv];
        }
    }
}

template <int DOF>
static cudaError_t launch_kron_gen_dm(GenDmArgs& a, const NoiseArgs& noise, cudaStream_t st) {
    using C = GenDmCfg<DOF>;
    a.ntiles = (int)((a.Ntot + C::TS - 1) / C::TS);
    const long long U = (long long)DOF * a.ntiles;
    const int grid = U < sm_count() ? (int)U : sm_count();
    {
        const char* v = getenv("MPB_DM_VARIANT");
        // a CTA's range must hold at most two dofs (two factor buffers, loaded once): always true for grid = min(U, SMs >= 7)
        // default: 24 producer warps with the roles folded (sample_gp_kron_gen_dm3_kernel); MPB_DM_VARIANT=2: the kernel with
        // dedicated loader / MMA / mat-vec warps and the timing hooks (two producer groups)
        if (!(v && atoi(v) == 2) && (U + grid - 1) / grid <= (long long)a.ntiles) {
            using C3 = GenDm3Cfg<DOF>;
            auto k3 = sample_gp_kron_gen_dm3_kernel<DOF>;
            const cudaError_t e3 = cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C3::SMEM);
            if (e3 != cudaSuccess) return e3;
            return launch_pdl(k3, dim3(grid), dim3(C3::THREADS), C3::SMEM, st, a, noise);
        }
    }
    auto kern = sample_gp_kron_gen_dm_kernel<DOF>;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    return launch_pdl(kern, dim3(grid), dim3(C::THREADS), C::SMEM, st, a, noise);
}

}  // namespace mpb

extern "C" int mpb_sample_gp_kron_gen_dm(const void* Limg, const float* mu, const mpb_noise_desc* nd, float* x_dm, int P, int S, int H,
                                         int dof, const float* Sigma_inv, float* y, float* mu_copy, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(Limg && mu && nd && x_dm, "mpb_sample_gp_kron_gen_dm: null pointer");
    MPB_REQUIRE((Sigma_inv == nullptr) == (y == nullptr), "mpb_sample_gp_kron_gen_dm: Sigma_inv and y go together");
    MPB_REQUIRE(!mu_copy || y, "mpb_sample_gp_kron_gen_dm: mu_copy needs the mat-vec warp (Sigma_inv, y)");
    MPB_REQUIRE(P >= 0 && S >= 0, "mpb_sample_gp_kron_gen_dm: bad sizes P=%d S=%d", P, S);
    MPB_REQUIRE(mpb_sample_gp_kron_gen_supported(H, dof), "mpb_sample_gp_kron_gen_dm: shape H=%d dof=%d not supported", H, dof);
    MPB_REQUIRE(((uintptr_t)Limg | (uintptr_t)mu | (uintptr_t)x_dm) % 16 == 0, "mpb_sample_gp_kron_gen_dm: pointers must be 16-byte aligned");
    if (P == 0 || S == 0) return MPB_OK;
    GenDmArgs a{};
    NoiseArgs noise{};
    const char* why = noise_args(*nd, P, noise);
    MPB_REQUIRE(!why, "mpb_sample_gp_kron_gen_dm: %s", why);
    a.Limg = static_cast<const unsigned char*>(Limg);
    a.mu = mu; a.x = x_dm; a.P = P; a.S = S;
    a.Sinv = Sigma_inv; a.y = y; a.mu_copy = mu_copy;
    a.Ntot = (long long)P * S;
    { const char* v = getenv("MPB_KRON_DM_DBG"); a.dbg = v ? atoi(v) : 0; }
    { const char* v = getenv("MPB_KRON_DM_TRACE"); a.trace = v ? reinterpret_cast<long long*>(strtoull(v, nullptr, 0)) : nullptr; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaErrorInvalidValue;
    switch (dof) {
        case 2: e = launch_kron_gen_dm<2>(a, noise, st); break;
        case 3: e = launch_kron_gen_dm<3>(a, noise, st); break;
        case 4: e = launch_kron_gen_dm<4>(a, noise, st); break;
        case 5: e = launch_kron_gen_dm<5>(a, noise, st); break;
        case 6: e = launch_kron_gen_dm<6>(a, noise, st); break;
        case 7: e = launch_kron_gen_dm<7>(a, noise, st); break;
    }
    if (e != cudaSuccess) { set_error("mpb_sample_gp_kron_gen_dm: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    return check_launch("mpb_sample_gp_kron_gen_dm");
}

extern "C" int mpb_traj_from_dof_major(const float* x_dm, float* x, long long B, int H, int dof, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(x_dm && x && x_dm != x, "mpb_traj_from_dof_major: null or aliased pointers");
    MPB_REQUIRE(B >= 0 && H >= 1 && dof >= 1, "mpb_traj_from_dof_major: bad sizes");
    if (B == 0) return MPB_OK;
    const long long total = B * 2 * H * dof;
    const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    traj_from_dof_major_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x_dm, x, B, H, dof, 0);
    return check_launch("mpb_traj_from_dof_major");
}

extern "C" int mpb_traj_to_dof_major(const float* x, float* x_dm, long long B, int H, int dof, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(x_dm && x && x_dm != x, "mpb_traj_to_dof_major: null or aliased pointers");
    MPB_REQUIRE(B >= 0 && H >= 1 && dof >= 1, "mpb_traj_to_dof_major: bad sizes");
    if (B == 0) return MPB_OK;
    const long long total = B * 2 * H * dof;
    const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    traj_from_dof_major_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_dm, B, H, dof, 1);
    return check_launch("mpb_traj_to_dof_major");
}
