// K1 (Blackwell path of the structured sampler, noise drawn in the kernel):
//     x[p,s,:] = mu[p,:] + L @ eps[s,p,:],   eps ~ N(0, I) generated on the fly
//
// Replaces MultiMPPrior.sample (mp_baselines/planners/costs/factors/mp_priors_multi.py:253-256) including the noise draw
// torch does inside MultivariateNormal.rsample.  The factor decouples over the dofs (sample_gp_kron.cu), so per dof j
//     X_j^T [2H x samples] = L_j [2H x 2H] * E_j^T [2H x samples]
// which this kernel runs on tcgen05 with the FACTOR as the M = 128 operand and a tile of 64 samples as N: the seven
// accumulators of a tile (7 x 64 fp32 columns) fit tensor memory at once, so every noise value is generated exactly once
// and every factor chunk is streamed once per tile.  FP32 accuracy comes from a two-term fp16 split of both operands
// (kind::f16, three MMAs per k-step: hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM; the factor is scaled per dof by a
// power of two so that its fp16 parts stay normal) -- the same arithmetic as the warp-MMA kernel it supersedes.
//
// One persistent CTA per SM, warp-specialised (832 threads for the 7-dof arm):
//   warp 8       factor loader: one elected lane streams the 56 KiB factor chunk of each k-step (16 k x 7 dofs x hi/lo,
//                pre-arranged on the host side of the C ABI in the exact shared-memory image) with ONE bulk-async copy
//                (TMA engine, mbarrier complete_tx) into a 2-stage ring
//   warp 9       MMA issuer: one elected lane, 21 tcgen05.mma.kind::f16 (M128 x N64 x K16) per k-step
//   warps 0-7    epilogue (TMEM lane quadrant x half of the samples of a 4-sample batch): tcgen05.ld the accumulators
//                (thread = output row n of every dof), x = mu + acc / scale, dofs
//                re-interleaved into full trajectory rows in shared memory (conflict-free: lane stride 7 words), rows
//                leave through bulk-async stores (14 KiB each), double buffered, issued by warp 10 (hand-off by mbarrier)
//   warps 11-24  noise producers: Philox4x32-10 + Box-Muller, split into fp16 hi / lo and written with 8-byte stores
//                straight into the canonical K-major (no swizzle) operand tiles of a 3-stage ring; the lane mapping makes
//                every store bank-conflict free
//   warp 25      (optional) y_p = Sigma^-1 mu_p of the CTA's particles for the cost kernel's importance-sampling term
// Noise layout MPB_NOISE_SPMD: the virtual global tensor is [S_glob, P_glob, dof, 2H] (dof-major inside a row), so the
// four normals of one Philox call are four consecutive k of ONE dof -- one 8-byte store per operand part.  As with the
// other layouts a value depends only on (seed, offset, global sample, global particle, element): results do not depend
// on the sharding.  mpb_philox_normal(MPB_NOISE_SPMD) dumps the same numbers in the [S,P,M] order the injected-noise
// kernels read, which is how the tests replay a run through the oracle.
#include <cuda_fp16.h>

#include <cstdlib>

#include "mpb_common.cuh"
#include "philox.cuh"
#include "tcgen05.cuh"

namespace mpb {

// TS_ = 64: one accumulator set (7 x 64 = 448 of the 512 tensor-memory columns): the epilogue of a tile and the MMAs of
// the next are serial.  TS_ = 32: TWO sets of 7 x 32 columns, ping-pong: the epilogue of tile t drains one set while the
// MMAs of tile t + 1 fill the other, so noise production (the longest stage) runs without the epilogue gap; the price is
// that every factor chunk is streamed once per 32 samples instead of once per 64 (twice the L2 -> shared-memory traffic,
// twice as many, half as wide MMAs).  A noise stage holds 64 / TS_ k-chunks, so a producer thread always runs four Philox
// chains per stage.
template <int DOF, int TS_>
struct GenCfg {
    static constexpr int NOUT = 128;                         // 2H = rows of a per-dof block (UMMA M); H = 64
    static constexpr int TS = TS_;                           // samples per tile (UMMA N)
    static constexpr int NSETS = TS == 32 ? 2 : 1;           // accumulator sets in tensor memory
    static constexpr int KPS = 64 / TS;                      // k-chunks per noise stage
    static constexpr int KC = 16;                            // k per factor stage = one kind::f16 MMA
    static constexpr int NKC = NOUT / KC;
    static constexpr int M = NOUT * DOF;                     // floats per trajectory row
    static constexpr uint32_t B_TILE = TS * KC * 2;          // noise tile of one (chunk, dof, part): 2 / 1 KiB
    static constexpr uint32_t A_TILE = NOUT * KC * 2;        // factor tile of one (dof, part): 4 KiB
    static constexpr uint32_t B_STAGE = KPS * DOF * 2 * B_TILE;
    static constexpr uint32_t A_STAGE = DOF * 2 * A_TILE;
    static_assert(TS == 64 || TS == 32, "tile of 64 or 32 samples");
    static constexpr int B_STAGES = 3, A_STAGES = 2;
    static constexpr int OUT_ROWS = 4;                       // samples per staged output batch
    static constexpr uint32_t OUT_BUF = OUT_ROWS * M * 4;
    static constexpr int PROD_WARPS = 2 * DOF;
    static constexpr int EPI_WARPS = 8;                      // warps 0-7: TMEM lane quadrant x half of a batch's samples
    static constexpr int LOAD_WARP = 8, MMA_WARP = 9, STORE_WARP = 10, FIRST_PROD_WARP = 11;
    static constexpr int MV_WARP = FIRST_PROD_WARP + PROD_WARPS;      // Sigma^-1 mu of this CTA's particles (optional)
    static constexpr int THREADS = (MV_WARP + 1) * 32;
    static constexpr uint32_t OFF_A = 0;
    static constexpr uint32_t OFF_B = OFF_A + A_STAGES * A_STAGE;
    static constexpr uint32_t OFF_OUT = OFF_B + B_STAGES * B_STAGE;
    static constexpr uint32_t OFF_BAR = OFF_OUT + 2 * OUT_BUF;
    static constexpr uint32_t SMEM = OFF_BAR + 256 + 128 /*alignment slack*/;
    static constexpr uint32_t TMEM_COLS = NSETS * DOF * TS <= 256 ? 256 : 512;
    static_assert(NSETS * DOF * TS <= 512, "accumulators must fit tensor memory");
    static_assert(SMEM <= 227 * 1024, "shared-memory budget");
};

struct GenArgs {
    const unsigned char* Limg;      // [NKC][DOF][hi|lo][A_TILE bytes] + inverse scales (DOF floats) behind it
    const float* mu;                // [P, M]
    float* x;                       // [P, S, M]
    int P, S;
    long long Ntot;                 // P * S rows
    int ntiles;
    const float* Sinv;              // optional [M, M] prior precision with the per-dof structure (mpb_prior_dof_structured)
    float* y;                       // optional [P, M]: y[p] = Sinv @ mu[p], computed by warp MV_WARP while the tiles run
    float* mu_copy;                 // optional [P, M]: copy of mu written by the same warp (the planner's pre-update means)
    long long* trace;               // optional [64] clock stamps of CTA 0 (MPB_KRON_GEN_TRACE = device pointer; timing experiments)
    int dbg;                        // MPB_KRON_GEN_DBG bit mask (timing experiments only): 1 no Philox (zeros), 2 no MMAs,
                                    // 4 no output stores, 8 no factor loads, 128 no epilogue TMEM loads, 256 no epilogue smem writes,
                                    // 512 factor tiles staged in tensor memory by tcgen05.cp (+ 1024: only the hi tile) -- bit-identical, slower
};

// fp16 hi / lo parts of four floats -> two 8-byte words
__device__ __forceinline__ void split4_f16(const float4 v, uint2& hi, uint2& lo) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(__fsub_rn(v.x, f01.x), __fsub_rn(v.y, f01.y));
    const __half2 l23 = __floats2half2_rn(__fsub_rn(v.z, f23.x), __fsub_rn(v.w, f23.y));
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// tcgen05.mma.kind::f16 with the two shared-memory descriptors given as (low word, common high word)
__device__ __forceinline__ void umma_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
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

// the same MMA with the A operand read from tensor memory (128 lanes x 8 columns per [128 x 16] fp16 tile)
__device__ __forceinline__ void umma_f16_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 db;\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// shared memory -> tensor memory: one [128 rows x 32 bytes] tile (the canonical no-swizzle K-major image) into 8 columns
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint32_t s_lo, uint32_t desc_hi) {
    asm volatile(
        "{\n"
        ".reg .b64 ds;\n"
        "mov.b64 ds, {%1, %2};\n"
        "tcgen05.cp.cta_group::1.128x256b [%0], ds;\n"
        "}\n" ::"r"(taddr),
        "r"(s_lo), "r"(desc_hi)
        : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

template <int DOF, int TS_>
__global__ void __launch_bounds__(GenCfg<DOF, TS_>::THREADS, 1)
sample_gp_kron_gen_kernel(const GenArgs a, const NoiseArgs noise) {
    using C = GenCfg<DOF, TS_>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* a_full = bars;                           // [A_STAGES] factor chunk landed
    uint64_t* a_empty = a_full + C::A_STAGES;          // [A_STAGES] MMAs that read it completed
    uint64_t* b_full = a_empty + C::A_STAGES;          // [B_STAGES] noise chunk written (PROD_WARPS arrivals)
    uint64_t* b_empty = b_full + C::B_STAGES;          // [B_STAGES]
    uint64_t* acc_full = b_empty + C::B_STAGES;        // [NSETS] accumulators of the tile complete
    uint64_t* acc_empty = acc_full + C::NSETS;         // [NSETS] accumulators drained (epilogue warps)
    uint64_t* out_full = acc_empty + C::NSETS;         // [2] staged rows written (one arrival per epilogue warp)
    uint64_t* out_empty = out_full + 2;                // [2] the bulk store has read the buffer (store warp)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(out_empty + 2);

    pdl_trigger();                                     // the cost kernel may be scheduled as CTAs of this grid exit
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // slot = 8 * tile ordinal + event: 0 MMA first chunk ready, 1 MMA tile committed, 2 epilogue start, 3 epilogue end,
    // 4 producer first chunk written, 5 producer last chunk written, 6 MMA got the accumulators
    auto stamp = [&](int ordinal, int ev) {
        if (a.trace && blockIdx.x == 0 && ordinal < 8) a.trace[8 * ordinal + ev] = clock64();
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < C::B_STAGES; ++s) { mbar_init(&b_full[s], C::PROD_WARPS); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < C::NSETS; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], C::EPI_WARPS); }
        for (int s = 0; s < 2; ++s) { mbar_init(&out_full[s], C::EPI_WARPS); mbar_init(&out_empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == C::MMA_WARP) tmem_alloc(tmem_base_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    pdl_wait();                                        // the means come from the previous iteration's update kernel

    if (warp == C::LOAD_WARP) {
        // ================================ factor loader ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
                for (int kc = 0; kc < C::NKC; ++kc) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    if (a.dbg & 8) {
                        mbar_arrive(&a_full[stage]);
                    } else {
                        mbar_expect_tx(&a_full[stage], C::A_STAGE);
                        bulk_load(sm + C::OFF_A + stage * C::A_STAGE, a.Limg + (size_t)kc * C::A_STAGE, C::A_STAGE, &a_full[stage]);
                    }
                    if (++stage == C::A_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == C::MMA_WARP) {
        // ================================ MMA issuer ===================================
        {
            // the whole warp runs the loop (uniform control flow keeps the descriptors in uniform registers); one elected
            // lane issues
            const bool leader = elect_one();
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            const uint32_t idesc = make_idesc_f16(C::NOUT, C::TS);
            const uint64_t adesc0 = make_nosw_desc(smem_u32(sm + C::OFF_A), 128u, 256u);
            const uint64_t bdesc0 = make_nosw_desc(smem_u32(sm + C::OFF_B), 128u, 256u);
            const uint32_t adesc_lo = (uint32_t)adesc0, bdesc_lo = (uint32_t)bdesc0, desc_hi = (uint32_t)(adesc0 >> 32);
            int ord = 0;
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++ord) {
                const int set = ord % C::NSETS;
                mbar_wait(&acc_empty[set], (((uint32_t)(ord / C::NSETS)) & 1u) ^ 1u);
                tc_fence_after();
                if (lane == 0) stamp(ord, 6);
                for (int kc = 0; kc < C::NKC; ++kc) {
                    const int u = kc % C::KPS;                      // chunk inside the noise stage
                    mbar_wait(&a_full[as], aph);
                    if (u == 0) mbar_wait(&b_full[bs], bph);
                    tc_fence_after();
                    if (kc == 0 && lane == 0) stamp(ord, 0);
                    // descriptors differ only in the 14-bit start-address field: one 32-bit add each
                    const uint32_t alo_w = adesc_lo + (uint32_t)((as * C::A_STAGE) >> 4);
                    const uint32_t blo_w = bdesc_lo + (uint32_t)((bs * C::B_STAGE + u * (DOF * 2 * C::B_TILE)) >> 4);
                    if (leader && (a.dbg & 512)) {
                        // experiment: the factor tiles go shared -> tensor memory once (tcgen05.cp) and the three MMAs of a
                        // (dof, k-step) read them from there; cp and mma of one thread execute in issue order
#pragma unroll
                        for (int j = 0; j < DOF; ++j) {
                            const uint32_t ahi = alo_w + (uint32_t)(((2 * j) * C::A_TILE) >> 4), alo = ahi + (C::A_TILE >> 4);
                            const uint32_t bhi = blo_w + (uint32_t)(((2 * j) * C::B_TILE) >> 4), blo = bhi + (C::B_TILE >> 4);
                            const uint32_t d = tmem_base + (uint32_t)(set * (DOF * C::TS) + j * C::TS);
                            const uint32_t slot = tmem_base + (uint32_t)(C::NSETS * DOF * C::TS) + (uint32_t)(((kc * DOF + j) & 3) * 16);
                            tmem_cp_128x256b(slot, ahi, desc_hi);
                            if (a.dbg & 1024) {                     // only the hi tile (used twice) goes through tensor memory
                                if (kc == 0) umma_f16_w(d, alo, bhi, desc_hi, idesc, 0u);
                                else umma_f16_w(d, alo, bhi, desc_hi, idesc, 1u);
                            } else {
                                tmem_cp_128x256b(slot + 8, alo, desc_hi);
                                if (kc == 0) umma_f16_ts_w(d, slot + 8, bhi, desc_hi, idesc, 0u);
                                else umma_f16_ts_w(d, slot + 8, bhi, desc_hi, idesc, 1u);
                            }
                            umma_f16_ts_w(d, slot, blo, desc_hi, idesc, 1u);
                            umma_f16_ts_w(d, slot, bhi, desc_hi, idesc, 1u);
                        }
                    } else if (leader && !(a.dbg & 2)) {
#pragma unroll
                        for (int j = 0; j < DOF; ++j) {
                            const uint32_t ahi = alo_w + (uint32_t)(((2 * j) * C::A_TILE) >> 4), alo = ahi + (C::A_TILE >> 4);
                            const uint32_t bhi = blo_w + (uint32_t)(((2 * j) * C::B_TILE) >> 4), blo = bhi + (C::B_TILE >> 4);
                            const uint32_t d = tmem_base + (uint32_t)(set * (DOF * C::TS) + j * C::TS);
                            if (kc == 0) umma_f16_w(d, alo, bhi, desc_hi, idesc, 0u);          // small terms first
                            else umma_f16_w(d, alo, bhi, desc_hi, idesc, 1u);
                            umma_f16_w(d, ahi, blo, desc_hi, idesc, 1u);
                            umma_f16_w(d, ahi, bhi, desc_hi, idesc, 1u);
                        }
                    }
                    if (leader) {
                        umma_commit(&a_empty[as]);
                        if (u == C::KPS - 1) umma_commit(&b_empty[bs]);
                    }
                    __syncwarp();
                    if (++as == C::A_STAGES) { as = 0; aph ^= 1; }
                    if (u == C::KPS - 1 && ++bs == C::B_STAGES) { bs = 0; bph ^= 1; }
                }
                if (leader) umma_commit(&acc_full[set]);
                if (lane == 0) stamp(ord, 1);
            }
        }
    } else if (warp < C::EPI_WARPS) {
        // ================================ epilogue =====================================
        // warp (q4, half): TMEM lane quadrant q4, samples 2 half, 2 half + 1 of every 4-sample batch
        const int q4 = warp & 3, half = warp >> 2;
        const int n_out = 32 * q4 + lane;                       // TMEM lane = output row of every dof
        const int et = threadIdx.x;                             // 0..255
        const float* inv_scale_g = reinterpret_cast<const float*>(a.Limg + (size_t)C::NKC * C::A_STAGE);
        float inv_scale[DOF];
#pragma unroll
        for (int j = 0; j < DOF; ++j) inv_scale[j] = __ldg(inv_scale_g + j);
        const bool one_particle = (a.S % C::TS) == 0;
        uint32_t use = 0;                                       // batches staged so far (all tiles)
        int ord = 0;
        for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++ord) {
            const int set = ord % C::NSETS;
            const long long row0 = (long long)t * C::TS;
            float mrow[DOF];
            if (one_particle) {
                const float* mp = a.mu + (size_t)(row0 / a.S) * C::M + DOF * n_out;
#pragma unroll
                for (int j = 0; j < DOF; ++j) mrow[j] = __ldg(mp + j);
            }
            mbar_wait(&acc_full[set], ((uint32_t)(ord / C::NSETS)) & 1u);
            tc_fence_after();
            if (et == 0) stamp(ord, 2);
            for (int b = 0; b < C::TS / C::OUT_ROWS; ++b, ++use) {
                const int buf = use & 1;
                uint32_t r[DOF][2];
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(set * (DOF * C::TS) + b * C::OUT_ROWS + 2 * half);
                if (!(a.dbg & 128)) {
#pragma unroll
                    for (int j = 0; j < DOF; ++j) tmem_ld2_nowait(taddr + (uint32_t)(j * C::TS), r[j]);
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < DOF; ++j) r[j][0] = r[j][1] = 0u;
                }
                mbar_wait(&out_empty[buf], ((use >> 1) & 1) ^ 1);      // the store that last read this buffer has drained it
                float* ob = reinterpret_cast<float*>(sm + C::OFF_OUT + buf * C::OUT_BUF) + DOF * n_out;
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    const int sl = 2 * half + s2;
                    if (!one_particle) {
                        const long long row = row0 + b * C::OUT_ROWS + sl;
                        const long long p = (row < a.Ntot ? row : a.Ntot - 1) / a.S;
                        const float* mp = a.mu + (size_t)p * C::M + DOF * n_out;
#pragma unroll
                        for (int j = 0; j < DOF; ++j) mrow[j] = __ldg(mp + j);
                    }
                    if (a.dbg & 256) continue;
#pragma unroll
                    for (int j = 0; j < DOF; ++j) ob[sl * C::M + j] = fmaf(__uint_as_float(r[j][s2]), inv_scale[j], mrow[j]);
                }
                fence_async_proxy();
                __syncwarp();
                if (lane == 0) mbar_arrive(&out_full[buf]);     // one arrival per warp (arrivals on one barrier serialise);
                                                                // the store warp takes it from here
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[set]);
            if (et == 0) stamp(ord, 3);
        }
    } else if (warp == C::STORE_WARP) {
        // ================================ store issuer ==================================
        // one lane: one bulk-async store (14 KiB of finished rows) per staged batch, so that no epilogue thread ever
        // spends the ~200 cycles a bulk copy takes to issue; a buffer goes back once the store that read it has drained
        if (lane == 0) {
            uint32_t n = 0;
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
                const long long row0 = (long long)t * C::TS;
                for (int b = 0; b < C::TS / C::OUT_ROWS; ++b, ++n) {
                    const int buf = n & 1;
                    mbar_wait_spin(&out_full[buf], (n >> 1) & 1);
                    const long long first = row0 + b * C::OUT_ROWS;
                    long long rows = a.Ntot - first;
                    if (rows > C::OUT_ROWS) rows = C::OUT_ROWS;
                    if (rows > 0 && !(a.dbg & 4))
                        bulk_store(a.x + (size_t)first * C::M, sm + C::OFF_OUT + buf * C::OUT_BUF, (uint32_t)(rows * C::M * 4));
                    bulk_commit();
                    bulk_wait_read<0>();                        // the store has read its buffer: hand it back (the epilogue
                    mbar_arrive(&out_empty[buf]);               // fills the other buffer meanwhile)
                }
            }
            bulk_wait<0>();
        }
    } else if (warp >= C::FIRST_PROD_WARP && warp < C::MV_WARP) {
        // ================================ noise producers ==============================
        const int pw = warp - C::FIRST_PROD_WARP;
        const int j = pw % DOF, g2 = pw / DOF;                  // dof, one bit of the 8-sample group
        const int r8 = lane & 7, par = (lane >> 3) & 1, g1 = lane >> 4;
        int bs = 0;
        uint32_t bph = 0;
        for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
            // the thread's two work units of a noise stage: TS = 64 -- two samples (8-sample groups g and g + 4) of ONE
            // k-chunk; TS = 32 -- one sample of TWO consecutive k-chunks.  Either way four Philox calls per stage.
            long long grow[2];
            int sgrp[2];
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                sgrp[it] = g1 + 2 * g2 + (C::KPS == 1 ? 4 * it : 0);
                long long n = (long long)t * C::TS + sgrp[it] * 8 + r8;
                if (n >= a.Ntot) n = a.Ntot - 1;                  // rows past the end are never stored
                const long long p = n / a.S, s = n - p * a.S;
                grow[it] = (((noise.s_off + s) * noise.P_glob + noise.p_off + p) * DOF + j) * (C::NOUT / 4);
            }
            for (int ks = 0; ks < C::NKC / C::KPS; ++ks) {
                mbar_wait(&b_empty[bs], bph ^ 1);
                const uint32_t stage = smem_u32(sm + C::OFF_B + bs * C::B_STAGE + (2 * j) * C::B_TILE);
                // the four Philox / Box-Muller chains of this thread are generated together (no branches in between) so
                // that their long dependency chains interleave; rows past the end draw (unused) values like any other
                uint2 hi[4], lo[4];
                if (!(a.dbg & 1)) {
                    float4 e[4];
#pragma unroll
                    for (int it = 0; it < 2; ++it)
#pragma unroll
                        for (int qq = 0; qq < 2; ++qq) {
                            const int kc = C::KPS == 1 ? ks : 2 * ks + it;
                            e[2 * it + qq] = philox_normal4((unsigned long long)grow[it] + (unsigned)(kc * 4 + 2 * qq + par), noise);
                        }
#pragma unroll
                    for (int i = 0; i < 4; ++i) split4_f16(e[i], hi[i], lo[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) hi[i] = lo[i] = make_uint2(0u, 0u);
                }
#pragma unroll
                for (int it = 0; it < 2; ++it)
#pragma unroll
                    for (int qq = 0; qq < 2; ++qq) {
                        const int q = 2 * qq + par;                            // 4-k group inside the chunk
                        // (sample row, k) -> core matrix (row / 8, k / 8): 16-byte rows, K groups 128 B apart, row groups 256 B
                        const uint32_t off = stage + (uint32_t)((C::KPS == 1 ? 0 : it) * (DOF * 2 * C::B_TILE)) +
                                             (uint32_t)(sgrp[it] * 256 + (q >> 1) * 128 + r8 * 16 + (q & 1) * 8);
                        st_shared_v2(off, hi[2 * it + qq]);
                        st_shared_v2(off + C::B_TILE, lo[2 * it + qq]);
                    }
                fence_async_proxy();
                __syncwarp();
                if (lane == 0) mbar_arrive(&b_full[bs]);
                if (pw == 0 && lane == 0 && (ks == 0 || ks == C::NKC / C::KPS - 1)) stamp((t - blockIdx.x) / gridDim.x, ks == 0 ? 4 : 5);
                if (++bs == C::B_STAGES) { bs = 0; bph ^= 1; }
            }
        }
    } else if (warp == C::MV_WARP && a.y) {
        // ================================ Sigma^-1 mu ===================================
        // The importance-sampling term of the cost kernel needs y_p = Sigma^-1 mu_p (stoch_gpmp.py:239-241).  It depends
        // only on the means this kernel reads anyway, so one otherwise idle warp computes it for the particles
        // p = blockIdx.x, blockIdx.x + gridDim.x, ... while the tiles run: the 12 us launch + latency chain of a separate
        // prior_matvec_dof_kernel disappears from the iteration.  Same arithmetic as that kernel (sample_gp.cu: 7
        // non-zeros per row, double-float accumulation in the order m = -3..3), so y is bit-identical.
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
                        if (a.mu_copy) a.mu_copy[(size_t)(pb + q * gridDim.x) * C::M + i] = mv[q][3];      // mv[q][3] = mu[p][i]
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C::MMA_WARP) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// Factor image: per k-chunk kc, dof j, part (hi | lo): the [128 x 16] fp16 tile  A[n][k] = L_j[n][16 kc + k] * scale_j
// in the canonical K-major no-swizzle layout (8-row x 16-byte core matrices, K groups 128 B apart, row groups 256 B apart).
__global__ void kron_gen_image_kernel(const float* __restrict__ LkT, unsigned char* __restrict__ img, const unsigned* __restrict__ maxbits,
                                      int dof) {
    constexpr int N = 128, KC = 16, NKC = 8;
    const long long total = (long long)dof * N * N;
    float* inv_scale = reinterpret_cast<float*>(img + (size_t)NKC * dof * 2 * N * KC * 2);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(idx % N);
        const int k = (int)((idx / N) % N);
        const int j = (int)(idx / ((long long)N * N));
        const float mx = __uint_as_float(maxbits[j]);
        const float scale = mx > 0.f ? exp2f((float)(13 - ilogbf(mx))) : 1.f;
        if (n == 0 && k == 0) inv_scale[j] = 1.f / scale;
        const float v = LkT[((size_t)j * N + k) * N + n] * scale;               // LkT[j][k][n] = L_j[n][k]
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(__fsub_rn(v, __half2float(hi)));
        const int kc = k / KC, kk = k % KC;
        const size_t tile = ((size_t)(kc * dof + j) * 2) * (N * KC * 2);
        const size_t off = (size_t)(n >> 3) * 256 + (kk >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(img + tile + off) = hi;
        *reinterpret_cast<__half*>(img + tile + N * KC * 2 + off) = lo;
    }
}
__global__ void kron_gen_max_kernel(const float* __restrict__ LkT, unsigned* __restrict__ maxbits, int N, int dof) {
    const long long total = (long long)dof * N * N;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        atomicMax(maxbits + (int)(idx / ((long long)N * N)), __float_as_uint(fabsf(LkT[idx])));
}

}  // namespace mpb

// H = 64 (2H = 128 = the UMMA M) and 2..7 dofs: two accumulator sets of dof x 32 columns fit tensor memory up to 8 dofs, the
// shared-memory rings (32 KiB per dof) up to 7
extern "C" int mpb_sample_gp_kron_gen_supported(int H, int dof) { return (H == 64 && dof >= 2 && dof <= 7) ? 1 : 0; }

namespace mpb {
template <int DOF>
static cudaError_t launch_kron_gen(GenArgs& a, const NoiseArgs& noise, int ts, cudaStream_t st) {
    using C = GenCfg<DOF, 32>;
    a.ntiles = (int)((a.Ntot + ts - 1) / ts);
    const int grid = a.ntiles < sm_count() ? a.ntiles : sm_count();     // >= 1: the mat-vec warp strides the particles by the grid
    if (DOF == 7 && ts == 64) {         // the one-accumulator-set variant is kept for the 7-dof arm only (A/B timing)
        using C64 = GenCfg<7, 64>;
        auto kern = sample_gp_kron_gen_kernel<7, 64>;
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C64::SMEM);
        if (e != cudaSuccess) return e;
        return launch_pdl(kern, dim3(grid), dim3(C64::THREADS), C64::SMEM, st, a, noise);
    }
    auto kern = sample_gp_kron_gen_kernel<DOF, 32>;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    return launch_pdl(kern, dim3(grid), dim3(C::THREADS), C::SMEM, st, a, noise);
}
}  // namespace mpb

extern "C" long long mpb_sample_gp_kron_gen_bytes(int H, int dof) {
    return (long long)dof * 2 * (2 * H) * (2 * H) * 2 + 64 * 4;       // fp16 hi + lo image, inverse scales (16 words), scratch
}

extern "C" int mpb_sample_gp_kron_gen_prepare(const float* LkT, void* Limg, int H, int dof, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(LkT && Limg, "mpb_sample_gp_kron_gen_prepare: null pointer");
    MPB_REQUIRE(mpb_sample_gp_kron_gen_supported(H, dof), "mpb_sample_gp_kron_gen_prepare: shape H=%d dof=%d not supported", H, dof);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int N = 2 * H;
    const long long total = (long long)dof * N * N;
    unsigned char* img = static_cast<unsigned char*>(Limg);
    unsigned* tail = reinterpret_cast<unsigned*>(img + total * 4);
    cudaError_t e = cudaMemsetAsync(tail, 0, 64 * 4, st);
    if (e != cudaSuccess) { set_error("mpb_sample_gp_kron_gen_prepare: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    const int grid = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
    kron_gen_max_kernel<<<grid, 256, 0, st>>>(LkT, tail + 16, N, dof);
    kron_gen_image_kernel<<<grid, 256, 0, st>>>(LkT, img, tail + 16, dof);
    return check_launch("mpb_sample_gp_kron_gen_prepare");
}

extern "C" int mpb_sample_gp_kron_gen(const void* Limg, const float* mu, const mpb_noise_desc* nd, float* x, int P, int S, int H,
                                      int dof, void* stream) {
    return mpb_sample_gp_kron_gen_mv(Limg, mu, nd, x, P, S, H, dof, nullptr, nullptr, nullptr, stream);
}

extern "C" int mpb_sample_gp_kron_gen_mv(const void* Limg, const float* mu, const mpb_noise_desc* nd, float* x, int P, int S, int H,
                                         int dof, const float* Sigma_inv, float* y, float* mu_copy, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(Limg && mu && nd && x, "mpb_sample_gp_kron_gen: null pointer");
    MPB_REQUIRE((Sigma_inv == nullptr) == (y == nullptr), "mpb_sample_gp_kron_gen_mv: Sigma_inv and y go together");
    MPB_REQUIRE(!mu_copy || y, "mpb_sample_gp_kron_gen_mv: mu_copy needs the mat-vec warp (Sigma_inv, y)");
    MPB_REQUIRE(P >= 0 && S >= 0, "mpb_sample_gp_kron_gen: bad sizes P=%d S=%d", P, S);
    MPB_REQUIRE(mpb_sample_gp_kron_gen_supported(H, dof), "mpb_sample_gp_kron_gen: shape H=%d dof=%d not supported", H, dof);
    MPB_REQUIRE(((uintptr_t)Limg | (uintptr_t)mu | (uintptr_t)x) % 16 == 0, "mpb_sample_gp_kron_gen: pointers must be 16-byte aligned");
    if (P == 0 || S == 0) return MPB_OK;
    GenArgs a{};
    NoiseArgs noise{};
    const char* why = noise_args(*nd, P, noise);
    MPB_REQUIRE(!why, "mpb_sample_gp_kron_gen: %s", why);
    a.Limg = static_cast<const unsigned char*>(Limg);
    a.mu = mu; a.x = x; a.P = P; a.S = S;
    a.Sinv = Sigma_inv; a.y = y; a.mu_copy = mu_copy;
    a.Ntot = (long long)P * S;
    // MPB_KRON_GEN_TS=64: the one-accumulator-set variant (A/B timing, 7 dofs); default: 32-sample tiles, two sets
    int ts = 32;
    { const char* v = getenv("MPB_KRON_GEN_TS"); if (v && atoi(v) == 64 && dof == 7) ts = 64; }
    { const char* v = getenv("MPB_KRON_GEN_DBG"); a.dbg = v ? atoi(v) : 0; }
    { const char* v = getenv("MPB_KRON_GEN_TRACE"); a.trace = v ? reinterpret_cast<long long*>(strtoull(v, nullptr, 0)) : nullptr; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaErrorInvalidValue;
    switch (dof) {
        case 2: e = launch_kron_gen<2>(a, noise, ts, st); break;
        case 3: e = launch_kron_gen<3>(a, noise, ts, st); break;
        case 4: e = launch_kron_gen<4>(a, noise, ts, st); break;
        case 5: e = launch_kron_gen<5>(a, noise, ts, st); break;
        case 6: e = launch_kron_gen<6>(a, noise, ts, st); break;
        case 7: e = launch_kron_gen<7>(a, noise, ts, st); break;
    }
    if (e != cudaSuccess) { set_error("mpb_sample_gp_kron_gen: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    return check_launch("mpb_sample_gp_kron_gen");
}
