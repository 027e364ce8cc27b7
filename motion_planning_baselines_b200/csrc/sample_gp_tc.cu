// K1 (tensor-core variant): GP-prior sampling  x[p,s,:] = mu[p,:] + L @ eps[s,p,:]  on tcgen05.
//
// Replaces MultiMPPrior.sample (mp_baselines/planners/costs/factors/mp_priors_multi.py:253-256).  The shape
// is a real dense contraction -- X[N,M] = E[N,M] * L^T with N = P*S = 32768 rows and M = K = 896 at the
// BASELINE.json C4 shape -- so it runs on the 5th-generation tensor cores.  fp32 parity (1e-5) rules out
// plain TF32 (10-bit mantissa); we use the error-compensated 3xTF32 split
//        E*L^T  ~=  E_hi*L_hi^T + E_lo*L_hi^T + E_hi*L_lo^T        (fp32 accumulation in TMEM)
// whose dropped term E_lo*L_lo is O(2^-22) relative.
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the raw eps tile [128 x 16] and of the pre-split factor
//               tiles L_hi / L_lo [224 x 16] (64-byte swizzle) into a 6-stage shared-memory ring (36 KiB per stage)
//   warps 4-7   transform: each thread owns one eps row of the tile, splits it into hi (low 13 mantissa bits cleared)
//               and lo = e - hi and writes both with tcgen05.st into a 2-slot A-operand ring in TENSOR MEMORY
//   warp 1      MMA issuer: one elected thread issues 6 tcgen05.mma.kind::tf32 per stage (2 k-steps x 3 products) with
//               the A operand read from TMEM and B from shared memory; accumulators [128 x 224] fp32 in TMEM, double
//               buffered (columns 0-223 and 224-447; the A ring lives in columns 448-511)
//   warps 8-11  epilogue: tcgen05.ld the accumulator, add mu_p, store the row of x (particle-major)
// Why A sits in TMEM: with both operands in shared memory the kernel was bound by shared-memory bandwidth (every
// M128 x N256 x K8 tf32 MMA re-reads 4 KiB of A and 8 KiB of B; three products per k-step), ncu: tensor pipe 47 % busy.
// Keeping E_hi / E_lo in tensor memory removes the A reads and the transform's write-back from the shared-memory port.
// L is lower triangular, so k-blocks right of the diagonal block are never loaded or multiplied.
// Pipelines: smem full/empty mbarriers (TMA -> transform/MMA -> TMA), A-ring full/empty (transform <-> MMA) and TMEM
// accumulator full/empty (MMA <-> epilogue).
#include <cuda.h>

#include <cstdlib>

#include "mpb_common.cuh"
#include "tcgen05.cuh"

namespace mpb {

constexpr int TC_BM = 128;          // rows of eps per tile  (UMMA M)
constexpr int TC_BN = 224;          // columns of x per tile (UMMA N, rows of L): 896 = 4 x 224, 2 x 224 + 64 = 512 TMEM columns
constexpr int TC_BK = 16;           // k per stage: 16 fp32 = one 64-byte swizzle row
constexpr int TC_STAGES = 6;
constexpr int TC_ASLOTS = 2;        // A-operand slots in tensor memory (hi | lo, 2 x TC_BK columns each)
constexpr int TC_SWIZZLE_BYTES = TC_BK * 4;
constexpr int TC_THREADS = 384;
constexpr uint32_t A_TILE_BYTES = TC_BM * TC_BK * 4;      // 8 KiB   raw eps
constexpr uint32_t B_TILE_BYTES = TC_BN * TC_BK * 4;      // 14 KiB  per factor part
constexpr uint32_t STAGE_BYTES = A_TILE_BYTES + 2 * B_TILE_BYTES;         // eps | L_hi | L_lo  = 36 KiB
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr uint32_t TC_ACC_COLS = 2 * TC_BN;               // 448: two accumulators
constexpr uint32_t TC_A_COL0 = TC_ACC_COLS;               // A ring starts here (64 columns)
static_assert(TC_ACC_COLS + TC_ASLOTS * 2 * TC_BK <= 512, "tensor memory budget");
static_assert((2 * TC_STAGES + 2 * TC_ASLOTS + 4) * 8 + 4 <= 256, "barrier block too small");
static_assert(STAGE_BYTES % 1024 == 0 && A_TILE_BYTES % 1024 == 0 && B_TILE_BYTES % 512 == 0, "swizzle atom alignment");

// K-major operand tile with 128-byte (64-byte) swizzle: rows of 128 B (64 B), 8-row swizzle atoms of 1024 B (512 B)
// stacked along M/N.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8 * TC_SWIZZLE_BYTES) >> 4) << 32;   // stride byte offset: 8 rows x swizzle row
    d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
    d |= (uint64_t)(TC_SWIZZLE_BYTES == 128 ? 2 : 4) << 61;   // SWIZZLE_128B | SWIZZLE_64B
    return d;
}
// kind::tf32, fp32 accumulate, both operands K-major, M x N.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcArgs {
    const float* mu;
    float* x;
    int P, S, M, N;             // N = P*S rows
    int n_row_tiles, n_col_tiles;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
sample_gp_tc_kernel(const __grid_constant__ CUtensorMap map_eps, const __grid_constant__ CUtensorMap map_lhi,
                    const __grid_constant__ CUtensorMap map_llo, const TcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + TC_STAGES * STAGE_BYTES);
    uint64_t* full = bars;                                   // [STAGES] TMA landed (eps + L_hi + L_lo)
    uint64_t* empty = bars + TC_STAGES;                      // [STAGES] MMAs that read the stage completed
    uint64_t* a_full = bars + 2 * TC_STAGES;                 // [ASLOTS] E_hi | E_lo written to tensor memory
    uint64_t* a_empty = a_full + TC_ASLOTS;                  // [ASLOTS] MMAs that read the A slot completed
    uint64_t* tmem_full = a_empty + TC_ASLOTS;               // [2] accumulator ready
    uint64_t* tmem_empty = tmem_full + 2;                    // [2] accumulator drained
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < TC_ASLOTS; ++s) {
            mbar_init(&a_full[s], 4);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    const int n_tiles = a.n_row_tiles * a.n_col_tiles;
    // tile t -> (column tile jc, row tile r); column tiles are visited from the right (longest k range) first
    auto tile_coords = [&](int t, int& r, int& jc, int& n_cols, int& n_kb) {
        const int jj = t / a.n_row_tiles;
        r = t - jj * a.n_row_tiles;
        jc = a.n_col_tiles - 1 - jj;
        const int c0 = jc * TC_BN;
        n_cols = min(TC_BN, a.M - c0);
        const int k_end = min(a.M, c0 + TC_BN);          // L[i,k] = 0 for k > i
        n_kb = (k_end + TC_BK - 1) / TC_BK;
    };

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int r, jc, n_cols, n_kb;
                tile_coords(t, r, jc, n_cols, n_kb);
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* st = tiles + stage * STAGE_BYTES;
                    mbar_expect_tx(&full[stage], A_TILE_BYTES + 2 * B_TILE_BYTES);
                    tma_load_2d(&map_eps, &full[stage], st, kb * TC_BK, r * TC_BM);
                    tma_load_2d(&map_lhi, &full[stage], st + A_TILE_BYTES, kb * TC_BK, jc * TC_BN);
                    tma_load_2d(&map_llo, &full[stage], st + A_TILE_BYTES + B_TILE_BYTES, kb * TC_BK, jc * TC_BN);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            int stage = 0, slot = 0, acc = 0;
            uint32_t phase = 0, slot_phase = 0, acc_phase = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int r, jc, n_cols, n_kb;
                tile_coords(t, r, jc, n_cols, n_kb);
                const uint32_t idesc = make_idesc_tf32(TC_BM, (n_cols + 15) & ~15);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_BN);
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&full[stage], phase);                // L_hi / L_lo tiles landed
                    mbar_wait(&a_full[slot], slot_phase);          // E_hi / E_lo in tensor memory
                    tc_fence_after();
                    const uint32_t st = smem_u32(tiles + stage * STAGE_BYTES);
                    const uint64_t b_hi = make_sw128_desc(st + A_TILE_BYTES);
                    const uint64_t b_lo = make_sw128_desc(st + A_TILE_BYTES + B_TILE_BYTES);
                    const uint32_t a_hi = tmem_base + TC_A_COL0 + (uint32_t)(slot * 2 * TC_BK);
                    const uint32_t a_lo = a_hi + TC_BK;
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);      // +32 B per k-step inside the swizzle row
                        umma_tf32_ts(d_tmem, a_lo + 8 * k, b_hi + koff, idesc, (kb | k) ? 1u : 0u);   // small terms first
                        umma_tf32_ts(d_tmem, a_hi + 8 * k, b_lo + koff, idesc, 1u);
                        umma_tf32_ts(d_tmem, a_hi + 8 * k, b_hi + koff, idesc, 1u);
                    }
                    umma_commit(&empty[stage]);
                    umma_commit(&a_empty[slot]);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    if (++slot == TC_ASLOTS) { slot = 0; slot_phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================================ transform: eps row -> E_hi | E_lo in tensor memory =====================
        const int m = threadIdx.x - 128;             // row of the tile = TMEM lane (warp q may access lanes 32q..32q+31)
        const int q = warp & 3;
        // 64-byte swizzle of the TMA tile: 16-byte chunk c of row m sits at physical chunk c ^ ((m >> 1) & 3)
        const uint32_t row_off = (uint32_t)((m >> 3) * 512 + (m & 7) * 64);
        const int sw = (m >> 1) & 3;
        int stage = 0, slot = 0;
        uint32_t phase = 0, slot_phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int r, jc, n_cols, n_kb;
            tile_coords(t, r, jc, n_cols, n_kb);
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                const unsigned char* rowp = tiles + stage * STAGE_BYTES + row_off;
                float hi[16], lo[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ sw) << 4));
                    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float h = __uint_as_float(__float_as_uint(e[i]) & 0xffffe000u);
                        hi[4 * c + i] = h;
                        lo[4 * c + i] = e[i] - h;
                    }
                }
                mbar_wait(&a_empty[slot], slot_phase ^ 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + TC_A_COL0 + (uint32_t)(slot * 2 * TC_BK);
                tmem_st16(taddr, hi);
                tmem_st16(taddr + TC_BK, lo);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[slot]);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                if (++slot == TC_ASLOTS) { slot = 0; slot_phase ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // ================================ epilogue ====================================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int r, jc, n_cols, n_kb;
            tile_coords(t, r, jc, n_cols, n_kb);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int m = r * TC_BM + q * 32 + lane;            // row of eps: m = s*P + p
            const bool row_ok = m < a.N;
            const int s = row_ok ? m / a.P : 0, p = row_ok ? m - s * a.P : 0;
            const float* mrow = a.mu + (size_t)p * a.M + jc * TC_BN;
            float* xrow = a.x + ((size_t)p * a.S + s) * a.M + jc * TC_BN;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_BN);
            for (int c = 0; c < n_cols; c += 32) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)c, v);
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        if (c + i + 3 < n_cols) {
                            const float4 m4 = __ldg(reinterpret_cast<const float4*>(mrow + c + i));
                            float4 o;
                            o.x = m4.x + v[i]; o.y = m4.y + v[i + 1]; o.z = m4.z + v[i + 2]; o.w = m4.w + v[i + 3];
                            *reinterpret_cast<float4*>(xrow + c + i) = o;
                        } else {
                            for (int e = 0; e < 4; ++e)
                                if (c + i + e < n_cols) xrow[c + i + e] = __ldg(mrow + c + i + e) + v[i + e];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// L -> L_hi (low 13 mantissa bits cleared) and L_lo = tf32-truncated (L - L_hi); both exactly representable in TF32.
__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = src[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[i] = h;
    lo[i] = __uint_as_float(__float_as_uint(v - h) & 0xffffe000u);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 row-major [rows, cols] tensor, box [box_rows, 32 floats], 128-byte swizzle, zero fill out of bounds.
static bool make_map(CUtensorMap* map, const float* base, int rows, int cols, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, TC_SWIZZLE_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


// =====================================================================================================================
// Structured tcgen05 sampler: the factor decouples over the dofs (see sample_gp_kron.cu), so  x = mu + L @ eps  is `dof`
// independent [2H,2H] lower-triangular products.  Same warp-specialised pipeline as above; what changes:
//   * an output tile is 16 waypoints = 32 outputs of each dof = 32*DOF contiguous columns of x; its accumulator holds the
//     dofs side by side (TMEM column j*32 + i_local), one M128 x N32 MMA chain per dof.  ONE accumulator, and the other
//     512 - 32*DOF columns are a deep A ring (>= DOF slots): a slot feeds only six small MMAs, and tcgen05.st + wait::st
//     + the mbarrier handshake cost ~900 cycles per round trip, so the transform writes the slots of ALL dofs of a stage,
//     waits once, and releases them together (measured: per-dof waits made a stage cost 6-7k cycles whatever N was);
//   * a stage is one k-chunk of 8 waypoints = 16 k of every dof: DOF raw eps boxes [128 x 16] (the 16*DOF contiguous raw
//     columns) + the packed factor tiles L_hi / L_lo [32*DOF x 16] of (output tile, k-chunk); triangular: output tile `to`
//     needs the k-chunks 0 .. 2*to+1 only;
//   * the transform thread of a row reads its 16*DOF raw values once (conflict-free 128-bit loads), and per dof gathers
//     the 16 values (raw column DOF*kk + j), splits them into hi | lo and writes them into the TMEM A ring -- the stride-DOF
//     gather happens in registers, no shared-memory transpose;
//   * the epilogue interleaves the dofs back (4 outputs x DOF dofs = 4*DOF contiguous floats of the row of x), stages a
//     quarter of the columns of the warp's 32 rows in shared memory at a time and writes them out COALESCED with mu_p
//     added (a thread-per-row store touches 32 cache lines per instruction and made the epilogue the bottleneck).
template <int DOF>
struct KtCfg {
    static constexpr int OW = 32;                                // outputs of one dof per tile (UMMA N)
    static constexpr int BN = OW * DOF;                          // accumulator columns per output tile
    static constexpr uint32_t EPS_BYTES = DOF * A_TILE_BYTES;    // DOF boxes [128 x 16] fp32
    static constexpr uint32_t L_BYTES = BN * TC_BK * 4;          // one packed factor tile
    static constexpr uint32_t STAGE = EPS_BYTES + 2 * L_BYTES;
    static constexpr int STAGES = 2;
    static constexpr int ASLOTS = (512 - BN) / (2 * TC_BK) < 12 ? (512 - BN) / (2 * TC_BK) : 12;
    static constexpr uint32_t A_COL0 = BN;
    static constexpr int EPI_PASSES = 4, EPI_COLS = BN / EPI_PASSES;   // columns staged per pass (2 groups of 4 outputs x DOF)
    static constexpr int EPI_STRIDE = EPI_COLS + 4;              // floats per staged row (keeps 128-bit accesses conflict-free)
    static constexpr uint32_t EPI_WARP_BYTES = 32 * EPI_STRIDE * 4 + 32 * 16;     // staged rows + per-row (x offset, mu offset)
    static constexpr uint32_t BAR_BYTES = 512;
    static constexpr uint32_t SMEM = STAGES * STAGE + 4 * EPI_WARP_BYTES + 1024 + BAR_BYTES;
    static_assert(BN + ASLOTS * 2 * TC_BK <= 512 && ASLOTS >= DOF, "tensor memory budget");
    static_assert(EPI_COLS % (4 * DOF) == 0, "epilogue pass = whole groups of 4 outputs");
    static_assert(STAGE % 1024 == 0 && L_BYTES % 512 == 0, "swizzle atom alignment");
    static_assert((2 * STAGES + 2 * ASLOTS + 2) * 8 + 4 <= BAR_BYTES, "barrier block too small");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct KtArgs {
    const float* mu;
    float* x;
    int P, S, M, N;             // N = P*S rows
    int n_row_tiles, n_to, n_kc;   // output tiles (16 waypoints) and k-chunks (8 waypoints) per dof
    int dbg;                       // MPB_KRON_UMMA_DBG bit mask (timing experiments only): 1 no MMA, 2 no tcgen05.st, 4 no write-out
};

template <int DOF>
__global__ void __launch_bounds__(TC_THREADS, 1)
sample_gp_kron_umma_kernel(const __grid_constant__ CUtensorMap map_eps, const __grid_constant__ CUtensorMap map_lhi,
                           const __grid_constant__ CUtensorMap map_llo, const KtArgs a) {
    using Cfg = KtCfg<DOF>;
    constexpr int STAGES = Cfg::STAGES, ASLOTS = Cfg::ASLOTS, OW = Cfg::OW, BN = Cfg::BN;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* epi = tiles + STAGES * Cfg::STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi + 4 * Cfg::EPI_WARP_BYTES);
    uint64_t* full = bars;                                   // [STAGES] TMA landed (eps boxes + L_hi + L_lo)
    uint64_t* empty = bars + STAGES;                         // [STAGES] MMAs that read the stage completed
    uint64_t* a_full = bars + 2 * STAGES;                    // [ASLOTS] E_hi | E_lo of one dof written to tensor memory
    uint64_t* a_empty = a_full + ASLOTS;                     // [ASLOTS] MMAs that read the A slot completed
    uint64_t* tmem_full = a_empty + ASLOTS;                  // accumulator ready
    uint64_t* tmem_empty = tmem_full + 1;                    // accumulator drained
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < ASLOTS; ++s) {
            mbar_init(&a_full[s], 4);
            mbar_init(&a_empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    const int n_tiles = a.n_row_tiles * a.n_to;
    // tile t -> (output tile to, row tile r); output tiles are visited from the last (longest k range) first
    // tile t -> (row tile r, output tile to).  The output tiles of one row tile are adjacent tiles, i.e. they run on adjacent
    // CTAs at the same time and share their eps chunks through L2; the rotation by r gives every CTA all tile sizes.
    auto tile_coords = [&](int t, int& r, int& to, int& n_kc) {
        r = t / a.n_to;
        to = (t + r) % a.n_to;
        n_kc = 2 * to + 2;
    };

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int r, to, n_kc;
                tile_coords(t, r, to, n_kc);
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* st = tiles + stage * Cfg::STAGE;
                    mbar_expect_tx(&full[stage], Cfg::STAGE);
#pragma unroll
                    for (int b = 0; b < DOF; ++b)
                        tma_load_2d(&map_eps, &full[stage], st + b * A_TILE_BYTES, (kc * DOF + b) * TC_BK, r * TC_BM);
                    const int lrow = (to * a.n_kc + kc) * BN;
                    tma_load_2d(&map_lhi, &full[stage], st + Cfg::EPS_BYTES, 0, lrow);
                    tma_load_2d(&map_llo, &full[stage], st + Cfg::EPS_BYTES + Cfg::L_BYTES, 0, lrow);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            int stage = 0, slot = 0;
            uint32_t phase = 0, slot_phase = 0, acc_phase = 0;
            constexpr uint32_t idesc = make_idesc_tf32(TC_BM, OW);
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int r, to, n_kc;
                tile_coords(t, r, to, n_kc);
                mbar_wait(tmem_empty, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base;
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(&full[stage], phase);                // factor tiles landed
                    const uint32_t st = smem_u32(tiles + stage * Cfg::STAGE);
                    // The six MMAs of one dof accumulate into the same 32 columns, i.e. they form a dependent chain through the
                    // tensor pipe; consecutive MMAs therefore go to DIFFERENT dofs (independent accumulator columns).
                    uint32_t a_hi[DOF];
                    {
                        int sl = slot;
                        uint32_t ph = slot_phase;
#pragma unroll
                        for (int j = 0; j < DOF; ++j) {
                            mbar_wait(&a_full[sl], ph);            // E_hi / E_lo of dof j in tensor memory
                            a_hi[j] = tmem_base + Cfg::A_COL0 + (uint32_t)(sl * 2 * TC_BK);
                            if (++sl == ASLOTS) { sl = 0; ph ^= 1; }
                        }
                    }
                    tc_fence_after();
                    const uint64_t b_hi0 = make_sw128_desc(st + Cfg::EPS_BYTES);
                    const uint64_t b_lo0 = make_sw128_desc(st + Cfg::EPS_BYTES + Cfg::L_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);          // +32 B per k-step inside the swizzle row
#pragma unroll
                        for (int term = 0; term < 3; ++term) {                       // lo*hi, hi*lo, hi*hi (small terms first)
#pragma unroll
                            for (int j = 0; j < DOF; ++j) {
                                const uint64_t joff = (uint64_t)((j * OW * TC_SWIZZLE_BYTES) >> 4);
                                const uint32_t av = a_hi[j] + (term == 0 ? TC_BK : 0) + 8 * k;
                                const uint64_t bv = (term == 1 ? b_lo0 : b_hi0) + joff + koff;
                                if (!(a.dbg & 1)) umma_tf32_ts(d_tmem + (uint32_t)(j * OW), av, bv, idesc, (term == 0 && k == 0 && kc == 0) ? 0u : 1u);
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < DOF; ++j) {
                        umma_commit(&a_empty[slot]);
                        if (++slot == ASLOTS) { slot = 0; slot_phase ^= 1; }
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tmem_full);
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================================ transform: raw eps row -> per-dof E_hi | E_lo in tensor memory ==========
        const int m = threadIdx.x - 128;             // row of the tile = TMEM lane (warp q may access lanes 32q..32q+31)
        const int q = warp & 3;
        // 64-byte swizzle of the TMA tile: 16-byte chunk c of row m sits at physical chunk c ^ ((m >> 1) & 3)
        const uint32_t row_off = (uint32_t)((m >> 3) * 512 + (m & 7) * 64);
        const int sw = (m >> 1) & 3;
        int stage = 0, slot = 0;
        uint32_t phase = 0, slot_phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int r, to, n_kc;
            tile_coords(t, r, to, n_kc);
            for (int kc = 0; kc < n_kc; ++kc) {
                mbar_wait(&full[stage], phase);
                const unsigned char* rowp = tiles + stage * Cfg::STAGE + row_off;
                float raw[16 * DOF];                  // raw column c = DOF*kk + j  (kk = 2*waypoint + [pos|vel])
#pragma unroll
                for (int b = 0; b < DOF; ++b)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 v = *reinterpret_cast<const float4*>(rowp + b * A_TILE_BYTES + ((c ^ sw) << 4));
                        raw[16 * b + 4 * c + 0] = v.x; raw[16 * b + 4 * c + 1] = v.y;
                        raw[16 * b + 4 * c + 2] = v.z; raw[16 * b + 4 * c + 3] = v.w;
                    }
                // all DOF slots of this stage: wait for them, write them, ONE wait::st, release them together
                {
                    int sl = slot;
                    uint32_t ph = slot_phase;
#pragma unroll
                    for (int j = 0; j < DOF; ++j) {
                        mbar_wait(&a_empty[sl], ph ^ 1);
                        if (++sl == ASLOTS) { sl = 0; ph ^= 1; }
                    }
                }
                tc_fence_after();
                {
                    int sl = slot;
#pragma unroll
                    for (int j = 0; j < DOF; ++j) {
                        float hi[16], lo[16];
#pragma unroll
                        for (int kk = 0; kk < 16; ++kk) {
                            const float e = raw[DOF * kk + j];
                            const float h = __uint_as_float(__float_as_uint(e) & 0xffffe000u);
                            hi[kk] = h;
                            lo[kk] = e - h;
                        }
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::A_COL0 + (uint32_t)(sl * 2 * TC_BK);
                        if (!(a.dbg & 2)) {
                            tmem_st16(taddr, hi);
                            tmem_st16(taddr + TC_BK, lo);
                        }
                        if (++sl == ASLOTS) sl = 0;
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < DOF; ++j) {
                    if (lane == 0) mbar_arrive(&a_full[slot]);
                    if (++slot == ASLOTS) { slot = 0; slot_phase ^= 1; }
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // ================================ epilogue ====================================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        float* stg = reinterpret_cast<float*>(epi + q * Cfg::EPI_WARP_BYTES);                  // [32][EPI_STRIDE]
        long long* xoff = reinterpret_cast<long long*>(stg + 32 * Cfg::EPI_STRIDE);             // [32] float offset of the x row, -1 = none
        long long* moff = xoff + 32;                                                            // [32] float offset of the mu row
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int r, to, n_kc;
            tile_coords(t, r, to, n_kc);
            {
                const int m = r * TC_BM + q * 32 + lane;            // row of eps: m = s*P + p
                long long xo = -1, mo = 0;
                if (m < a.N) {
                    const int s = m / a.P, p = m - s * a.P;
                    xo = ((long long)p * a.S + s) * a.M + to * BN;
                    mo = (long long)p * a.M + to * BN;
                }
                xoff[lane] = xo;
                moff[lane] = mo;
            }
            mbar_wait(tmem_full, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int pass = 0; pass < Cfg::EPI_PASSES; ++pass) {
                constexpr int GPP = Cfg::EPI_COLS / (4 * DOF);      // groups of 4 outputs per pass
#pragma unroll
                for (int gl = 0; gl < GPP; ++gl) {                  // outputs i_local = 4*g4 .. 4*g4+3 of every dof
                    const int g4 = pass * GPP + gl;
                    uint32_t v[DOF][4];
#pragma unroll
                    for (int j = 0; j < DOF; ++j) tmem_ld4_nowait(taddr + (uint32_t)(j * OW + 4 * g4), v[j]);
                    tmem_wait_ld();
                    float o[4 * DOF];                                // x column DOF*i_local + j
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int j = 0; j < DOF; ++j) o[DOF * e + j] = __uint_as_float(v[j][e]);
#pragma unroll
                    for (int c = 0; c < DOF; ++c)
                        *reinterpret_cast<float4*>(stg + lane * Cfg::EPI_STRIDE + 4 * DOF * gl + 4 * c) =
                            make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                }
                if (pass == Cfg::EPI_PASSES - 1) {                   // accumulator drained: the next tile's MMAs may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty);
                    acc_phase ^= 1;
                } else {
                    __syncwarp();
                }
                // coalesced write-out of the warp's 32 rows x EPI_COLS columns (+ mu_p)
                constexpr int V4 = Cfg::EPI_COLS / 4;            // 32 * V4 float4 per pass = V4 per lane
                constexpr int UB = (V4 % 7 == 0) ? 7 : (V4 % 4 == 0 ? 4 : 2);   // loads in flight per batch
                static_assert(V4 % UB == 0, "write-out batch");
#pragma unroll 1
                for (int i0 = (a.dbg & 4) ? V4 : 0; i0 < V4; i0 += UB) {
                    float4 m4[UB];
                    long long xo[UB];
                    int so[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) {               // all global loads of the batch first: one latency per batch
                        const int idx = (i0 + u) * 32 + lane;
                        const int row = idx / V4, c4 = idx - row * V4;
                        xo[u] = xoff[row];
                        so[u] = row * Cfg::EPI_STRIDE + 4 * c4;
                        m4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (xo[u] >= 0) {
                            m4[u] = __ldg(reinterpret_cast<const float4*>(a.mu + moff[row] + pass * Cfg::EPI_COLS + 4 * c4));
                            xo[u] += pass * Cfg::EPI_COLS + 4 * c4;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        if (xo[u] >= 0) {
                            const float4 nz = *reinterpret_cast<const float4*>(stg + so[u]);
                            float4 w;
                            w.x = m4[u].x + nz.x; w.y = m4[u].y + nz.y; w.z = m4[u].z + nz.z; w.w = m4[u].w + nz.w;
                            *reinterpret_cast<float4*>(a.x + xo[u]) = w;
                        }
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Lp_hi / Lp_lo [n_to][n_kc][32*dof][16]: row j*32 + il, column kk  <-  LkT[j][16*kc + kk][32*to + il]  split like split_tf32_kernel
__global__ void kron_umma_pack_kernel(const float* __restrict__ LkT, float* __restrict__ hi, float* __restrict__ lo, int H, int dof) {
    const int N = 2 * H, n_to = N / 32, n_kc = N / 16;
    const long long total = (long long)n_to * n_kc * 32 * dof * 16;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long r = idx;
        const int kk = (int)(r & 15); r >>= 4;
        const int il = (int)(r & 31); r >>= 5;
        const int j = (int)(r % dof); r /= dof;
        const int kc = (int)(r % n_kc);
        const int to = (int)(r / n_kc);
        const float v = LkT[((size_t)j * N + 16 * kc + kk) * N + 32 * to + il];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        hi[idx] = h;
        lo[idx] = __uint_as_float(__float_as_uint(v - h) & 0xffffe000u);
    }
}

template <int DOF>
static int launch_kron_umma(const float* Lp_hi, const float* Lp_lo, const float* mu, const float* eps, float* x, int P, int S,
                            int H, cudaStream_t st) {
    using Cfg = KtCfg<DOF>;
    const int N = P * S, M = 2 * H * DOF, n_to = 2 * H / 32, n_kc = 2 * H / 16;
    CUtensorMap m_eps, m_hi, m_lo;
    if (!make_map(&m_eps, eps, N, M, TC_BM) || !make_map(&m_hi, Lp_hi, n_to * n_kc * Cfg::BN, TC_BK, Cfg::BN) ||
        !make_map(&m_lo, Lp_lo, n_to * n_kc * Cfg::BN, TC_BK, Cfg::BN)) {
        set_error("mpb_sample_gp_kron_umma: cuTensorMapEncodeTiled failed");
        return MPB_ECUDA;
    }
    KtArgs a;
    a.mu = mu; a.x = x; a.P = P; a.S = S; a.M = M; a.N = N;
    a.n_row_tiles = (N + TC_BM - 1) / TC_BM;
    a.n_to = n_to; a.n_kc = n_kc;
    { const char* v = getenv("MPB_KRON_UMMA_DBG"); a.dbg = v ? atoi(v) : 0; }
    cudaError_t e = cudaFuncSetAttribute(sample_gp_kron_umma_kernel<DOF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) { set_error("mpb_sample_gp_kron_umma: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    const int n_tiles = a.n_row_tiles * n_to;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    sample_gp_kron_umma_kernel<DOF><<<grid, TC_THREADS, Cfg::SMEM, st>>>(m_eps, m_hi, m_lo, a);
    return check_launch("mpb_sample_gp_kron_umma");
}

}  // namespace mpb

extern "C" int mpb_split_tf32(const float* src, float* hi, float* lo, long long n, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(src && hi && lo && n >= 0, "mpb_split_tf32: bad arguments");
    if (n == 0) return MPB_OK;
    split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, hi, lo, (size_t)n);
    return check_launch("mpb_split_tf32");
}

extern "C" int mpb_sample_gp_tc_supported(int P, int S, int M) {
    return (M % 16 == 0) && M >= 32 && (long long)P * S >= 1 && mpb::get_encode() != nullptr;
}

extern "C" int mpb_sample_gp_tc(const float* L_hi, const float* L_lo, const float* mu, const float* eps, float* x, int P,
                                int S, int M, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(L_hi && L_lo && mu && eps && x, "mpb_sample_gp_tc: null pointer");
    MPB_REQUIRE(P >= 0 && S >= 0 && M >= 32 && M % 16 == 0, "mpb_sample_gp_tc: M=%d must be a multiple of 16 and >= 32", M);
    if (P == 0 || S == 0) return MPB_OK;
    MPB_REQUIRE((reinterpret_cast<uintptr_t>(eps) & 15) == 0 && (reinterpret_cast<uintptr_t>(L_hi) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(L_lo) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(mu) & 15) == 0,
                "mpb_sample_gp_tc: pointers must be 16-byte aligned");
    const int N = P * S;
    CUtensorMap m_eps, m_hi, m_lo;
    if (!make_map(&m_eps, eps, N, M, TC_BM) || !make_map(&m_hi, L_hi, M, M, TC_BN) || !make_map(&m_lo, L_lo, M, M, TC_BN)) {
        set_error("mpb_sample_gp_tc: cuTensorMapEncodeTiled failed");
        return MPB_ECUDA;
    }
    TcArgs a;
    a.mu = mu; a.x = x; a.P = P; a.S = S; a.M = M; a.N = N;
    a.n_row_tiles = (N + TC_BM - 1) / TC_BM;
    a.n_col_tiles = (M + TC_BN - 1) / TC_BN;
    cudaError_t e = cudaFuncSetAttribute(sample_gp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("mpb_sample_gp_tc: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    const int n_tiles = a.n_row_tiles * a.n_col_tiles;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    sample_gp_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(m_eps, m_hi, m_lo, a);
    return check_launch("mpb_sample_gp_tc");
}

extern "C" int mpb_sample_gp_kron_umma_supported(int H, int dof) {
    return (dof == 2 || dof == 3 || dof == 7) && H >= 16 && H % 16 == 0 && H <= 256 && mpb::get_encode() != nullptr;
}

extern "C" long long mpb_sample_gp_kron_umma_floats(int H, int dof) {
    return 2LL * (2 * H / 32) * (2 * H / 16) * 32 * dof * 16;      // L_hi then L_lo
}

extern "C" int mpb_sample_gp_kron_umma_prepare(const float* LkT, float* Lp, int H, int dof, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(LkT && Lp, "mpb_sample_gp_kron_umma_prepare: null pointer");
    MPB_REQUIRE(mpb_sample_gp_kron_umma_supported(H, dof), "mpb_sample_gp_kron_umma_prepare: unsupported shape H=%d dof=%d", H, dof);
    const long long half = mpb_sample_gp_kron_umma_floats(H, dof) / 2;
    const int grid = (int)((half + 255) / 256 < 2048 ? (half + 255) / 256 : 2048);
    kron_umma_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(LkT, Lp, Lp + half, H, dof);
    return check_launch("mpb_sample_gp_kron_umma_prepare");
}

extern "C" int mpb_sample_gp_kron_umma(const float* Lp, const float* mu, const float* eps, float* x, int P, int S, int H,
                                       int dof, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(Lp && mu && eps && x, "mpb_sample_gp_kron_umma: null pointer");
    MPB_REQUIRE(P >= 0 && S >= 0, "mpb_sample_gp_kron_umma: bad sizes P=%d S=%d", P, S);
    MPB_REQUIRE(mpb_sample_gp_kron_umma_supported(H, dof), "mpb_sample_gp_kron_umma: unsupported shape H=%d dof=%d", H, dof);
    if (P == 0 || S == 0) return MPB_OK;
    MPB_REQUIRE((long long)P * S <= 0x7fffffffLL / 2, "mpb_sample_gp_kron_umma: P*S too large");
    MPB_REQUIRE(((uintptr_t)Lp | (uintptr_t)mu | (uintptr_t)eps | (uintptr_t)x) % 16 == 0, "mpb_sample_gp_kron_umma: pointers must be 16-byte aligned");
    const float* Lp_lo = Lp + mpb_sample_gp_kron_umma_floats(H, dof) / 2;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dof) {
        case 2: return launch_kron_umma<2>(Lp, Lp_lo, mu, eps, x, P, S, H, st);
        case 3: return launch_kron_umma<3>(Lp, Lp_lo, mu, eps, x, P, S, H, st);
        case 7: return launch_kron_umma<7>(Lp, Lp_lo, mu, eps, x, P, S, H, st);
    }
    return MPB_EINVAL;
}
