// K1s (structured FP32 variant): GP-prior sampling  x[p,s,:] = mu[p,:] + L @ eps[s,p,:]  for a factor that
// decouples over the degrees of freedom.
//
// Replaces MultiMPPrior.sample (mp_baselines/planners/costs/factors/mp_priors_multi.py:253-256).  The reference's
// prior precision is A^T Q^-1 A with K_s, K_g and Q_c proportional to the identity (unary_factor.py:19,
// gp_factor.py:23-26,42-50), so in the state ordering (t, [pos|vel], j) it only couples entries of the SAME dof j.
// Cholesky and the triangular solve (torch _precision_to_scale_tril) keep exact zeros exact, hence scale_tril has
// L[(t,a,j),(t',b,j')] == 0.0f whenever j != j': the [M,M] mat-vec is really `dof` independent [2H,2H] lower-triangular
// mat-vecs (7x fewer flops for the Panda; 2H = 128).  The per-dof blocks are NOT bit-identical to each other (fp32
// round-off of an ill-conditioned factorisation), so all `dof` blocks are kept.  mpb_sample_gp_kron_pack verifies the
// zero pattern bit-exactly and extracts the blocks; dropping exact zeros from an fp32 sum changes nothing, so the
// result equals the dense FP32 sum over the same ascending k order.
//
// Mapping: persistent CTAs, one tile = 32 sample rows x M columns staged in shared memory TRANSPOSED ([column][row],
// XOR-swizzled so that both the transposing stores and the 128-bit row reads are bank-conflict free).  One warp group
// per dof; warp a of a group owns the 16-row output blocks a and NB-1-a (the triangular work of the pair is the same
// for every a); lanes = 4 row quads x 8 sample quads, 4x4 register tile.  The factor block streams through a per-warp
// double-buffered cp.async ring (16 k x 16 rows per chunk), no CTA barrier inside the contraction.  Results go back
// into the tile in place; the store phase adds mu_p and writes full 32-byte sectors; the next tile's rows are already
// in flight in registers while the current tile is stored.  Measured bound: the shared-memory data pipe, not FMA issue --
// a 128-bit shared load is served per quarter-warp (4 wavefronts when the 8 lanes of a quarter read different addresses,
// 2 when they read one), i.e. 6 wavefronts per 16 FFMA: 0.210 ms at C4.  This kernel is the exact-FP32 parity anchor of
// the structured path (bit-identical to sample_gp_simt_kernel); the default is the tensor-core variant further down
// (sample_gp_kron_mma_kernel, 0.107 ms), and sample_gp_tc.cu holds a tcgen05 variant (sample_gp_kron_umma_kernel).
#include <cuda_fp16.h>

#include "mpb_common.cuh"
#include "philox.cuh"

namespace mpb {

constexpr int kTileRows = 32;

template <int DOF, int H>
struct KronCfg {
    static constexpr int D = 2 * DOF, M = H * D, N = 2 * H;       // N = rows/cols of one per-dof block
    static constexpr int NB = N / 16;                              // 16-row output blocks per dof
    static constexpr int WPD = NB / 2;                             // warps per dof
    static constexpr int WARPS = DOF * WPD, THREADS = WARPS * 32;
    static constexpr int TILE_FLOATS = kTileRows * M;
    static constexpr int LBUF_FLOATS = 2 * 16 * 16;                // per warp: two chunks of 16 k x 16 rows
    static constexpr int ITEMS = M / 4;                            // (half tile of 16 rows) x (8-column group)
    static constexpr int IPW = ITEMS / WARPS;                      // = 8 for every (DOF, H)
    static constexpr size_t SMEM = (size_t)(TILE_FLOATS + WARPS * LBUF_FLOATS) * sizeof(float);
    static_assert(H % 16 == 0, "H must be a multiple of 16");
    static_assert(ITEMS % WARPS == 0 && IPW == 8, "tile load mapping");
    static_assert(THREADS <= 1024, "too many warps");
};

__device__ __forceinline__ int tile_off(int c, int r) { return c * kTileRows + (r ^ ((c & 7) << 2)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

template <int DOF, int H>
__global__ void __launch_bounds__(KronCfg<DOF, H>::THREADS, 1)
sample_gp_kron_kernel(const float* __restrict__ LkT, const float* __restrict__ mu, const float* __restrict__ eps,
                      float* __restrict__ x, int P, int S) {
    using Cfg = KronCfg<DOF, H>;
    constexpr int D = Cfg::D, M = Cfg::M, N = Cfg::N, NB = Cfg::NB, WPD = Cfg::WPD, WARPS = Cfg::WARPS;
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* lbuf = smem + Cfg::TILE_FLOATS + warp * Cfg::LBUF_FLOATS;

    const long long Ntot = (long long)P * S;
    const int ntiles = (int)((Ntot + kTileRows - 1) / kTileRows);

    // contraction role
    const int j = warp / WPD, a = warp - j * WPD;
    const int ig = lane >> 3, sg = lane & 7;
    const float* Lj = LkT + (size_t)j * N * N;
    int xo[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) xo[m] = (4 * sg) ^ (((m + j) & 7) << 2);

    // tile load / store role: lane -> (row within a 16-row half, 4-column quad of an 8-column group)
    const int rl = lane >> 1, cl = lane & 1;

    float4 pre[Cfg::IPW];
    auto gload = [&](int t) {
        const long long n0 = (long long)t * kTileRows;
        const float* src[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const long long n = n0 + hf * 16 + rl;
            if (n < Ntot) {
                const int p = (int)(n / S), s = (int)(n - (long long)p * S);
                src[hf] = eps + ((size_t)s * P + p) * M;
            } else {
                src[hf] = nullptr;
            }
        }
#pragma unroll
        for (int it = 0; it < Cfg::IPW; ++it) {
            const int item = warp + it * WARPS;
            const int hf = item & 1, c0 = (item >> 1) * 8 + cl * 4;
            const float* sp = hf ? src[1] : src[0];
            pre[it] = sp ? __ldg(reinterpret_cast<const float4*>(sp + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int it = 0; it < Cfg::IPW; ++it) {
            const int item = warp + it * WARPS;
            const int hf = item & 1, c0 = (item >> 1) * 8 + cl * 4, r = hf * 16 + rl;
            tile[(c0 + 0) * kTileRows + (r ^ ((cl * 4 + 0) << 2))] = pre[it].x;
            tile[(c0 + 1) * kTileRows + (r ^ ((cl * 4 + 1) << 2))] = pre[it].y;
            tile[(c0 + 2) * kTileRows + (r ^ ((cl * 4 + 2) << 2))] = pre[it].z;
            tile[(c0 + 3) * kTileRows + (r ^ ((cl * 4 + 3) << 2))] = pre[it].w;
        }
    };

    // out[16*blk + 4*ig + ii][4*sg + ss] = sum_{k < 16*(blk+1)} LkT[k][16*blk + 4*ig + ii] * e[k][4*sg + ss]
    auto contract = [&](int blk, float (&acc)[4][4]) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) acc[ii][ss] = 0.f;
        const int nchunks = blk + 1;
        const float* lsrc = Lj + (size_t)(lane >> 2) * N + 16 * blk + 4 * (lane & 3);     // row kk = lane>>2 (+8)
        float* ldst = lbuf + (lane >> 2) * 16 + 4 * (lane & 3);
        auto issue = [&](int q) {
            float* dst = ldst + (q & 1) * 256;
            const float* s0 = lsrc + (size_t)q * 16 * N;
            cp_async16(dst, s0);
            cp_async16(dst + 8 * 16, s0 + (size_t)8 * N);
            cp_async_commit();
        };
        issue(0);
        for (int q = 0; q < nchunks; ++q) {
            cp_async_wait_all();
            __syncwarp();
            if (q + 1 < nchunks) issue(q + 1);
            const float* lb = lbuf + (q & 1) * 256 + 4 * ig;
            const float* eb = tile + (size_t)(q * 8 * D + j) * kTileRows;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                const int cK = (kk >> 1) * D + (kk & 1) * DOF;          // column of k within the chunk, minus j
                const float4 e = *reinterpret_cast<const float4*>(eb + cK * kTileRows + xo[cK & 7]);
                const float4 l = *reinterpret_cast<const float4*>(lb + kk * 16);
                const float lv[4] = {l.x, l.y, l.z, l.w}, ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                    for (int ss = 0; ss < 4; ++ss) acc[ii][ss] = fmaf(lv[ii], ev[ss], acc[ii][ss]);
            }
        }
        __syncwarp();
    };
    auto put = [&](int blk, const float (&acc)[4][4]) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int i = 16 * blk + 4 * ig + ii;
            const int c = (i >> 1) * D + (i & 1) * DOF + j;
            *reinterpret_cast<float4*>(tile + c * kTileRows + ((4 * sg) ^ ((c & 7) << 2))) =
                make_float4(acc[ii][0], acc[ii][1], acc[ii][2], acc[ii][3]);
        }
    };

    int t = blockIdx.x;
    if (t < ntiles) gload(t);
    for (; t < ntiles; t += gridDim.x) {
        sstore();
        __syncthreads();
        float acc0[4][4], acc1[4][4];
        contract(a, acc0);
        contract(NB - 1 - a, acc1);
        __syncthreads();                       // every read of the noise tile is done: overwrite it in place
        put(a, acc0);
        put(NB - 1 - a, acc1);
        __syncthreads();
        const int tn = t + gridDim.x;
        if (tn < ntiles) gload(tn);            // next tile's rows fly while this one is stored
        const long long n0 = (long long)t * kTileRows;
        const float* mrow[2];
        float* xrow[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const long long n = n0 + hf * 16 + rl;
            if (n < Ntot) {
                mrow[hf] = mu + (size_t)(n / S) * M;
                xrow[hf] = x + (size_t)n * M;
            } else {
                mrow[hf] = nullptr;
                xrow[hf] = nullptr;
            }
        }
#pragma unroll
        for (int it = 0; it < Cfg::IPW; ++it) {
            const int item = warp + it * WARPS;
            const int hf = item & 1, c0 = (item >> 1) * 8 + cl * 4, r = hf * 16 + rl;
            float* xr = hf ? xrow[1] : xrow[0];
            const float* mr = hf ? mrow[1] : mrow[0];
            if (xr) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(mr + c0));
                float4 o;
                o.x = m4.x + tile[(c0 + 0) * kTileRows + (r ^ ((cl * 4 + 0) << 2))];
                o.y = m4.y + tile[(c0 + 1) * kTileRows + (r ^ ((cl * 4 + 1) << 2))];
                o.z = m4.z + tile[(c0 + 2) * kTileRows + (r ^ ((cl * 4 + 2) << 2))];
                o.w = m4.w + tile[(c0 + 3) * kTileRows + (r ^ ((cl * 4 + 3) << 2))];
                *reinterpret_cast<float4*>(xr + c0) = o;
            }
        }
        __syncthreads();
    }
}

// ---- tensor-core variant (warp-level MMA, two-term fp16 split) ---------------------------------------------------------
// Same work split (one warp group per dof, warp a owns the 16-row blocks a and NB-1-a, per-warp cp.async ring for the
// factor block); the 16-row x 32-sample output block of a warp is four m16n8k16 MMAs per 16 k.  Both operands are split
// into two fp16 terms, v = hi + lo with hi = fp16(v), lo = fp16(v - hi) (22 significant bits); lo*hi + hi*lo + hi*hi
// accumulate in fp32 (the dropped lo*lo term is 2^-22 relative).  The factor block is scaled per dof by a power of two
// so that its largest entry sits at 2^13 (entries below 2^-37 of the largest flush to zero); noise is O(1) by contract
// (|eps| < 65504; entries below 2^-14 keep an absolute precision of 2^-25).  An fp16 MMA moves twice the k of a TF32
// MMA per issue, so this needs half the tensor-pipe time of a 3xTF32 split for the same accuracy.
// Fragments are gathered by the threads themselves, so the per-dof column stride of the noise rows costs nothing --
// which is why this kernel uses mma.sync and not tcgen05: a tcgen05 formulation needs the factor block resident next
// to a >= 128-sample tile in canonical core-matrix layout, and a 32-row tile of all dofs already fills shared memory.
// The tile keeps the rows in their NATURAL layout [32 rows][M + 4]: rows are copied global->shared with coalesced
// 16-byte cp.async (no registers) and converted in place to packed (hi | lo << 16) words by the thread that copied
// them; the padded stride (= 4 mod 32) makes the B-fragment gathers and the accumulator scatter bank-conflict free
// (the MMA's k slots are permuted so that the four lanes of a fragment row read consecutive k); the store phase is a
// coalesced row copy that adds mu_p and immediately refills the slot it has just read with the next tile's noise, so
// the next tile streams in while this one is written out.
template <int DOF, int H>
struct KronMmaCfg {
    static constexpr int D = 2 * DOF, M = H * D, N = 2 * H;       // N = rows/cols of one per-dof block
    static constexpr int NB = N / 16;                              // 16-row output blocks per dof
    static constexpr int ROWS = 32, NT = ROWS / 8;                 // sample rows per tile, m16n8 n-tiles per warp
    static constexpr int PPW = 1;                                  // block pairs (a, NB-1-a) per warp
    static constexpr int STAGES = 3;                               // factor-chunk ring: two chunks in flight
    static constexpr int WPD = NB / 2 / PPW;                       // warps per dof
    static constexpr int WARPS = DOF * WPD, THREADS = WARPS * 32;
    // resident CTAs the register file allows at 72 registers per thread (1 for the Panda's 896 threads, 2-4 for small robots)
    static constexpr int CTAS_PER_SM = 65536 / (72 * THREADS) < 1 ? 1 : (65536 / (72 * THREADS) > 4 ? 4 : 65536 / (72 * THREADS));
    static constexpr int RS = M + 4;                               // padded row stride in floats
    static constexpr int TILE_FLOATS = ROWS * RS;
    static constexpr int V4_PER_ROW = M / 4;
    static constexpr int IPT = ROWS * V4_PER_ROW / THREADS;        // float4 slots per thread = 8
    static constexpr int LBUF_FLOATS = STAGES * 256;               // per warp: STAGES x (hi + lo A fragments of one chunk)
    static constexpr size_t SMEM = (size_t)(TILE_FLOATS + WARPS * LBUF_FLOATS) * sizeof(float) + 2 * ROWS * 16;
    static_assert(H % 16 == 0 && (NB / 2) % PPW == 0, "H must be a multiple of 16 * PPW");
    static_assert(RS % 32 == 4, "row stride must be 4 mod 32");
    static_assert(ROWS * V4_PER_ROW % THREADS == 0 && IPT == 8, "tile copy mapping");
};

__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// v -> (fp16(v) | fp16(v - fp16(v)) << 16)
__device__ __forceinline__ uint32_t split_f16(float v) {
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(__fsub_rn(v, __half2float(hi)));
    return (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
}

struct KronRow {            // per tile row: where its noise comes from and which particle it belongs to
    long long src;          // float offset of the eps row (GEN: of the row in the virtual GLOBAL noise tensor), or -1 past the end
    int p, pad;
};

// (v0..v3) -> packed (fp16 hi | fp16 lo << 16) words, the tile's operand format
__device__ __forceinline__ uint4 pack_split4(const float4 v) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(__fsub_rn(v.x, f01.x), __fsub_rn(v.y, f01.y));
    const __half2 l23 = __floats2half2_rn(__fsub_rn(v.z, f23.x), __fsub_rn(v.w, f23.y));
    const uint32_t uh01 = *reinterpret_cast<const uint32_t*>(&h01), ul01 = *reinterpret_cast<const uint32_t*>(&l01);
    const uint32_t uh23 = *reinterpret_cast<const uint32_t*>(&h23), ul23 = *reinterpret_cast<const uint32_t*>(&l23);
    uint4 w;
    w.x = __byte_perm(uh01, ul01, 0x5410); w.y = __byte_perm(uh01, ul01, 0x7632);
    w.z = __byte_perm(uh23, ul23, 0x5410); w.w = __byte_perm(uh23, ul23, 0x7632);
    return w;
}

// GEN = false: noise rows are read from `eps` [S,P,M] (injected noise: parity runs).
// GEN = true : `eps` is unused; every 4-column slot is drawn in place with Philox4x32-10 keyed on the slot's index in the
//              virtual global noise tensor [S_glob, P_glob, M] (philox.cuh) and written straight in the packed operand
//              format, so the global read of the noise, the separate conversion pass and the torch.randn launch that
//              produced it all disappear.  Same bits as mpb_philox_normal(MPB_NOISE_SPM) followed by GEN = false.
template <int DOF, int H, bool GEN>
__global__ void __launch_bounds__(KronMmaCfg<DOF, H>::THREADS, KronMmaCfg<DOF, H>::CTAS_PER_SM)
sample_gp_kron_mma_kernel(const uint32_t* __restrict__ LkF, const float* __restrict__ mu, const float* __restrict__ eps,
                          float* __restrict__ x, int P, int S, const NoiseArgs noise) {
    using Cfg = KronMmaCfg<DOF, H>;
    constexpr int M = Cfg::M, N = Cfg::N, NB = Cfg::NB, WPD = Cfg::WPD, RS = Cfg::RS, THREADS = Cfg::THREADS;
    constexpr int ROWS = Cfg::ROWS, NT = Cfg::NT, PPW = Cfg::PPW;
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* lbuf = smem + Cfg::TILE_FLOATS + warp * Cfg::LBUF_FLOATS;
    KronRow* rows = reinterpret_cast<KronRow*>(smem + Cfg::TILE_FLOATS + Cfg::WARPS * Cfg::LBUF_FLOATS);   // [2][ROWS]

    const long long Ntot = (long long)P * S;
    const int ntiles = (int)((Ntot + ROWS - 1) / ROWS);

    const int j = warp / WPD, a0 = (warp - j * WPD) * PPW;       // this warp owns the block pairs a0 .. a0+PPW-1 of dof j
    const int g = lane >> 2, t4 = lane & 3;
    const uint32_t* Lj = LkF + (size_t)j * N * N;                      // [blk][q][hi|lo][lane][4]
    const float inv_scale = reinterpret_cast<const float*>(LkF + (size_t)DOF * N * N)[j];

    auto row_table = [&](int t, int buf) {          // threads 0..ROWS-1
        const long long n = (long long)t * ROWS + threadIdx.x;
        KronRow r;
        r.src = -1; r.p = 0; r.pad = 0;
        if (n < Ntot) {
            const int p = (int)(n / S), s = (int)(n - (long long)p * S);
            r.src = GEN ? ((noise.s_off + s) * noise.P_glob + noise.p_off + p) * M : ((long long)s * P + p) * M;
            r.p = p;
        }
        rows[buf * ROWS + threadIdx.x] = r;
    };
    auto fill_slot = [&](int buf, int r, int v) {   // float4 slot (row r, 4-column group v) <- next tile's noise
        const long long src = rows[buf * ROWS + r].src;
        float* dst = tile + r * RS + 4 * v;
        if (GEN) {
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            if (src >= 0) w = pack_split4(philox_normal4((unsigned long long)(src >> 2) + (unsigned)v, noise));
            *reinterpret_cast<uint4*>(dst) = w;
        } else {
            if (src >= 0) cp_async16_cg(dst, eps + src + 4 * v);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    // Factor staging: every lane copies exactly the 2 x 16 bytes (hi and lo A fragments of one 16-row x 16-k chunk) it
    // reads back itself, so the ring needs no warp synchronisation.  All chunks of a warp form ONE sequence (per pair:
    // block a with q = 0..a, then block NB-1-a with q = 0..NB-1-a -- NB+1 chunks whatever a is); chunk u+1 is in flight
    // while chunk u runs, across block, pair and tile boundaries.
    float* ldst = lbuf + lane * 4;
    constexpr int SEQ = PPW * (NB + 1);                // chunks per warp and tile
    auto issue = [&](int v, int stage) {               // chunk v of the sequence (v >= SEQ: nothing, but keep the group count)
        if (v < SEQ) {
            const int i = v / (NB + 1), r = v - i * (NB + 1), a = a0 + i;
            const int blk = r <= a ? a : NB - 1 - a, q = r <= a ? r : r - a - 1;
            const uint32_t* s0 = Lj + ((size_t)blk * NB + q) * 256 + lane * 4;
            float* dst = ldst + stage * 256;
            cp_async16(dst, s0);
            cp_async16(dst + 128, s0 + 128);
        }
        cp_async_commit();
    };
    // B fragment: element (k, sample) = tile[sample * RS + DOF*k + j]; MMA k slots (2*t4, 2*t4+1, 2*t4+8, 2*t4+9) hold
    // k = 16q + t4 + (0, 4, 8, 12) -- the same permutation is baked into the packed A fragments
    const uint32_t* ebase = reinterpret_cast<const uint32_t*>(tile) + g * RS + DOF * t4 + j;
    // acc[n][.] : m16n8 accumulator of n-tile n (samples 8n..8n+7) of the current 16-row block
    auto step = [&](int q, int stage, float (&acc)[NT][4]) {
        const uint4 ahi = *reinterpret_cast<const uint4*>(ldst + stage * 256);
        const uint4 alo = *reinterpret_cast<const uint4*>(ldst + stage * 256 + 128);
        const uint32_t* eb = ebase + q * 16 * DOF;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const uint32_t* e0 = eb + n * 8 * RS;
            const uint32_t w0 = e0[0], w1 = e0[4 * DOF], w2 = e0[8 * DOF], w3 = e0[12 * DOF];
            const uint32_t bhi0 = __byte_perm(w0, w1, 0x5410), blo0 = __byte_perm(w0, w1, 0x7632);
            const uint32_t bhi1 = __byte_perm(w2, w3, 0x5410), blo1 = __byte_perm(w2, w3, 0x7632);
            mma_f16(acc[n], alo, bhi0, bhi1);
            mma_f16(acc[n], ahi, blo0, blo1);
            mma_f16(acc[n], ahi, bhi0, bhi1);
        }
    };
    // x = mu_p + noise: the mean of the particle every row of the tile belongs to (one particle per tile when S % ROWS == 0)
    const bool one_particle = (S % ROWS) == 0;
    auto put = [&](int blk, const float (&acc)[NT][4], int buf, const float (&mrow)[2]) {
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int c = DOF * (16 * blk + g + 8 * hrow) + j;                       // == (i>>1)*D + (i&1)*DOF + j
            float* col = tile + c + 2 * t4 * RS;
            if (one_particle) {
                const float m = mrow[hrow];
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    col[(8 * n) * RS] = m + acc[n][2 * hrow] * inv_scale;
                    col[(8 * n + 1) * RS] = m + acc[n][2 * hrow + 1] * inv_scale;
                }
            } else {
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int p = rows[buf * ROWS + 8 * n + 2 * t4 + e].p;
                        col[(8 * n + e) * RS] = __ldg(mu + (size_t)p * M + c) + acc[n][2 * hrow + e] * inv_scale;
                    }
            }
        }
    };

    int t = blockIdx.x;
    int buf = 0;
    unsigned u = 0;                                    // running chunk counter (selects the ring stage)
    if (t < ntiles) {
        if (threadIdx.x < ROWS) row_table(t, 0);
        __syncthreads();
#pragma unroll
        for (int it = 0; it < Cfg::IPT; ++it) {
            const int e = threadIdx.x + it * THREADS;
            fill_slot(0, e / Cfg::V4_PER_ROW, e % Cfg::V4_PER_ROW);
        }
        cp_async_commit();
    }
    for (; t < ntiles; t += gridDim.x, buf ^= 1) {
        const int tn = t + gridDim.x;
        cp_async_wait_all();
        issue(0, u % Cfg::STAGES);                     // the first two factor chunks fly during the conversion pass
        issue(1, (u + 1) % Cfg::STAGES);
        float mrow[PPW][2][2];                         // means of this thread's output rows (one particle per tile)
#pragma unroll
        for (int i = 0; i < PPW; ++i)
#pragma unroll
            for (int hb = 0; hb < 2; ++hb)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int blk = hb ? NB - 1 - (a0 + i) : a0 + i;
                    mrow[i][hb][hrow] = one_particle ? __ldg(mu + (size_t)(((long long)t * ROWS) / S) * M + DOF * (16 * blk + g + 8 * hrow) + j) : 0.f;
                }
        if (!GEN) {
#pragma unroll
            for (int it = 0; it < Cfg::IPT; ++it) {    // own slots only: fp32 -> packed (fp16 hi | fp16 lo << 16)
                const int e = threadIdx.x + it * THREADS;
                float4* slot = reinterpret_cast<float4*>(tile + (e / Cfg::V4_PER_ROW) * RS + 4 * (e % Cfg::V4_PER_ROW));
                *reinterpret_cast<uint4*>(slot) = pack_split4(*slot);
            }
        }
        __syncthreads();                               // tile t has landed and is converted
        float acc[PPW][2][NT][4];
#pragma unroll
        for (int i = 0; i < PPW; ++i)
#pragma unroll
            for (int hb = 0; hb < 2; ++hb)
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[i][hb][n][e] = 0.f;
        int v = 0;                                     // chunk index within this tile's sequence
#pragma unroll
        for (int i = 0; i < PPW; ++i) {
            const int a = a0 + i;
            for (int q = 0; q <= a; ++q, ++u, ++v) {
                cp_async_wait_1();
                issue(v + 2, (u + 2) % Cfg::STAGES);
                step(q, u % Cfg::STAGES, acc[i][0]);
            }
            for (int q = 0; q < NB - a; ++q, ++u, ++v) {
                cp_async_wait_1();
                issue(v + 2, (u + 2) % Cfg::STAGES);
                step(q, u % Cfg::STAGES, acc[i][1]);
            }
        }
        if (tn < ntiles && threadIdx.x < ROWS) row_table(tn, buf ^ 1);
        __syncthreads();                               // every read of the noise tile is done: overwrite it in place
#pragma unroll
        for (int i = 0; i < PPW; ++i) {
            put(a0 + i, acc[i][0], buf, mrow[i][0]);
            put(NB - 1 - (a0 + i), acc[i][1], buf, mrow[i][1]);
        }
        __syncthreads();
        // store phase: read this thread's slots, refill them with the next tile's noise (in flight while the rows go out)
        const long long n0 = (long long)t * ROWS;
        float4 out[Cfg::IPT];
#pragma unroll
        for (int it = 0; it < Cfg::IPT; ++it) {
            const int e = threadIdx.x + it * THREADS;
            out[it] = *reinterpret_cast<const float4*>(tile + (e / Cfg::V4_PER_ROW) * RS + 4 * (e % Cfg::V4_PER_ROW));
        }
        if (tn < ntiles) {
#pragma unroll
            for (int it = 0; it < Cfg::IPT; ++it) {
                const int e = threadIdx.x + it * THREADS;
                fill_slot(buf ^ 1, e / Cfg::V4_PER_ROW, e % Cfg::V4_PER_ROW);   // same thread, same slot as the read above
            }
        }
        cp_async_commit();
#pragma unroll
        for (int it = 0; it < Cfg::IPT; ++it) {
            const int e = threadIdx.x + it * THREADS;
            const int r = e / Cfg::V4_PER_ROW, v = e % Cfg::V4_PER_ROW;
            if (n0 + r < Ntot) *reinterpret_cast<float4*>(x + (size_t)(n0 + r) * M + 4 * v) = out[it];
        }
    }
    cp_async_wait_all();
}

// LkT[j][k][i] = L[(t,a,j),(t',b,j)] with i = 2t+a, k = 2t'+b;  *bad |= 1 if any entry that the structured sampler
// drops (different dof, or above the diagonal) is not exactly zero.
__global__ void kron_pack_kernel(const float* __restrict__ L, float* __restrict__ LkT, int* __restrict__ bad, int H, int dof) {
    const int D = 2 * dof, M = H * D, N = 2 * H;
    const long long total = (long long)M * M;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(idx / M), col = (int)(idx - (long long)row * M);
        const int t = row / D, rr = row - t * D, a = rr / dof, j = rr - a * dof;
        const int t2 = col / D, cc = col - t2 * D, b = cc / dof, j2 = cc - b * dof;
        const float v = L[idx];
        if (j == j2) {
            const int i = 2 * t + a, k = 2 * t2 + b;
            if (k > i) {
                if (v != 0.f) atomicOr(bad, 1);
                LkT[((size_t)j * N + k) * N + i] = 0.f;
            } else {
                LkT[((size_t)j * N + k) * N + i] = v;
            }
        } else if (v != 0.f) {
            atomicOr(bad, 1);
        }
    }
}

// Tensor-core operand: per dof the largest |entry| (kron_max_kernel, bit pattern of a non-negative float), then the
// fragment-ready fp16 hi/lo words  LkF[j][blk][q][hi|lo][lane][4]  of the scaled block and the inverse scales.
__global__ void kron_max_kernel(const float* __restrict__ LkT, unsigned* __restrict__ maxbits, int N, int dof) {
    const long long total = (long long)dof * N * N;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        atomicMax(maxbits + (int)(idx / ((long long)N * N)), __float_as_uint(fabsf(LkT[idx])));
}
__global__ void kron_frag_kernel(const float* __restrict__ LkT, uint32_t* __restrict__ LkF, const unsigned* __restrict__ maxbits,
                                 int N, int dof) {
    const int NB = N / 16;
    const long long total = (long long)dof * N * N;
    float* inv_scale = reinterpret_cast<float*>(LkF + total);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long r = idx;
        const int reg = (int)(r & 3); r >>= 2;
        const int lane = (int)(r & 31); r >>= 5;
        const int hl = (int)(r & 1); r >>= 1;
        const int q = (int)(r % NB); r /= NB;
        const int blk = (int)(r % NB);
        const int j = (int)(r / NB);
        const float mx = __uint_as_float(maxbits[j]);
        const float scale = mx > 0.f ? exp2f((float)(13 - ilogbf(mx))) : 1.f;
        if (blk == 0 && q == 0 && hl == 0 && lane == 0 && reg == 0) inv_scale[j] = 1.f / scale;
        const int g = lane >> 2, t4 = lane & 3;
        const int i = 16 * blk + g + 8 * (reg & 1);
        const int k0 = 16 * q + t4 + (reg >= 2 ? 8 : 0);
        const float* Lj = LkT + (size_t)j * N * N;
        const uint32_t w0 = split_f16(Lj[(size_t)k0 * N + i] * scale);
        const uint32_t w1 = split_f16(Lj[(size_t)(k0 + 4) * N + i] * scale);
        LkF[idx] = hl ? ((w0 >> 16) | (w1 & 0xffff0000u)) : ((w0 & 0xffffu) | (w1 << 16));
    }
}

template <int DOF, int H, int KIND>      // KIND 0: exact FP32, 1: warp MMA on injected noise, 2: warp MMA with in-kernel Philox noise
static int launch_kron(const void* Lp, const float* mu, const float* eps, float* x, int P, int S, const NoiseArgs& noise,
                       cudaStream_t st) {
    constexpr bool MMA = KIND != 0;
    const void* kern = KIND == 2 ? (const void*)sample_gp_kron_mma_kernel<DOF, H, true>
                     : KIND == 1 ? (const void*)sample_gp_kron_mma_kernel<DOF, H, false> : (const void*)sample_gp_kron_kernel<DOF, H>;
    const size_t smem_bytes = MMA ? KronMmaCfg<DOF, H>::SMEM : KronCfg<DOF, H>::SMEM;
    const int threads = MMA ? KronMmaCfg<DOF, H>::THREADS : KronCfg<DOF, H>::THREADS;
    const int tile_rows = MMA ? KronMmaCfg<DOF, H>::ROWS : kTileRows;
    static thread_local int cached_dev = -1, per_sm = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) { set_error("mpb_sample_gp_kron: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
        if (MMA) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int n = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem_bytes);
        if (e != cudaSuccess || n < 1) { set_error("mpb_sample_gp_kron: kernel does not fit on this device"); return MPB_ECUDA; }
        per_sm = n;
        cached_dev = dev;
    }
    const long long ntiles = ((long long)P * S + tile_rows - 1) / tile_rows;
    const long long cap = (long long)sm_count() * per_sm;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    void* args[] = {(void*)&Lp, (void*)&mu, (void*)&eps, (void*)&x, (void*)&P, (void*)&S, (void*)&noise};
    cudaError_t le = cudaLaunchKernel(kern, dim3(grid), dim3(threads), args, smem_bytes, st);
    if (le != cudaSuccess) { set_error("mpb_sample_gp_kron: %s", cudaGetErrorString(le)); return MPB_ECUDA; }
    return check_launch("mpb_sample_gp_kron");
}

}  // namespace mpb

// Any (dof, H) with H % 16 == 0 and dof * H <= 512 (one warp per dof and 32-row output pair: <= 1024 threads) fits the
// kernels; the list instantiates the horizons the reference's examples use (32 / 64 / 128 support points, plus 16 and 48
// for short horizons) for 2-D / 3-D point robots and chains of 4..8 joints.  Other shapes take the dense samplers.
#define MPB_KRON_SHAPES(X)                                                                            \
    X(2, 16) X(2, 32) X(2, 48) X(2, 64) X(2, 128) X(3, 16) X(3, 32) X(3, 48) X(3, 64) X(3, 128)      \
    X(4, 32) X(4, 64) X(5, 32) X(5, 64) X(6, 32) X(6, 64) X(7, 16) X(7, 32) X(7, 48) X(7, 64)         \
    X(8, 32) X(8, 64)

extern "C" int mpb_sample_gp_kron_supported(int H, int dof) {
#define X(d, h) if (dof == d && H == h) return 1;
    MPB_KRON_SHAPES(X)
#undef X
    return 0;
}

extern "C" int mpb_sample_gp_kron_pack(const float* L, float* LkT, int H, int dof, int* structured, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(L && LkT && structured, "mpb_sample_gp_kron_pack: null pointer");
    MPB_REQUIRE(H >= 1 && dof >= 1 && (long long)H * dof <= 16384, "mpb_sample_gp_kron_pack: bad sizes H=%d dof=%d", H, dof);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int* bad = nullptr;
    cudaError_t e = cudaMalloc(&bad, sizeof(int));
    if (e != cudaSuccess) { set_error("mpb_sample_gp_kron_pack: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    int h_bad = 1;
    e = cudaMemsetAsync(bad, 0, sizeof(int), st);
    if (e == cudaSuccess) {
        const long long total = (long long)H * 2 * dof * H * 2 * dof;
        const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        kron_pack_kernel<<<grid, 256, 0, st>>>(L, LkT, bad, H, dof);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(bad);
    if (e != cudaSuccess) { set_error("mpb_sample_gp_kron_pack: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    *structured = h_bad ? 0 : 1;
    return MPB_OK;
}

static int sample_gp_kron_any(int kind, const void* LkT, const float* mu, const float* eps, float* x, int P, int S, int H,
                              int dof, const mpb_noise_desc* nd, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(LkT && mu && x && (kind == 2 ? nd != nullptr : eps != nullptr), "mpb_sample_gp_kron: null pointer");
    NoiseArgs noise{};
    if (kind == 2) {
        const char* why = noise_args(*nd, P, noise);
        MPB_REQUIRE(!why, "mpb_sample_gp_kron_tc_rng: %s", why);
    }
    MPB_REQUIRE(P >= 0 && S >= 0, "mpb_sample_gp_kron: bad sizes P=%d S=%d", P, S);
    MPB_REQUIRE(mpb_sample_gp_kron_supported(H, dof), "mpb_sample_gp_kron: shape H=%d dof=%d has no structured kernel", H, dof);
    MPB_REQUIRE(((uintptr_t)LkT | (uintptr_t)mu | (uintptr_t)eps | (uintptr_t)x) % 16 == 0, "mpb_sample_gp_kron: pointers must be 16-byte aligned");   // eps == NULL (rng) passes
    if (P == 0 || S == 0) return MPB_OK;
    MPB_REQUIRE((long long)P * S <= 0x7fffffffLL / 2, "mpb_sample_gp_kron: P*S too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define X(d, h)                                                                                  \
    if (dof == d && H == h)                                                                      \
        return kind == 2   ? launch_kron<d, h, 2>(LkT, mu, eps, x, P, S, noise, st)              \
               : kind == 1 ? launch_kron<d, h, 1>(LkT, mu, eps, x, P, S, noise, st)              \
                           : launch_kron<d, h, 0>(LkT, mu, eps, x, P, S, noise, st);
    MPB_KRON_SHAPES(X)
#undef X
    return MPB_EINVAL;
}

extern "C" int mpb_sample_gp_kron(const float* LkT, const float* mu, const float* eps, float* x, int P, int S, int H,
                                  int dof, void* stream) {
    return sample_gp_kron_any(0, LkT, mu, eps, x, P, S, H, dof, nullptr, stream);
}

extern "C" long long mpb_sample_gp_kron_tc_bytes(int H, int dof) {
    return (long long)dof * 2 * H * 2 * H * 4 + 64 * 4;        // fragments + inverse scales (16) + scratch (48 words)
}

extern "C" int mpb_sample_gp_kron_tc_prepare(const float* LkT, void* LkF, int H, int dof, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(LkT && LkF, "mpb_sample_gp_kron_tc_prepare: null pointer");
    MPB_REQUIRE(mpb_sample_gp_kron_supported(H, dof) && dof <= 16, "mpb_sample_gp_kron_tc_prepare: shape H=%d dof=%d has no structured kernel", H, dof);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int N = 2 * H;
    const long long total = (long long)dof * N * N;
    uint32_t* frag = static_cast<uint32_t*>(LkF);
    unsigned* maxbits = frag + total + 16;
    cudaError_t e = cudaMemsetAsync(frag + total, 0, 64 * 4, st);
    if (e != cudaSuccess) { set_error("mpb_sample_gp_kron_tc_prepare: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    const int grid = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
    kron_max_kernel<<<grid, 256, 0, st>>>(LkT, maxbits, N, dof);
    kron_frag_kernel<<<grid, 256, 0, st>>>(LkT, frag, maxbits, N, dof);
    return check_launch("mpb_sample_gp_kron_tc_prepare");
}

extern "C" int mpb_sample_gp_kron_tc(const void* LkF, const float* mu, const float* eps, float* x, int P, int S, int H,
                                     int dof, void* stream) {
    return sample_gp_kron_any(1, LkF, mu, eps, x, P, S, H, dof, nullptr, stream);
}

extern "C" int mpb_sample_gp_kron_tc_rng(const void* LkF, const float* mu, const mpb_noise_desc* noise, float* x, int P, int S,
                                         int H, int dof, void* stream) {
    return sample_gp_kron_any(2, LkF, mu, nullptr, x, P, S, H, dof, noise, stream);
}
