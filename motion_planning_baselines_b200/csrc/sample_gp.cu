// K1 (FP32 SIMT variant): GP-prior sampling  x[p,s,:] = mu[p,:] + L @ eps[s,p,:].
//
// Replaces MultiMPPrior.sample (mp_baselines/planners/costs/factors/mp_priors_multi.py:253-256), i.e.
// torch MultivariateNormal.rsample's broadcast batched mat-vec against a per-particle [P,M,M]
// scale_tril, by ONE triangular GEMM  X[N,M] = E[N,M] * L^T  (N = P*S rows) that reads the single
// [M,M] factor: k-tiles above the diagonal are skipped, the epilogue adds mu_p and writes the
// sample-major reference layout [S,P,M] back as particle-major [P,S,M].
//
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile split 4+4 so that the 128-bit shared-memory
// reads are bank-conflict free; global->shared prefetch is double buffered through registers.
// Bound: FP32 FMA.  (The tcgen05 3xTF32 variant lives in sample_gp_tc.cu.)
#include "mpb_common.cuh"
#include "philox.cuh"

namespace mpb {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

__global__ void __launch_bounds__(256) sample_gp_simt_kernel(const float* __restrict__ L, const float* __restrict__ mu,
                                                             const float* __restrict__ eps, float* __restrict__ x,
                                                             int P, int S, int M) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];   // eps tile, k-major
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];   // L tile,   k-major
    const int N = P * S;
    const int n0 = blockIdx.x * BM, i0 = blockIdx.y * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const bool vec = (M & 3) == 0;

    // each thread moves two 4-float k-segments of A and of B per k-tile
    int lrow[2], lk[2];
    const float* asrc[2];
    const float* bsrc[2];
    bool aok[2], bok[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int idx = tid * 2 + j;
        lrow[j] = idx >> 2;
        lk[j] = (idx & 3) * 4;
        const int n = n0 + lrow[j];
        aok[j] = n < N;
        const int p = aok[j] ? n / S : 0, s = aok[j] ? n - p * S : 0;
        asrc[j] = eps + ((size_t)s * P + p) * M;
        const int i = i0 + lrow[j];
        bok[j] = i < M;
        bsrc[j] = L + (size_t)(bok[j] ? i : 0) * M;
    }
    const int k_end = min(M, i0 + BN);                  // L[i,k] = 0 for k > i
    const int n_kt = (k_end + BK - 1) / BK;

    float4 ra[2], rb[2];
    auto gload = [&](int kt) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int k = kt * BK + lk[j];
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (vec) {
                if (aok[j] && k < M) va = __ldg(reinterpret_cast<const float4*>(asrc[j] + k));
                if (bok[j] && k < M) vb = __ldg(reinterpret_cast<const float4*>(bsrc[j] + k));
            } else {
                float ta[4] = {0.f, 0.f, 0.f, 0.f}, tb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (aok[j] && k + c < M) ta[c] = __ldg(asrc[j] + k + c);
                    if (bok[j] && k + c < M) tb[c] = __ldg(bsrc[j] + k + c);
                }
                va = make_float4(ta[0], ta[1], ta[2], ta[3]);
                vb = make_float4(tb[0], tb[1], tb[2], tb[3]);
            }
            ra[j] = va;
            rb[j] = vb;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            As[buf][lk[j] + 0][lrow[j]] = ra[j].x; As[buf][lk[j] + 1][lrow[j]] = ra[j].y;
            As[buf][lk[j] + 2][lrow[j]] = ra[j].z; As[buf][lk[j] + 3][lrow[j]] = ra[j].w;
            Bs[buf][lk[j] + 0][lrow[j]] = rb[j].x; Bs[buf][lk[j] + 1][lrow[j]] = rb[j].y;
            Bs[buf][lk[j] + 2][lrow[j]] = rb[j].z; Bs[buf][lk[j] + 3][lrow[j]] = rb[j].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < n_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_kt) gload(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
        }
        if (kt + 1 < n_kt) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue: + mu_p, particle-major store
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int n = n0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
        if (n >= N) continue;
        const int p = n / S;
        const float* mrow = mu + (size_t)p * M;
        float* xrow = x + (size_t)n * M;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = i0 + h * 64 + tx * 4;
            if (vec && i + 3 < M) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(mrow + i));
                float4 o;
                o.x = m4.x + acc[r][h * 4 + 0]; o.y = m4.y + acc[r][h * 4 + 1];
                o.z = m4.z + acc[r][h * 4 + 2]; o.w = m4.w + acc[r][h * 4 + 3];
                *reinterpret_cast<float4*>(xrow + i) = o;
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (i + c < M) xrow[i + c] = __ldg(mrow + i + c) + acc[r][h * 4 + c];
            }
        }
    }
}

// STOMP noise: x[p,s,h,j] = mu[p,h,j] + (endpoint ? 0 : sum_{k<=h} L_R[h,k] eps[s,j,p,k])
// (mp_baselines/planners/stomp.py:97-108).  One CTA stages L_R once and loops over (p,s) pairs.
// GEN: eps is unused, the noise block is drawn in place (layout MPB_NOISE_STOMP, philox.cuh).
template <bool GEN>
__global__ void __launch_bounds__(256) sample_stomp_kernel(const float* __restrict__ LR, const float* __restrict__ mu,
                                                           const float* __restrict__ eps, float* __restrict__ x,
                                                           int P, int S, int H, int D, const NoiseArgs noise) {
    extern __shared__ __align__(16) float sm[];
    float* Ls = sm;                         // [H][H+1]
    float* es = sm + H * (H + 1);           // [D][H+1]
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) Ls[(i / H) * (H + 1) + (i % H)] = LR[i];
    const int M = H * D;
    for (int ps = blockIdx.x; ps < P * S; ps += gridDim.x) {
        const int p = ps / S, s = ps - p * S;
        __syncthreads();
        if (GEN) {
            // global element of (s, j, p, k):  (((s_off+s) D + j) P_glob + p_off+p) H + k
            if ((H & 3) == 0) {
                const int HQ = H >> 2;
                for (int i = threadIdx.x; i < D * HQ; i += blockDim.x) {
                    const int j = i / HQ, kq = i - j * HQ;
                    const unsigned long long e = ((unsigned long long)((noise.s_off + s) * D + j) * noise.P_glob + noise.p_off + p) * H + 4 * kq;
                    const float4 q = philox_normal4(e >> 2, noise);
                    float* d = es + j * (H + 1) + 4 * kq;
                    d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
                }
            } else {
                for (int i = threadIdx.x; i < M; i += blockDim.x) {
                    const int j = i / H, k = i - j * H;
                    const unsigned long long e = ((unsigned long long)((noise.s_off + s) * D + j) * noise.P_glob + noise.p_off + p) * H + k;
                    es[j * (H + 1) + k] = philox_normal1(e, noise);
                }
            }
        } else {
            for (int i = threadIdx.x; i < M; i += blockDim.x) {      // i = j*H + k
                const int j = i / H, k = i - j * H;
                es[j * (H + 1) + k] = __ldg(eps + (((size_t)s * D + j) * P + p) * H + k);
            }
        }
        __syncthreads();
        for (int o = threadIdx.x; o < M; o += blockDim.x) {      // o = h*D + j
            const int h = o / D, j = o - h * D;
            float acc = 0.f;
            if (h != 0 && h != H - 1) {
                const float* lrow = Ls + h * (H + 1);
                const float* erow = es + j * (H + 1);
                for (int k = 0; k <= h; ++k) acc = fmaf(lrow[k], erow[k], acc);
            }
            x[(size_t)ps * M + o] = __ldg(mu + (size_t)p * M + o) + acc;
        }
    }
}

// Wide variant for many samples (the C5 sweep: 10^4 - 10^6 samples of one problem): one THREAD per (particle, sample,
// state column j) keeps its H = 64 noise values in registers and walks the rows of L_R, which every thread reads at the
// same address (shared-memory broadcast, 16 bytes per 4 FMAs).  Rows are cut into chunks of 16 k; a chunk right of the
// diagonal is skipped, inside the diagonal chunk the exact zeros of the lower-triangular factor make the extra FMAs
// no-ops, so the sum is the one of sample_stomp_kernel bit for bit (same ascending k order).  The CTA-per-sample kernel
// above needs two barriers and a 224-thread Philox phase per sample: 1.07 ms for 10^5 Panda samples, this one 0.2 ms.
template <bool GEN>
__global__ void __launch_bounds__(128) sample_stomp_wide_kernel(const float* __restrict__ LR, const float* __restrict__ mu,
                                                                const float* __restrict__ eps, float* __restrict__ x,
                                                                int P, int S, int D, const NoiseArgs noise) {
    constexpr int H = 64;
    __shared__ __align__(16) float Ls[H * H];
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) Ls[i] = LR[i];
    __syncthreads();
    const long long total = (long long)P * S * D;
    const int M = H * D;
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(id % D);
        const long long ps = id / D;
        const int p = (int)(ps / S), s = (int)(ps - (long long)p * S);
        float e[H];
        if (GEN) {
            const unsigned long long g0 = (((unsigned long long)((noise.s_off + s) * D + j) * noise.P_glob + noise.p_off + p) * H) >> 2;
#pragma unroll
            for (int q = 0; q < H / 4; ++q) {
                const float4 v = philox_normal4(g0 + q, noise);
                e[4 * q] = v.x; e[4 * q + 1] = v.y; e[4 * q + 2] = v.z; e[4 * q + 3] = v.w;
            }
        } else {
            const float4* src = reinterpret_cast<const float4*>(eps + (((size_t)s * D + j) * P + p) * H);
#pragma unroll
            for (int q = 0; q < H / 4; ++q) {
                const float4 v = __ldg(src + q);
                e[4 * q] = v.x; e[4 * q + 1] = v.y; e[4 * q + 2] = v.z; e[4 * q + 3] = v.w;
            }
        }
        const float* mp = mu + (size_t)p * M + j;
        float* xp = x + (size_t)ps * M + j;
        xp[0] = __ldg(mp);
        xp[(size_t)(H - 1) * D] = __ldg(mp + (size_t)(H - 1) * D);
#pragma unroll 1
        for (int h = 1; h < H - 1; ++h) {
            const float4* lrow = reinterpret_cast<const float4*>(Ls + h * H);
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < H / 16; ++c) {
                if (16 * c <= h) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 l = lrow[4 * c + u];
                        acc = fmaf(l.x, e[16 * c + 4 * u], acc);
                        acc = fmaf(l.y, e[16 * c + 4 * u + 1], acc);
                        acc = fmaf(l.z, e[16 * c + 4 * u + 2], acc);
                        acc = fmaf(l.w, e[16 * c + 4 * u + 3], acc);
                    }
                }
            }
            xp[(size_t)h * D] = __ldg(mp + (size_t)h * D) + acc;
        }
    }
}

// The normals a kernel consumes for a noise descriptor, written in the LOCAL layout of the call (mpb_philox_normal).
// Quads of the innermost axis are group aligned when its extent is a multiple of 4; otherwise element by element.
__global__ void __launch_bounds__(256) philox_dump_kernel(const NoiseArgs noise, int layout, float* __restrict__ out,
                                                          long long n0, long long n1, long long n2, long long n3, int vec) {
    const long long inner = layout == MPB_NOISE_STOMP ? n3 : n2;
    const long long total = n0 * n1 * n2 * (layout == MPB_NOISE_STOMP ? n3 : 1);
    const long long items = vec ? total >> 2 : total;
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
        const long long flat = vec ? it << 2 : it;
        const long long k = flat % inner;
        long long r = flat / inner;
        unsigned long long e;
        if (layout == MPB_NOISE_SPM) {                 // r = s * P + p
            const long long p = r % n1, s = r / n1;
            e = ((unsigned long long)(noise.s_off + s) * noise.P_glob + noise.p_off + p) * inner + k;
        } else if (layout == MPB_NOISE_SPMD) {         // local [S,P,M] with m = n * dof + j; virtual [S_glob,P_glob,dof,M/dof]
            const long long p = r % n1, s = r / n1, dof = n3, j = k % dof, n = k / dof;
            e = (((unsigned long long)(noise.s_off + s) * noise.P_glob + noise.p_off + p) * dof + j) * (inner / dof) + n;
        } else if (layout == MPB_NOISE_STOMP) {        // r = (s * D + j) * P + p
            const long long p = r % n2; r /= n2;
            const long long j = r % n1, s = r / n1;
            e = ((unsigned long long)((noise.s_off + s) * n1 + j) * noise.P_glob + noise.p_off + p) * inner + k;
        } else {                                       // MPPI: r = i * N + n
            const long long n = r % n1, i = r / n1;
            e = ((unsigned long long)i * noise.P_glob + noise.s_off + n) * inner + k;
        }
        if (vec) *reinterpret_cast<float4*>(out + flat) = philox_normal4(e >> 2, noise);
        else out[flat] = philox_normal1(e, noise);
    }
}

// y[p,:] = Sigma_inv @ mu[p,:] for banded Sigma_inv, fp64 accumulation.  Sigma_inv = A^T Q^-1 A is symmetric, so
// thread i walks COLUMN i of the band (Sigma_inv[j][i], j = i-hbw..i+hbw): consecutive threads read consecutive
// addresses (coalesced) where walking row i would stride by M.  A CTA serves kMvP particles at once (their mu rows
// staged in shared memory) so that every band entry is fetched once per kMvP particles.
constexpr int kMvP = 8;
__global__ void __launch_bounds__(256) prior_matvec_kernel(const float* __restrict__ Sinv, const float* __restrict__ mu,
                                                           float* __restrict__ y, int P, int M, int hbw) {
    extern __shared__ __align__(16) float ms[];
    const int p0 = blockIdx.y * kMvP;
    const int i0 = blockIdx.x * blockDim.x;
    const int lo_j = max(0, i0 - hbw), hi_j = min(M - 1, i0 + (int)blockDim.x - 1 + hbw);
    const int span = blockDim.x + 2 * hbw;
    for (int q = 0; q < kMvP; ++q) {
        const int p = min(p0 + q, P - 1);
        for (int j = lo_j + threadIdx.x; j <= hi_j; j += blockDim.x) ms[q * span + j - lo_j] = __ldg(mu + (size_t)p * M + j);
    }
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i >= M) return;
    const int j0 = max(0, i - hbw), j1 = min(M - 1, i + hbw);
    // Terms reach 1e7 and cancel to O(1): accumulate in double-float (hi + lo fp32 pairs, error-free TwoProd / TwoSum).
    // Native FP64 issues at 1/32 of the FP32 rate on this part and made this kernel FP64-pipe bound.
    float hi[kMvP], lo[kMvP];
#pragma unroll
    for (int q = 0; q < kMvP; ++q) hi[q] = lo[q] = 0.f;
    constexpr int kU = 6;                // band entries fetched together: the loads are independent, the sums are not
    for (int jb = j0; jb <= j1; jb += kU) {
        float svv[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) svv[u] = (jb + u <= j1) ? __ldg(Sinv + (size_t)(jb + u) * M + i) : 0.f;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
        const int j = min(jb + u, j1);
        const float sv = svv[u];
#pragma unroll
        for (int q = 0; q < kMvP; ++q) {
            const float m = ms[q * span + j - lo_j];
            const float p = __fmul_rn(sv, m);
            const float e = fmaf(sv, m, -p);                      // sv*m = p + e exactly
            const float t = __fadd_rn(hi[q], p);
            const float z = __fsub_rn(t, hi[q]);
            lo[q] = __fadd_rn(lo[q], __fadd_rn(__fadd_rn(__fsub_rn(hi[q], __fsub_rn(t, z)), __fsub_rn(p, z)), e));
            hi[q] = t;
        }
        }
    }
#pragma unroll
    for (int q = 0; q < kMvP; ++q)
        if (p0 + q < P) y[(size_t)(p0 + q) * M + i] = __fadd_rn(hi[q], lo[q]);
}

// Structured variant of prior_matvec_kernel.  With the state order (t, [pos|vel], j) the index is (2t+a)*dof + j, and the
// reference's precision A^T Q^-1 A (mp_priors_multi.py:213-251) couples only entries of the same dof j whose block
// indices 2t+a differ by at most 3: row i has at most 7 non-zeros, at columns i + m*dof, m = -3..3, instead of the
// 2*(2D-1)+1 = 55 the band walk visits.  mpb_prior_dof_structured verifies the zero pattern bit-exactly; skipping exact
// zeros leaves every double-float partial sum unchanged, so the result is bit-identical to prior_matvec_kernel.
constexpr int kMvDofP = 4;
__global__ void __launch_bounds__(256) prior_matvec_dof_kernel(const float* __restrict__ Sinv, const float* __restrict__ mu,
                                                               float* __restrict__ y, int P, int M, int dof) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = blockIdx.y * kMvDofP;
    if (i >= M) return;
    float sv[7];
#pragma unroll
    for (int m = 0; m < 7; ++m) {
        const int j = i + (m - 3) * dof;
        sv[m] = (j >= 0 && j < M) ? __ldg(Sinv + (size_t)j * M + i) : 0.f;      // symmetric: column i read row-wise (coalesced)
    }
    float mv[kMvDofP][7];
#pragma unroll
    for (int q = 0; q < kMvDofP; ++q) {
        const float* mrow = mu + (size_t)min(p0 + q, P - 1) * M;
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            const int j = i + (m - 3) * dof;
            mv[q][m] = (j >= 0 && j < M) ? __ldg(mrow + j) : 0.f;
        }
    }
#pragma unroll
    for (int q = 0; q < kMvDofP; ++q) {
        float hi = 0.f, lo = 0.f;
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            const float p = __fmul_rn(sv[m], mv[q][m]);
            const float e = fmaf(sv[m], mv[q][m], -p);
            const float t = __fadd_rn(hi, p);
            const float z = __fsub_rn(t, hi);
            lo = __fadd_rn(lo, __fadd_rn(__fadd_rn(__fsub_rn(hi, __fsub_rn(t, z)), __fsub_rn(p, z)), e));
            hi = t;
        }
        if (p0 + q < P) y[(size_t)(p0 + q) * M + i] = __fadd_rn(hi, lo);
    }
}

// *bad |= 1 if Sinv[r][c] != 0 for some pair that the structured mat-vec skips
__global__ void prior_dof_check_kernel(const float* __restrict__ Sinv, int* __restrict__ bad, int M, int dof) {
    const long long total = (long long)M * M;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / M), c = (int)(idx - (long long)r * M);
        const int dj = (r % dof) - (c % dof), db = r / dof - c / dof;
        if ((dj != 0 || db > 3 || db < -3) && Sinv[idx] != 0.f) atomicOr(bad, 1);
    }
}

}  // namespace mpb

extern "C" int mpb_sample_gp(const float* L, const float* mu, const float* eps, float* x, int P, int S, int M,
                             void* stream) {
    using namespace mpb;
    MPB_REQUIRE(L && mu && eps && x, "mpb_sample_gp: null pointer");
    MPB_REQUIRE(P >= 0 && S >= 0 && M >= 1, "mpb_sample_gp: bad sizes P=%d S=%d M=%d", P, S, M);
    if (P == 0 || S == 0) return MPB_OK;
    MPB_REQUIRE((long long)P * S <= 0x7fffffffLL / 2, "mpb_sample_gp: P*S too large");
    dim3 grid((P * S + BM - 1) / BM, (M + BN - 1) / BN);
    sample_gp_simt_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(L, mu, eps, x, P, S, M);
    return check_launch("mpb_sample_gp");
}

extern "C" int mpb_philox_normal(const mpb_noise_desc* nd, int layout, float* out, int n0, int n1, int n2, int n3, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(nd && out, "mpb_philox_normal: null pointer");
    MPB_REQUIRE(layout == MPB_NOISE_SPM || layout == MPB_NOISE_STOMP || layout == MPB_NOISE_MPPI || layout == MPB_NOISE_SPMD,
                "mpb_philox_normal: unknown layout %d", layout);
    MPB_REQUIRE(layout != MPB_NOISE_SPMD || (n3 >= 1 && n2 % n3 == 0), "mpb_philox_normal: MPB_NOISE_SPMD takes the dof count in n3 (M %% dof == 0)");
    MPB_REQUIRE(n0 >= 0 && n1 >= 0 && n2 >= 0 && (layout != MPB_NOISE_STOMP || n3 >= 0), "mpb_philox_normal: negative extent");
    NoiseArgs noise{};
    const char* why = noise_args(*nd, 0, noise);
    MPB_REQUIRE(!why, "mpb_philox_normal: %s", why);
    const long long inner = layout == MPB_NOISE_STOMP ? n3 : n2;
    const long long total = (long long)n0 * n1 * n2 * (layout == MPB_NOISE_STOMP ? n3 : 1);
    if (total == 0) return MPB_OK;
    const int vec = (inner % 4 == 0) && ((uintptr_t)out % 16 == 0) && layout != MPB_NOISE_SPMD;
    const long long items = vec ? total / 4 : total;
    const long long blocks = (items + 255) / 256;
    const int grid = (int)(blocks < (long long)sm_count() * 8 ? blocks : (long long)sm_count() * 8);
    philox_dump_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(noise, layout, out, n0, n1, n2, n3, vec);
    return check_launch("mpb_philox_normal");
}

static int sample_stomp_any(const float* L_R, const float* mu, const float* eps, const mpb_noise_desc* nd, float* x, int P,
                            int S, int H, int D, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(L_R && mu && (eps || nd) && x, "mpb_sample_stomp: null pointer");
    MPB_REQUIRE(P >= 0 && S >= 0 && H >= 2 && D >= 1, "mpb_sample_stomp: bad sizes");
    if (P == 0 || S == 0) return MPB_OK;
    const size_t smem = (size_t)(H + D) * (H + 1) * sizeof(float);
    MPB_REQUIRE(smem <= 200 * 1024, "mpb_sample_stomp: H=%d too large for shared memory", H);
    NoiseArgs noise{};
    if (!eps) {
        const char* why = noise_args(*nd, P, noise);
        MPB_REQUIRE(!why, "mpb_sample_stomp_rng: %s", why);
    }
    cudaError_t e = eps ? cudaFuncSetAttribute(sample_stomp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                        : cudaFuncSetAttribute(sample_stomp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mpb_sample_stomp: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    const long long work = (long long)P * S;
    // many samples: one thread per (particle, sample, column); the factor must be lower triangular with exact zeros above
    // the diagonal (torch's scale_tril is), the injected noise 16-byte aligned
    if (H == 64 && work * D >= 16384 && (!eps || (reinterpret_cast<uintptr_t>(eps) & 15) == 0)) {
        const long long threads = work * D;
        const long long blocks = (threads + 127) / 128;
        const int wgrid = (int)(blocks < (long long)sm_count() * 16 ? blocks : (long long)sm_count() * 16);
        if (eps) sample_stomp_wide_kernel<false><<<wgrid, 128, 0, static_cast<cudaStream_t>(stream)>>>(L_R, mu, eps, x, P, S, D, noise);
        else sample_stomp_wide_kernel<true><<<wgrid, 128, 0, static_cast<cudaStream_t>(stream)>>>(L_R, mu, eps, x, P, S, D, noise);
        return check_launch("mpb_sample_stomp");
    }
    const int grid = (int)(work < (long long)sm_count() * 4 ? work : (long long)sm_count() * 4);
    if (eps) sample_stomp_kernel<false><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(L_R, mu, eps, x, P, S, H, D, noise);
    else sample_stomp_kernel<true><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(L_R, mu, eps, x, P, S, H, D, noise);
    return check_launch("mpb_sample_stomp");
}

extern "C" int mpb_sample_stomp(const float* L_R, const float* mu, const float* eps, float* x, int P, int S, int H,
                                int D, void* stream) {
    MPB_REQUIRE(eps, "mpb_sample_stomp: eps is null (use mpb_sample_stomp_rng for in-kernel noise)");
    return sample_stomp_any(L_R, mu, eps, nullptr, x, P, S, H, D, stream);
}

extern "C" int mpb_sample_stomp_rng(const float* L_R, const float* mu, const mpb_noise_desc* noise, float* x, int P, int S,
                                    int H, int D, void* stream) {
    MPB_REQUIRE(noise, "mpb_sample_stomp_rng: noise descriptor is null");
    return sample_stomp_any(L_R, mu, nullptr, noise, x, P, S, H, D, stream);
}

extern "C" int mpb_prior_matvec(const float* Sigma_inv, const float* mu, float* y, int P, int M, int half_bw,
                                void* stream) {
    using namespace mpb;
    MPB_REQUIRE(Sigma_inv && mu && y, "mpb_prior_matvec: null pointer");
    MPB_REQUIRE(P >= 0 && M >= 1 && half_bw >= 0, "mpb_prior_matvec: bad sizes");
    if (P == 0) return MPB_OK;
    MPB_REQUIRE(P <= 65535 * kMvP, "mpb_prior_matvec: too many particles per call");
    dim3 grid((M + 255) / 256, (P + kMvP - 1) / kMvP);
    const size_t smem = (size_t)kMvP * (256 + 2 * half_bw) * sizeof(float);
    MPB_REQUIRE(smem <= 48 * 1024, "mpb_prior_matvec: half bandwidth %d too large", half_bw);
    prior_matvec_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(Sigma_inv, mu, y, P, M, half_bw);
    return check_launch("mpb_prior_matvec");
}

extern "C" int mpb_prior_dof_structured(const float* Sigma_inv, int H, int dof, int* structured, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(Sigma_inv && structured, "mpb_prior_dof_structured: null pointer");
    MPB_REQUIRE(H >= 1 && dof >= 1 && (long long)H * dof <= 16384, "mpb_prior_dof_structured: bad sizes H=%d dof=%d", H, dof);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int M = 2 * H * dof;
    int* bad = nullptr;
    cudaError_t e = cudaMalloc(&bad, sizeof(int));
    if (e != cudaSuccess) { set_error("mpb_prior_dof_structured: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    int h_bad = 1;
    e = cudaMemsetAsync(bad, 0, sizeof(int), st);
    if (e == cudaSuccess) {
        const long long total = (long long)M * M;
        prior_dof_check_kernel<<<(int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096), 256, 0, st>>>(Sigma_inv, bad, M, dof);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(bad);
    if (e != cudaSuccess) { set_error("mpb_prior_dof_structured: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    *structured = h_bad ? 0 : 1;
    return MPB_OK;
}

extern "C" int mpb_prior_matvec_dof(const float* Sigma_inv, const float* mu, float* y, int P, int H, int dof, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(Sigma_inv && mu && y, "mpb_prior_matvec_dof: null pointer");
    MPB_REQUIRE(P >= 0 && H >= 1 && dof >= 1, "mpb_prior_matvec_dof: bad sizes");
    if (P == 0) return MPB_OK;
    const int M = 2 * H * dof;
    MPB_REQUIRE(P <= 65535 * kMvDofP, "mpb_prior_matvec_dof: too many particles per call");
    dim3 grid((M + 255) / 256, (P + kMvDofP - 1) / kMvDofP);
    prior_matvec_dof_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(Sigma_inv, mu, y, P, M, dof);
    return check_launch("mpb_prior_matvec_dof");
}
