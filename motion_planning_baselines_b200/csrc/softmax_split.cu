// K3b: the importance-weight update split over CTAs (and over GPUs) through packed partial records.
//
// Same maths as softmax_update.cu (StochGPMP._update_distribution, mp_baselines/planners/stoch_gpmp.py:267-279;
// STOMP._update_distribution, stomp.py:199-220; MPPI.update_controller + _save_best, mppi.py:72-86,164-169) for the
// cases one CTA per particle cannot serve: one problem with 10^3..10^6 samples (BASELINE.json configs[4]) and one
// problem whose samples are sharded over several GPUs.
//
//   partial : for particle p and a chunk of its samples,
//               m = max_s a_s, a_s = -cost_s/temp;   Z = sum_s exp(a_s - m);   v = sum_s exp(a_s - m) (x_s - mu)
//               (cmin, argmin) with first-occurrence ties (torch.argmin, mppi.py:166)
//             packed as one record [m, Z, cmin, argmin(int bits), v[0..Mw)].
//   combine : log-sum-exp merge of R records per particle in FIXED order r = 0..R-1 (deterministic; every GPU that
//             holds the same all-gathered records computes bit-identical means):
//               m* = max_r m_r;  Z* = sum_r Z_r e^{m_r - m*};  g = sum_r e^{m_r - m*} v_r / Z*;  mu += step * (SigmaR g | g)
//   weights : w_s = exp(a_s - m*) / Z*.
// Multi-GPU: records [n_chunks,P,REC] of every rank are concatenated along the first axis by one all-gather
// (chunk-major layout makes that a plain concatenation), then `combine` runs redundantly on every rank.
//
// Samples may be a column window [c0, c0+Dw) of wider rows [H, Dfull] (MPPI: the control part of (state | control)).
// Bound: HBM read of x (partial); combine is latency-bound (R * P small records).
#include "mpb_common.cuh"

namespace mpb {

constexpr int kPartThreads = 256;
constexpr int kRecHead = 4;

__global__ void __launch_bounds__(kPartThreads) softmax_partial_kernel(
    const float* __restrict__ cost, const float* __restrict__ x, const float* __restrict__ mu, float* __restrict__ rec,
    float temp, int P, int S, int H, int Dfull, int c0, int Dw, int chunk, long long sample_offset) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float red_f[32];
    __shared__ int red_i[32];
    const int Mw = H * Dw;
    const int p = blockIdx.y, c = blockIdx.x;
    const int s0 = c * chunk, s1 = min(S, s0 + chunk);
    const int ns = max(0, s1 - s0);
    float* es = sm;                                     // [chunk] unnormalised weights
    float* part = sm + ((chunk + 3) & ~3);              // [nsl][Mw] partial sums
    const float* cp = cost + (size_t)p * S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kPartThreads / 32;

    // ---- chunk max of a = -cost/temp, chunk min of cost with its first index -----------------------------
    float mx = -CUDART_INF_F, cmin = CUDART_INF_F;
    int imin = 0x7fffffff;
    for (int i = threadIdx.x; i < ns; i += kPartThreads) {
        const float cs = __ldg(cp + s0 + i);
        const float a = -cs / temp;
        es[i] = a;
        mx = fmaxf(mx, a);
        if (cs < cmin) { cmin = cs; imin = s0 + i; }    // i ascends per thread: first occurrence kept
    }
    mx = warp_max(mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float oc = __shfl_xor_sync(MPB_FULL_MASK, cmin, o);
        const int oi = __shfl_xor_sync(MPB_FULL_MASK, imin, o);
        if (oc < cmin || (oc == cmin && oi < imin)) { cmin = oc; imin = oi; }
    }
    if (lane == 0) { red_f[warp] = mx; red_f[8 + warp] = cmin; red_i[warp] = imin; }
    __syncthreads();
    mx = red_f[0]; cmin = red_f[8]; imin = red_i[0];
    for (int w = 1; w < nw; ++w) {
        mx = fmaxf(mx, red_f[w]);
        const float oc = red_f[8 + w];
        const int oi = red_i[w];
        if (oc < cmin || (oc == cmin && oi < imin)) { cmin = oc; imin = oi; }
    }
    __syncthreads();
    // ---- Z = sum exp(a - m) (fixed order: per-thread strided, warp tree, then warps 0..nw-1) ------------------
    float z = 0.f;
    for (int i = threadIdx.x; i < ns; i += kPartThreads) {
        const float e = expf(es[i] - mx);
        es[i] = e;
        z += e;
    }
    z = warp_sum(z);
    if (lane == 0) red_f[16 + warp] = z;
    __syncthreads();
    z = 0.f;
    for (int w = 0; w < nw; ++w) z += red_f[16 + w];

    // ---- v = sum_s e_s (x_s - mu) over the column window ---------------------------------------------------------
    const int width = Mw < kPartThreads ? Mw : kPartThreads;      // threads across columns
    const int nsl = kPartThreads / width;                         // sample slices
    const int sl = threadIdx.x / width, tc = threadIdx.x - sl * width;
    const float* xp = x + ((size_t)p * S + s0) * (size_t)H * Dfull;
    const size_t rs = (size_t)H * Dfull;
    if (sl < nsl && Mw <= 4 * width) {
        // up to four columns per thread, all walked together: 16 independent loads in flight per thread instead of 4 (the
        // kernel streams x once and was latency-, not bandwidth-, bound: 1.3 TB/s at 10^5 samples).  Every column still
        // accumulates its samples in the same order with the same operations, so the records are bit-identical.
        float acc[4], m1[4];
        const float* xc[4];
        bool ok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = tc + q * width;
            ok[q] = col < Mw;
            const int cc = ok[q] ? col : 0;
            const int h = cc / Dw, j = cc - h * Dw;
            xc[q] = xp + (size_t)h * Dfull + c0 + j;
            m1[q] = ok[q] ? __ldg(mu + (size_t)p * Mw + cc) : 0.f;
            acc[q] = 0.f;
        }
        int i = sl;
        for (; i + 3 * nsl < ns; i += 4 * nsl) {
            float e[4], v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) e[u] = es[i + u * nsl];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    v[u][q] = (e[u] != 0.f && ok[q]) ? __ldg(xc[q] + (size_t)(i + u * nsl) * rs) : m1[q];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = fmaf(e[u], v[u][q] - m1[q], acc[q]);
        }
        for (; i < ns; i += nsl) {
            const float e = es[i];
            if (e != 0.f) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (ok[q]) acc[q] = fmaf(e, __ldg(xc[q] + (size_t)i * rs) - m1[q], acc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (ok[q]) part[(size_t)sl * Mw + tc + q * width] = acc[q];
    } else if (sl < nsl) {
        for (int col = tc; col < Mw; col += width) {
            const int h = col / Dw, j = col - h * Dw;
            const float* xc = xp + (size_t)h * Dfull + c0 + j;
            const float m1 = __ldg(mu + (size_t)p * Mw + col);
            float acc = 0.f;
            int i = sl;
            for (; i + 3 * nsl < ns; i += 4 * nsl) {
                const float e0 = es[i], e1 = es[i + nsl], e2 = es[i + 2 * nsl], e3 = es[i + 3 * nsl];
                const float v0 = (e0 != 0.f) ? __ldg(xc + (size_t)i * rs) : m1;
                const float v1 = (e1 != 0.f) ? __ldg(xc + (size_t)(i + nsl) * rs) : m1;
                const float v2 = (e2 != 0.f) ? __ldg(xc + (size_t)(i + 2 * nsl) * rs) : m1;
                const float v3 = (e3 != 0.f) ? __ldg(xc + (size_t)(i + 3 * nsl) * rs) : m1;
                acc = fmaf(e0, v0 - m1, acc);
                acc = fmaf(e1, v1 - m1, acc);
                acc = fmaf(e2, v2 - m1, acc);
                acc = fmaf(e3, v3 - m1, acc);
            }
            for (; i < ns; i += nsl) {
                const float e = es[i];
                if (e != 0.f) acc = fmaf(e, __ldg(xc + (size_t)i * rs) - m1, acc);
            }
            part[(size_t)sl * Mw + col] = acc;
        }
    }
    __syncthreads();
    float* r = rec + ((size_t)c * P + p) * (size_t)(kRecHead + Mw);
    for (int col = threadIdx.x; col < Mw; col += kPartThreads) {
        float acc = part[col];
        for (int k = 1; k < nsl; ++k) acc += part[(size_t)k * Mw + col];
        r[kRecHead + col] = acc;
    }
    if (threadIdx.x == 0) {
        r[0] = mx;                                   // -inf for an empty chunk: contributes nothing to the merge
        r[1] = z;
        r[2] = cmin;
        long long gi = (long long)imin + sample_offset;
        r[3] = __int_as_float(ns > 0 ? (int)gi : 0x7fffffff);
    }
}

__global__ void __launch_bounds__(256) softmax_combine_kernel(
    const float* __restrict__ rec, int R, float* __restrict__ mu, float* __restrict__ grad, float* __restrict__ lse,
    float* __restrict__ best_cost, int* __restrict__ best_idx, float step, const float* __restrict__ SigmaR, int P, int H,
    int Dw) {
    extern __shared__ __align__(16) float sm[];
    const int Mw = H * Dw, REC = kRecHead + Mw;
    const int p = blockIdx.x;
    float* sc = sm;                      // [R] scale factors e^{m_r - m*}
    float* gs = sm + ((R + 3) & ~3);     // [Mw] merged mean (SigmaR path)
    __shared__ float s_m, s_Z;
    if (threadIdx.x == 0) {
        float m = -CUDART_INF_F, cmin = CUDART_INF_F;
        int imin = 0x7fffffff;
        for (int r = 0; r < R; ++r) {
            const float* q = rec + ((size_t)r * P + p) * REC;
            m = fmaxf(m, q[0]);
            const float oc = q[2];
            const int oi = __float_as_int(q[3]);
            if (oc < cmin || (oc == cmin && oi < imin)) { cmin = oc; imin = oi; }
        }
        float Z = 0.f;
        for (int r = 0; r < R; ++r) {
            const float* q = rec + ((size_t)r * P + p) * REC;
            const float s = (q[0] == -CUDART_INF_F) ? 0.f : expf(q[0] - m);
            sc[r] = s;
            Z = fmaf(s, q[1], Z);
        }
        s_m = m; s_Z = Z;
        if (lse) { lse[2 * p] = m; lse[2 * p + 1] = Z; }
        if (best_cost) best_cost[p] = cmin;
        if (best_idx) best_idx[p] = imin;
    }
    __syncthreads();
    const float Z = s_Z;
    float* mp = mu + (size_t)p * Mw;
    for (int col = threadIdx.x; col < Mw; col += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < R; ++r) {
            const float s = sc[r];
            if (s != 0.f) acc = fmaf(s, __ldg(rec + ((size_t)r * P + p) * REC + kRecHead + col), acc);
        }
        const float g = acc / Z;
        if (grad) grad[(size_t)p * Mw + col] = g;
        if (SigmaR) gs[col] = g; else mp[col] = fmaf(step, g, mp[col]);
    }
    if (SigmaR) {        // STOMP: mu[h,j] += step * sum_k SigmaR[h,k] g[k,j]   (stomp.py:206-211)
        __syncthreads();
        for (int o = threadIdx.x; o < Mw; o += blockDim.x) {
            const int h = o / Dw, j = o - h * Dw;
            const float* srow = SigmaR + (size_t)h * H;
            float acc = 0.f;
            for (int k = 0; k < H; ++k) acc = fmaf(__ldg(srow + k), gs[k * Dw + j], acc);
            mp[o] = fmaf(step, acc, mp[o]);
        }
    }
}

// Multi-CTA combine: same arithmetic in the same order as softmax_combine_kernel (per column the records are folded in
// ascending order with one FMA each; the normaliser is one serial FMA chain), but (i) the record heads are staged and the
// exponentials computed by all threads, (ii) the columns are spread over gridDim.x CTAs with one column per thread and
// (iii) the loads of eight records are in flight per thread.  The serial single-CTA kernel above took 0.27-0.44 ms for the
// 296 records of one problem and grew linearly with the number of ranks of a sample split.
constexpr int kCombThreads = 128;
__global__ void __launch_bounds__(kCombThreads) softmax_combine_cols_kernel(
    const float* __restrict__ rec, int R, float* __restrict__ mu, float* __restrict__ grad, float* __restrict__ lse,
    float* __restrict__ best_cost, int* __restrict__ best_idx, float step, int apply, int P, int Mw) {
    extern __shared__ __align__(16) float sm[];
    const int REC = kRecHead + Mw;
    const int p = blockIdx.y;
    float* sc = sm;                          // [R] e^{m_r - m*}
    float* hm = sm + ((R + 3) & ~3);         // [R] m_r, then reused
    float* hz = hm + ((R + 3) & ~3);         // [R] Z_r
    float* hc = hz + ((R + 3) & ~3);         // [R] cmin_r
    int* hi = reinterpret_cast<int*>(hc + ((R + 3) & ~3));   // [R] argmin_r
    __shared__ float s_m, s_Z;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const float* q = rec + ((size_t)r * P + p) * REC;
        hm[r] = __ldg(q); hz[r] = __ldg(q + 1); hc[r] = __ldg(q + 2); hi[r] = __float_as_int(__ldg(q + 3));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = -CUDART_INF_F, cmin = CUDART_INF_F;
        int imin = 0x7fffffff;
        for (int r = 0; r < R; ++r) {
            m = fmaxf(m, hm[r]);
            const float oc = hc[r];
            const int oi = hi[r];
            if (oc < cmin || (oc == cmin && oi < imin)) { cmin = oc; imin = oi; }
        }
        s_m = m;
        if (blockIdx.x == 0) {
            if (best_cost) best_cost[p] = cmin;
            if (best_idx) best_idx[p] = imin;
        }
    }
    __syncthreads();
    const float m = s_m;
    for (int r = threadIdx.x; r < R; r += blockDim.x) sc[r] = (hm[r] == -CUDART_INF_F) ? 0.f : expf(hm[r] - m);
    __syncthreads();
    if (threadIdx.x == 0) {
        float Z = 0.f;
        for (int r = 0; r < R; ++r) Z = fmaf(sc[r], hz[r], Z);
        s_Z = Z;
        if (blockIdx.x == 0 && lse) { lse[2 * p] = m; lse[2 * p + 1] = Z; }
    }
    __syncthreads();
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= Mw) return;
    const float Z = s_Z;
    const float* base = rec + (size_t)p * REC + kRecHead + col;
    const size_t stride = (size_t)P * REC;
    float acc = 0.f;
    int r = 0;
    for (; r + 8 <= R; r += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(base + (size_t)(r + u) * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float s = sc[r + u];
            if (s != 0.f) acc = fmaf(s, v[u], acc);
        }
    }
    for (; r < R; ++r) {
        const float s = sc[r];
        if (s != 0.f) acc = fmaf(s, __ldg(base + (size_t)r * stride), acc);
    }
    const float g = acc / Z;
    if (grad) grad[(size_t)p * Mw + col] = g;
    if (apply) mu[(size_t)p * Mw + col] = fmaf(step, g, mu[(size_t)p * Mw + col]);
}

// STOMP: mu[h,j] += step * sum_k SigmaR[h,k] g[k,j]   (stomp.py:206-211), after the multi-CTA combine wrote g
__global__ void __launch_bounds__(256) sigma_update_kernel(const float* __restrict__ g, const float* __restrict__ SigmaR,
                                                           float* __restrict__ mu, float step, int H, int Dw) {
    extern __shared__ __align__(16) float gs[];
    const int Mw = H * Dw, p = blockIdx.x;
    for (int i = threadIdx.x; i < Mw; i += blockDim.x) gs[i] = g[(size_t)p * Mw + i];
    __syncthreads();
    float* mp = mu + (size_t)p * Mw;
    for (int o = threadIdx.x; o < Mw; o += blockDim.x) {
        const int h = o / Dw, j = o - h * Dw;
        const float* srow = SigmaR + (size_t)h * H;
        float acc = 0.f;
        for (int k = 0; k < H; ++k) acc = fmaf(__ldg(srow + k), gs[k * Dw + j], acc);
        mp[o] = fmaf(step, acc, mp[o]);
    }
}

__global__ void softmax_weights_kernel(const float* __restrict__ cost, const float* __restrict__ lse,
                                       float* __restrict__ weights, float temp, int P, int S) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)P * S) return;
    const int p = (int)(i / S);
    weights[i] = expf(-__ldg(cost + i) / temp - lse[2 * p]) / lse[2 * p + 1];
}

// Deterministic sum of n floats in fp64: per-CTA partials in a fixed layout, then one thread adds them in order.
__global__ void __launch_bounds__(256) sum_stage1_kernel(const float* __restrict__ v, long long n, double* __restrict__ partial) {
    __shared__ double red[8];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) acc += (double)__ldg(v + i);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}
__global__ void sum_stage2_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += partial[i];
        out[0] = s;
    }
}

}  // namespace mpb

extern "C" int mpb_softmax_record_len(int H, int Dw) { return mpb::kRecHead + H * Dw; }

extern "C" int mpb_softmax_partial(const float* cost, const float* x, const float* mu, float* rec, float temp, int P,
                                   int S, int H, int Dfull, int c0, int Dw, int n_chunks, long long sample_offset,
                                   void* stream) {
    using namespace mpb;
    MPB_REQUIRE(cost && x && mu && rec, "mpb_softmax_partial: null pointer");
    MPB_REQUIRE(P >= 0 && S >= 0 && H >= 1 && Dfull >= 1 && Dw >= 1 && c0 >= 0 && c0 + Dw <= Dfull, "mpb_softmax_partial: bad sizes");
    MPB_REQUIRE(n_chunks >= 1 && n_chunks <= 65535 * 16, "mpb_softmax_partial: bad n_chunks %d", n_chunks);
    MPB_REQUIRE(temp > 0.f, "mpb_softmax_partial: temperature must be positive");
    MPB_REQUIRE(P <= 65535, "mpb_softmax_partial: at most 65535 particles per call");
    if (P == 0) return MPB_OK;
    const int chunk = (S + n_chunks - 1) / n_chunks > 0 ? (S + n_chunks - 1) / n_chunks : 1;
    const int Mw = H * Dw;
    const int width = Mw < kPartThreads ? Mw : kPartThreads;
    const int nsl = kPartThreads / width;
    const size_t smem = ((size_t)((chunk + 3) & ~3) + (size_t)nsl * Mw) * sizeof(float);
    MPB_REQUIRE(smem <= 200 * 1024, "mpb_softmax_partial: chunk of %d samples too large; use more chunks", chunk);
    cudaError_t e = cudaFuncSetAttribute(softmax_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mpb_softmax_partial: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    dim3 grid(n_chunks, P);
    softmax_partial_kernel<<<grid, kPartThreads, smem, static_cast<cudaStream_t>(stream)>>>(cost, x, mu, rec, temp, P, S, H,
                                                                                           Dfull, c0, Dw, chunk, sample_offset);
    return check_launch("mpb_softmax_partial");
}

extern "C" int mpb_softmax_combine(const float* rec, int R, float* mu, float* grad, float* lse, float* best_cost,
                                   int32_t* best_idx, float step, const float* SigmaR, int P, int H, int Dw, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(rec && mu, "mpb_softmax_combine: null pointer");
    MPB_REQUIRE(R >= 1 && P >= 0 && H >= 1 && Dw >= 1, "mpb_softmax_combine: bad sizes");
    if (P == 0) return MPB_OK;
    // multi-CTA path (identical arithmetic): needs the merged mean in global memory when Sigma_R is applied afterwards
    if ((!SigmaR || grad) && P <= 65535 && (size_t)5 * ((R + 3) & ~3) * sizeof(float) <= 200 * 1024) {
        const int Mw = H * Dw;
        const size_t smem2 = (size_t)5 * ((R + 3) & ~3) * sizeof(float);
        cudaError_t e2 = cudaFuncSetAttribute(softmax_combine_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e2 != cudaSuccess) { set_error("mpb_softmax_combine: %s", cudaGetErrorString(e2)); return MPB_ECUDA; }
        dim3 grid((Mw + kCombThreads - 1) / kCombThreads, P);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        softmax_combine_cols_kernel<<<grid, kCombThreads, smem2, st>>>(rec, R, mu, grad, lse, best_cost, best_idx, step, SigmaR ? 0 : 1, P, Mw);
        if (SigmaR) {
            MPB_REQUIRE((size_t)Mw * sizeof(float) <= 48 * 1024, "mpb_softmax_combine: H * Dw too large for the Sigma_R update");
            sigma_update_kernel<<<P, 256, (size_t)Mw * sizeof(float), st>>>(grad, SigmaR, mu, step, H, Dw);
        }
        return check_launch("mpb_softmax_combine");
    }
    const size_t smem = ((size_t)((R + 3) & ~3) + (SigmaR ? (size_t)H * Dw : 0)) * sizeof(float);
    MPB_REQUIRE(smem <= 200 * 1024, "mpb_softmax_combine: too many records (%d)", R);
    cudaError_t e = cudaFuncSetAttribute(softmax_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mpb_softmax_combine: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    softmax_combine_kernel<<<P, 256, smem, static_cast<cudaStream_t>(stream)>>>(rec, R, mu, grad, lse, best_cost, best_idx, step,
                                                                               SigmaR, P, H, Dw);
    return check_launch("mpb_softmax_combine");
}

extern "C" int mpb_softmax_weights(const float* cost, const float* lse, float* weights, float temp, int P, int S,
                                   void* stream) {
    using namespace mpb;
    MPB_REQUIRE(cost && lse && weights, "mpb_softmax_weights: null pointer");
    MPB_REQUIRE(P >= 0 && S >= 0 && temp > 0.f, "mpb_softmax_weights: bad arguments");
    const long long n = (long long)P * S;
    if (n == 0) return MPB_OK;
    softmax_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cost, lse, weights, temp, P, S);
    return check_launch("mpb_softmax_weights");
}

extern "C" int mpb_sum_f64(const float* v, long long n, double* out, double* scratch, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(v && out && scratch && n >= 0, "mpb_sum_f64: bad arguments");
    const int blocks = 1024;           // scratch holds 1024 doubles
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sum_stage1_kernel<<<blocks, 256, 0, st>>>(v, n, scratch);
    sum_stage2_kernel<<<1, 32, 0, st>>>(scratch, blocks, out);
    return check_launch("mpb_sum_f64");
}
