// K3: importance weights + weighted-mean trajectory update.
//
// Replaces StochGPMP._update_distribution (mp_baselines/planners/stoch_gpmp.py:267-279) and
// STOMP._update_distribution / _calc_sample_weights (mp_baselines/planners/stomp.py:199-220):
//   w = softmax(-cost/temp) over the S samples of a particle
//   g = sum_s w_s (x_s - mu)            (returned as approx_grad)
//   mu += step * g                      (Stoch-GPMP)    or    mu += step * SigmaR @ g   (STOMP)
//
// One CTA per particle.  Phase 1: block-wide max / sum (warp shuffles + one shared-memory hop).
// Phase 2: threads own 4 consecutive trajectory columns and stream the S sample rows with 128-bit
// coalesced loads; rows whose weight underflowed to exactly 0 contribute exactly 0 and are not
// fetched (the update then reads only the samples that matter).  Bound: HBM/L2 read of x.
#include "mpb_common.cuh"

namespace mpb {

constexpr int kUpdThreads = 256;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float r = (lane < nw) ? red[lane] : (is_max ? -CUDART_INF_F : 0.f);
    r = is_max ? warp_max(r) : warp_sum(r);
    return r;
}

__global__ void __launch_bounds__(kUpdThreads) softmax_update_kernel(
    const float* __restrict__ cost, const float* __restrict__ x, float* __restrict__ mu, float* __restrict__ weights,
    float* __restrict__ grad, float temp, float step, const float* __restrict__ SigmaR, float* __restrict__ mu_copy, int S, int H,
    int D, int x_dm) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float red[32];
    const int M = H * D;
    float* ws = sm;                    // [S] weights of this particle
    int* nzs = reinterpret_cast<int*>(sm + ((S + 3) & ~3));      // [S] indices of the samples with a non-zero weight, ascending
    float* gs = sm + 2 * ((S + 3) & ~3);   // [M] weighted mean (only when SigmaR)
    __shared__ int s_nnz;
    const int p = blockIdx.x;
    const float* cp = cost + (size_t)p * S;
    pdl_trigger();
    pdl_wait();                        // the costs come from the kernel in front

    // ---- phase 1: softmax over samples ----------------------------------------------------
    float mx = -CUDART_INF_F;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const float a = -__ldg(cp + s) / temp;
        ws[s] = a;
        mx = fmaxf(mx, a);
    }
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const float e = expf(ws[s] - mx);
        ws[s] = e;
        sum += e;
    }
    sum = block_reduce(sum, red, false);
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const float w = ws[s] / sum;
        ws[s] = w;
        weights[(size_t)p * S + s] = w;
    }
    __syncthreads();
    // rows whose weight underflowed to exactly 0 are not fetched: list the others (ascending, so the sums keep their
    // order); at T = 1 the weights are nearly one-hot and walking all S of them per column was most of the kernel
    if (threadIdx.x < 32) {
        int n = 0;
        for (int s0 = 0; s0 < S; s0 += 32) {
            const int s = s0 + (int)threadIdx.x;
            const bool nz = s < S && ws[s] != 0.f;
            const unsigned m = __ballot_sync(MPB_FULL_MASK, nz);
            if (nz) nzs[n + __popc(m & ((1u << threadIdx.x) - 1u))] = s;
            n += __popc(m);
        }
        if (threadIdx.x == 0) s_nnz = n;
    }
    __syncthreads();
    const int nnz = s_nnz;

    // ---- phase 2: g = sum_s w_s (x_s - mu) ------------------------------------------------
    const float* xp = x + (size_t)p * S * M;
    float* mp = mu + (size_t)p * M;
    const bool vec = ((M & 3) == 0);
    if (x_dm) {
        // The sample rows are DOF-MAJOR (column 2H j + 2 h + pv, sample_gp_kron_gen_dm.cu); mu / grad / mu_copy keep the
        // reference layout (column D h + dof pv + j).  A thread still streams four consecutive floats of every listed row
        // with one 128-bit load; they are (h, pos), (h, vel), (h + 1, pos), (h + 1, vel) of one dof.  Per element the same
        // operations in the same sample order as below: bit-identical means.
        const int dof = D >> 1, N2 = 2 * H;
        for (int c = threadIdx.x * 4; c < M; c += blockDim.x * 4) {
            const int j = c / N2, n = c - j * N2;
            const int c0 = (n >> 1) * D + j;
            const int col[4] = {c0, c0 + dof, c0 + D, c0 + D + dof};
            float m4[4], g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i) m4[i] = mp[col[i]];
            for (int k = 0; k < nnz; ++k) {
                const int s = nzs[k];
                const float w = ws[s];
                const float4 v = __ldg(reinterpret_cast<const float4*>(xp + (size_t)s * M + c));
                g[0] = fmaf(w, v.x - m4[0], g[0]); g[1] = fmaf(w, v.y - m4[1], g[1]);
                g[2] = fmaf(w, v.z - m4[2], g[2]); g[3] = fmaf(w, v.w - m4[3], g[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (grad) grad[(size_t)p * M + col[i]] = g[i];
                const float o = fmaf(step, g[i], m4[i]);
                mp[col[i]] = o;
                if (mu_copy) mu_copy[(size_t)p * M + col[i]] = o;
            }
        }
    } else if (vec) {
        for (int c = threadIdx.x * 4; c < M; c += blockDim.x * 4) {
            const float4 m4 = *reinterpret_cast<const float4*>(mp + c);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int k = 0; k < nnz; ++k) {
                const int s = nzs[k];
                const float w = ws[s];
                const float4 v = __ldg(reinterpret_cast<const float4*>(xp + (size_t)s * M + c));
                g.x = fmaf(w, v.x - m4.x, g.x); g.y = fmaf(w, v.y - m4.y, g.y);
                g.z = fmaf(w, v.z - m4.z, g.z); g.w = fmaf(w, v.w - m4.w, g.w);
            }
            if (grad) *reinterpret_cast<float4*>(grad + (size_t)p * M + c) = g;
            if (SigmaR) {
                *reinterpret_cast<float4*>(gs + c) = g;
            } else {
                float4 o;
                o.x = fmaf(step, g.x, m4.x); o.y = fmaf(step, g.y, m4.y);
                o.z = fmaf(step, g.z, m4.z); o.w = fmaf(step, g.w, m4.w);
                *reinterpret_cast<float4*>(mp + c) = o;
                if (mu_copy) *reinterpret_cast<float4*>(mu_copy + (size_t)p * M + c) = o;
            }
        }
    } else {
        for (int c = threadIdx.x; c < M; c += blockDim.x) {
            const float m1 = mp[c];
            float g = 0.f;
            for (int k = 0; k < nnz; ++k) {
                const int s = nzs[k];
                g = fmaf(ws[s], __ldg(xp + (size_t)s * M + c) - m1, g);
            }
            if (grad) grad[(size_t)p * M + c] = g;
            if (SigmaR) gs[c] = g;
            else {
                const float o = fmaf(step, g, m1);
                mp[c] = o;
                if (mu_copy) mu_copy[(size_t)p * M + c] = o;
            }
        }
    }
    if (SigmaR) {       // STOMP: mu[h,j] += step * sum_k SigmaR[h,k] g[k,j]
        __syncthreads();
        for (int o = threadIdx.x; o < M; o += blockDim.x) {
            const int h = o / D, j = o - h * D;
            const float* srow = SigmaR + (size_t)h * H;
            float acc = 0.f;
            for (int k = 0; k < H; ++k) acc = fmaf(__ldg(srow + k), gs[k * D + j], acc);
            const float v = fmaf(step, acc, mp[o]);
            mp[o] = v;
            if (mu_copy) mu_copy[(size_t)p * M + o] = v;
        }
    }
}

}  // namespace mpb

extern "C" int mpb_softmax_update(const float* cost, const float* x, float* mu, float* weights, float* grad,
                                  float temp, float step, const float* SigmaR, int P, int S, int H, int D,
                                  void* stream) {
    return mpb_softmax_update_ex(cost, x, mu, weights, grad, temp, step, SigmaR, nullptr, P, S, H, D, stream);
}

static int softmax_update_impl(const float* cost, const float* x, float* mu, float* weights, float* grad, float temp, float step,
                               const float* SigmaR, float* mu_copy, int P, int S, int H, int D, int x_dm, void* stream);

// the same update reading DOF-MAJOR sample rows (sample_gp_kron_gen_dm.cu); mu / grad / mu_out in the reference layout
extern "C" int mpb_softmax_update_dm(const float* cost, const float* x_dm, float* mu, float* weights, float* grad, float temp,
                                     float step, float* mu_copy, int P, int S, int H, int D, void* stream) {
    MPB_REQUIRE((D & 1) == 0 && ((2 * H) & 3) == 0, "mpb_softmax_update_dm: need an even state dimension and 2H a multiple of 4");
    return softmax_update_impl(cost, x_dm, mu, weights, grad, temp, step, nullptr, mu_copy, P, S, H, D, 1, stream);
}

extern "C" int mpb_softmax_update_ex(const float* cost, const float* x, float* mu, float* weights, float* grad,
                                     float temp, float step, const float* SigmaR, float* mu_copy, int P, int S, int H, int D,
                                     void* stream) {
    return softmax_update_impl(cost, x, mu, weights, grad, temp, step, SigmaR, mu_copy, P, S, H, D, 0, stream);
}

static int softmax_update_impl(const float* cost, const float* x, float* mu, float* weights, float* grad, float temp, float step,
                               const float* SigmaR, float* mu_copy, int P, int S, int H, int D, int x_dm, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(cost && x && mu && weights, "mpb_softmax_update: null pointer");
    MPB_REQUIRE(P >= 0 && S >= 1 && H >= 1 && D >= 1, "mpb_softmax_update: bad sizes");
    MPB_REQUIRE(temp > 0.f, "mpb_softmax_update: temperature must be positive");
    if (P == 0) return MPB_OK;
    const size_t smem = (2 * (size_t)((S + 3) & ~3) + (SigmaR ? (size_t)H * D : 0)) * sizeof(float);
    MPB_REQUIRE(smem <= 200 * 1024, "mpb_softmax_update: S=%d too large for the single-CTA path", S);
    cudaError_t e = cudaFuncSetAttribute(softmax_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mpb_softmax_update: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    e = launch_pdl(softmax_update_kernel, dim3(P), dim3(kUpdThreads), smem, static_cast<cudaStream_t>(stream), cost, x, mu, weights, grad,
                   temp, step, SigmaR, mu_copy, S, H, D, x_dm);
    if (e != cudaSuccess) { set_error("mpb_softmax_update: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    return check_launch("mpb_softmax_update");
}
