// K6: MPPI control sampling + rollout + quadratic cost + importance-sampling term.
//
// Replaces, per iteration of MPPI.sample_and_eval (mp_baselines/planners/mppi.py:88-134):
//   ControlTrajectoryGaussian.sample     priors/gaussian.py:276-298   U[n,:,i] = mean[:,i] + L_i eps[i,n,:]
//                                        (python loop over control dims; L_i = cholesky(Cov[:,:,i]))
//   get_state_trajectories_rollout       mppi.py:190-210 + dynamics/point.py:102-140 (velocity control, deterministic):
//                                        x_0 = state, x_{t+1} = x_t + clamp(u_t, ctrl_min, ctrl_max) * dt
//                                        (T-1 sequential torch launches in the reference)
//   PointParticleDynamics.traj_cost      point.py:154-226: sum_t disc_t (w_pos |x_t - g|^2 + w_ctrl |u_t|^2)
//                                        + disc_{T-1} w_posT |x_{T-1} - g|^2   (the w_vel slice is empty for velocity control)
//   IS term                              mppi.py:125-128: temp * U[:, :, i]^T Cov_i^-1 mean[:, i]
//
// The obstacle ("energy") cost of point.py:192-196 is evaluated on the (state | control) rows this kernel writes by
// mpb_cost_eval; mpb_mppi_finalize then adds -- faithfully to quirk B2 -- its SUM OVER THE BATCH to every sample.
//
// Mapping: one warp per control sample; lanes over time for the sampling mat-vec (L_i rows staged in shared memory,
// padded to avoid bank conflicts) and for the cost terms; the rollout itself is a sequential fp32 recurrence per
// state dimension (lane = dimension) so that it rounds exactly like the reference's step-by-step loop.
// Bound: FP32 (T(T+1)C/2 FMAs per sample) / HBM write of the (state | control) rows.
#include "mpb_common.cuh"
#include "philox.cuh"

namespace mpb {

constexpr int kMppiWarps = 8;

struct MppiArgs {
    const float* L;        // [C,T,T] lower factors
    const float* Cov_inv;  // [C,T,T]
    const float* mean;     // [T,C] mean of the IS term (MPPI._mean)
    const float* mean_s;   // [T,C] mean the controls are sampled around (ctrl_dist.loc: differs from `mean` after shift())
    const float* eps;      // [C,N,T], or NULL: drawn in the kernel (layout MPB_NOISE_MPPI)
    NoiseArgs noise;
    const float* state0;   // [sd]
    const float* goal;     // [>=sd]
    const float* ctrl_min; // [C]
    const float* ctrl_max; // [C]
    float* xu;             // [N,T,sd+C]
    float* quad;           // [N]
    float* isv;            // [N,C]
    int N, T, C, sd;
    float dt, discount, w_pos, w_ctrl, w_posT;
    int l_in_smem;
    int l_shared;          // every control dimension has the same factor: only L[0] is staged (caller's promise, mpb_mppi_rollout_opt)
    int tp;                // row pitch of the staged factors in floats (T + 4 when T % 4 == 0: 16-byte rows, else T + 1)
};

__global__ void __launch_bounds__(kMppiWarps * 32) mppi_rollout_kernel(const __grid_constant__ MppiArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int T = a.T, C = a.C, sd = a.sd, W = sd + C, TP = a.tp;
    // shared layout: v[C][T] (Cov_inv_i @ mean_i) | disc[T] | per-warp: es[C][T], us[T][C], xs[T][sd] | L[C][T][T+1] (optional)
    float* vs = sm;
    float* disc = vs + C * T;
    float* wbase = disc + T;
    const int per_warp = C * T + T * C + T * sd;
    float* Ls = wbase + kMppiWarps * per_warp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // v_i = Cov_i^-1 mean_i  (second factor of the IS term), discount sequence, optional L staging
    for (int o = threadIdx.x; o < C * T; o += blockDim.x) {
        const int i = o / T, t = o - i * T;
        const float* row = a.Cov_inv + ((size_t)i * T + t) * T;
        double acc = 0.0;
        for (int k = 0; k < T; ++k) acc = fma((double)__ldg(row + k), (double)__ldg(a.mean + (size_t)k * C + i), acc);
        vs[o] = (float)acc;
    }
    if (threadIdx.x == 0) {
        float dsc = 1.f;                      // cumprod(discount)/discount: 1, d, d^2, ...   (point.py:145-152)
        for (int t = 0; t < T; ++t) { disc[t] = dsc; dsc *= a.discount; }
    }
    if (a.l_in_smem)
        for (int o = threadIdx.x; o < (a.l_shared ? 1 : C) * T * T; o += blockDim.x) {
            const int i = o / (T * T), r = (o / T) % T, k = o % T;
            Ls[((size_t)i * T + r) * TP + k] = __ldg(a.L + o);
        }
    __syncthreads();

    float* es = wbase + warp * per_warp;
    float* us = es + C * T;
    float* xs = us + T * C;
    for (int n = blockIdx.x * kMppiWarps + warp; n < a.N; n += gridDim.x * kMppiWarps) {
        // stage this sample's noise rows (coalesced over time), or draw them: global element (i, s_off+n, k) of [C,N_glob,T]
        if (a.eps) {
            for (int i = 0; i < C; ++i)
                for (int k = lane; k < T; k += 32) es[i * T + k] = __ldg(a.eps + ((size_t)i * a.N + n) * T + k);
        } else if ((T & 3) == 0) {
            const int TQ = T >> 2;
            for (int o = lane; o < C * TQ; o += 32) {
                const int i = o / TQ, kq = o - i * TQ;
                const unsigned long long e = ((unsigned long long)i * a.noise.P_glob + a.noise.s_off + n) * T + 4 * kq;
                const float4 q = philox_normal4(e >> 2, a.noise);
                float* d = es + i * T + 4 * kq;
                d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
            }
        } else {
            for (int o = lane; o < C * T; o += 32) {
                const int i = o / T, k = o - i * T;
                es[o] = philox_normal1(((unsigned long long)i * a.noise.P_glob + a.noise.s_off + n) * T + k, a.noise);
            }
        }
        __syncwarp();
        // U[t,i] = mean[t,i] + sum_{k<=t} L_i[t,k] eps_i[k]
        if (T == 64 && a.l_in_smem && TP == 68) {
            // lane l owns rows l and 63 - l: together 65 factor entries whatever l is, so the triangular work is balanced
            // across the lanes (rows l and l + 32 made the warp run 32 + 64 iterations for 32.5 useful ones on average).
            // 16-byte loads of the factor rows (pitch 68 floats: the eight lanes of a quarter-warp hit 32 distinct banks)
            // and of the noise (one address per chunk: broadcast).  A chunk that crosses the diagonal multiplies the exact
            // zeros torch's scale_tril holds above it: the sum is the ascending-k sum of the generic loop bit for bit.
            const int ta = lane, tb = 63 - lane;
            const int na = (ta >> 2) + 1, nb = (tb >> 2) + 1;           // 1..8 and 9..16 chunks of four
            for (int i = 0; i < C; ++i) {
                const float4* er4 = reinterpret_cast<const float4*>(es + i * T);
                const int il = a.l_shared ? 0 : i;
                const float4* la4 = reinterpret_cast<const float4*>(Ls + ((size_t)il * T + ta) * TP);
                const float4* lb4 = reinterpret_cast<const float4*>(Ls + ((size_t)il * T + tb) * TP);
                float acc_a = 0.f, acc_b = 0.f;
#pragma unroll 4
                for (int c = 0; c < 8; ++c) {
                    const float4 e = er4[c];
                    const float4 lb = lb4[c];
                    acc_b = fmaf(lb.x, e.x, acc_b); acc_b = fmaf(lb.y, e.y, acc_b);
                    acc_b = fmaf(lb.z, e.z, acc_b); acc_b = fmaf(lb.w, e.w, acc_b);
                    if (c < na) {
                        const float4 la = la4[c];
                        acc_a = fmaf(la.x, e.x, acc_a); acc_a = fmaf(la.y, e.y, acc_a);
                        acc_a = fmaf(la.z, e.z, acc_a); acc_a = fmaf(la.w, e.w, acc_a);
                    }
                }
#pragma unroll 4
                for (int c = 8; c < 16; ++c) {
                    if (c < nb) {
                        const float4 e = er4[c];
                        const float4 lb = lb4[c];
                        acc_b = fmaf(lb.x, e.x, acc_b); acc_b = fmaf(lb.y, e.y, acc_b);
                        acc_b = fmaf(lb.z, e.z, acc_b); acc_b = fmaf(lb.w, e.w, acc_b);
                    }
                }
                us[ta * C + i] = __ldg(a.mean_s + (size_t)ta * C + i) + acc_a;
                us[tb * C + i] = __ldg(a.mean_s + (size_t)tb * C + i) + acc_b;
            }
        } else
        for (int t = lane; t < T; t += 32) {
            for (int i = 0; i < C; ++i) {
                float acc = 0.f;
                const float* er = es + i * T;
                if (a.l_in_smem) {
                    const float* lr = Ls + ((size_t)(a.l_shared ? 0 : i) * T + t) * TP;
                    for (int k = 0; k <= t; ++k) acc = fmaf(lr[k], er[k], acc);
                } else {
                    const float* lr = a.L + ((size_t)(a.l_shared ? 0 : i) * T + t) * T;
                    for (int k = 0; k <= t; ++k) acc = fmaf(__ldg(lr + k), er[k], acc);
                }
                us[t * C + i] = __ldg(a.mean_s + (size_t)t * C + i) + acc;
            }
        }
        __syncwarp();
        // rollout: sequential fp32 recurrence per state dimension (velocity control: xdot = clamp(u))
        for (int j = lane; j < sd; j += 32) {
            float xj = __ldg(a.state0 + j);
            const float lo = __ldg(a.ctrl_min + j), hi = __ldg(a.ctrl_max + j);
            xs[j] = xj;
            int t = 0;
            for (; t + 8 < T; t += 8) {             // eight increments fetched before the dependent adds (the stores to xs
                float du[8];                        // would otherwise order every load of us behind them)
#pragma unroll
                for (int q = 0; q < 8; ++q) du[q] = __fmul_rn(fminf(fmaxf(us[(t + q) * C + j], lo), hi), a.dt);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    xj = __fadd_rn(xj, du[q]);
                    xs[(t + q + 1) * sd + j] = xj;
                }
            }
            for (; t + 1 < T; ++t) {
                const float u = fminf(fmaxf(us[t * C + j], lo), hi);
                xj = __fadd_rn(xj, __fmul_rn(u, a.dt));
                xs[(t + 1) * sd + j] = xj;
            }
        }
        __syncwarp();
        // costs + IS dot + write-out, lanes over time
        double pos = 0.0, ctl = 0.0, isd[MPB_MAX_DOF];
#pragma unroll
        for (int i = 0; i < MPB_MAX_DOF; ++i) isd[i] = 0.0;
        float* row = a.xu + (size_t)n * T * W;
        for (int t = lane; t < T; t += 32) {
            float p2 = 0.f, c2 = 0.f;
            for (int j = 0; j < sd; ++j) {
                const float xv = xs[t * sd + j];
                const float dx = xv - __ldg(a.goal + j);
                p2 = fmaf(dx * dx, a.w_pos, p2);
                row[t * W + j] = xv;
            }
#pragma unroll
            for (int i = 0; i < MPB_MAX_DOF; ++i) {
                if (i < C) {
                    const float u = us[t * C + i];
                    c2 = fmaf(u * u, a.w_ctrl, c2);
                    isd[i] = fma((double)u, (double)vs[i * T + t], isd[i]);
                    row[t * W + sd + i] = u;
                }
            }
            pos += (double)(p2 * disc[t]);
            ctl += (double)(c2 * disc[t]);
        }
        pos = warp_sum(pos);
        ctl = warp_sum(ctl);
#pragma unroll
        for (int i = 0; i < MPB_MAX_DOF; ++i)
            if (i < C) isd[i] = warp_sum(isd[i]);
        if (lane == 0) {
            float term = 0.f;
            for (int j = 0; j < sd; ++j) {
                const float dx = xs[(T - 1) * sd + j] - __ldg(a.goal + j);
                term = fmaf(dx * dx, a.w_posT, term);
            }
            term *= disc[T - 1];
            // costs = pos + vel(=0) + ctrl + terminal   (point.py:225; the energy term is added by mpb_mppi_finalize)
            a.quad[n] = ((float)pos + (float)ctl) + term;
#pragma unroll
            for (int i = 0; i < MPB_MAX_DOF; ++i)
                if (i < C) a.isv[(size_t)n * C + i] = (float)isd[i];
        }
        __syncwarp();
    }
}

// cost[n] = (quad[n] + energy) + temp*is[n,0] + temp*is[n,1] + ...   in the reference's order (mppi.py:117-128)
__global__ void mppi_finalize_kernel(const float* __restrict__ quad, const float* __restrict__ isv,
                                     const double* __restrict__ energy, float temp, float* __restrict__ cost, int N, int C) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float c = quad[n];
    if (energy) c += (float)energy[0];
    for (int i = 0; i < C; ++i) c += temp * isv[(size_t)n * C + i];
    cost[n] = c;
}

}  // namespace mpb

extern "C" int mpb_mppi_rollout_ex(const float* L_ctrl, const float* Cov_inv, const float* mean, const float* mean_sample,
                                   const float* eps, const mpb_noise_desc* noise, const float* state0, const float* goal,
                                   const float* ctrl_min, const float* ctrl_max, float* xu, float* quad, float* isv, int N,
                                   int T, int C, int sd, float dt, float discount, float w_pos, float w_ctrl, float w_posT,
                                   void* stream) {
    return mpb_mppi_rollout_opt(L_ctrl, Cov_inv, mean, mean_sample, eps, noise, state0, goal, ctrl_min, ctrl_max, xu, quad, isv, N, T,
                                C, sd, dt, discount, w_pos, w_ctrl, w_posT, 0, stream);
}

extern "C" int mpb_mppi_rollout_opt(const float* L_ctrl, const float* Cov_inv, const float* mean, const float* mean_sample,
                                    const float* eps, const mpb_noise_desc* noise, const float* state0, const float* goal,
                                    const float* ctrl_min, const float* ctrl_max, float* xu, float* quad, float* isv, int N,
                                    int T, int C, int sd, float dt, float discount, float w_pos, float w_ctrl, float w_posT,
                                    int flags, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(N >= 0, "mpb_mppi_rollout: negative N");
    if (N == 0) return MPB_OK;
    MPB_REQUIRE(L_ctrl && Cov_inv && mean && (eps || noise) && state0 && goal && ctrl_min && ctrl_max && xu && quad && isv,
                "mpb_mppi_rollout: null pointer");
    MPB_REQUIRE(T >= 2 && C >= 1 && C <= MPB_MAX_DOF, "mpb_mppi_rollout: need T >= 2 and 1 <= C <= %d", MPB_MAX_DOF);
    MPB_REQUIRE(sd == C, "mpb_mppi_rollout: velocity control needs state_dim == control_dim (got %d, %d)", sd, C);
    MppiArgs a{};
    a.L = L_ctrl; a.Cov_inv = Cov_inv; a.mean = mean; a.mean_s = mean_sample ? mean_sample : mean; a.eps = eps;
    if (!eps) {
        const char* why = noise_args(*noise, N, a.noise);
        MPB_REQUIRE(!why, "mpb_mppi_rollout: %s", why);
    }
    a.state0 = state0; a.goal = goal;
    a.ctrl_min = ctrl_min; a.ctrl_max = ctrl_max; a.xu = xu; a.quad = quad; a.isv = isv;
    a.N = N; a.T = T; a.C = C; a.sd = sd; a.dt = dt; a.discount = discount; a.w_pos = w_pos; a.w_ctrl = w_ctrl; a.w_posT = w_posT;
    const size_t base = (size_t)(C * T + T + kMppiWarps * (C * T + T * C + T * sd)) * sizeof(float);
    a.tp = (T % 4 == 0) ? T + 4 : T + 1;
    a.l_shared = (flags & MPB_MPPI_SHARED_FACTOR) ? 1 : 0;
    const size_t lbytes = (size_t)(a.l_shared ? 1 : C) * T * a.tp * sizeof(float);
    // L in shared memory when two CTAs still fit per SM (or when it fits at all); else it is read through L1
    a.l_in_smem = (base + lbytes <= 200 * 1024) ? 1 : 0;
    const size_t smem = base + (a.l_in_smem ? lbytes : 0);
    MPB_REQUIRE(smem <= 227 * 1024, "mpb_mppi_rollout: T*C too large for shared memory");
    cudaError_t e = cudaFuncSetAttribute(mppi_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mpb_mppi_rollout: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mppi_rollout_kernel, kMppiWarps * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int blocks = (N + kMppiWarps - 1) / kMppiWarps;
    const int cap = sm_count() * per_sm;
    const int grid = blocks < cap ? blocks : cap;
    mppi_rollout_kernel<<<grid, kMppiWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("mpb_mppi_rollout");
}

extern "C" int mpb_mppi_rollout(const float* L_ctrl, const float* Cov_inv, const float* mean, const float* eps,
                                const float* state0, const float* goal, const float* ctrl_min, const float* ctrl_max,
                                float* xu, float* quad, float* isv, int N, int T, int C, int sd, float dt, float discount,
                                float w_pos, float w_ctrl, float w_posT, void* stream) {
    MPB_REQUIRE(eps, "mpb_mppi_rollout: eps is null (use mpb_mppi_rollout_ex for in-kernel noise)");
    return mpb_mppi_rollout_ex(L_ctrl, Cov_inv, mean, mean, eps, nullptr, state0, goal, ctrl_min, ctrl_max, xu, quad, isv, N, T,
                               C, sd, dt, discount, w_pos, w_ctrl, w_posT, stream);
}

extern "C" int mpb_mppi_finalize(const float* quad, const float* isv, const double* energy, float temp, float* cost, int N,
                                 int C, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(N >= 0 && C >= 1, "mpb_mppi_finalize: bad sizes");
    if (N == 0) return MPB_OK;
    MPB_REQUIRE(quad && isv && cost, "mpb_mppi_finalize: null pointer");
    mppi_finalize_kernel<<<(N + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(quad, isv, energy, temp, cost, N, C);
    return check_launch("mpb_mppi_finalize");
}
