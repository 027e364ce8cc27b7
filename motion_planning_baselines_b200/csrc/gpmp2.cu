// K5: one GPMP2 Gauss-Newton / Levenberg-Marquardt step without ever materialising A, K or J^T J.
//
// Replaces CostComposite.get_linear_system (mp_baselines/planners/costs/cost_functions.py:107-144) with
// CostGP / CostGoalPrior / CostCollision.get_linear_system (:291-314, :538-554, :191-231),
// FieldFactor.get_error(calc_jacobian=True) (costs/factors/field_factor.py:41-57) and
// GPMP2._get_grad_terms / get_torch_solve('cholesky') / _step / _get_costs (gpmp2.py:308-368,451-452,493-495).
//
// The reference builds dense A [B,rows,N], K [B,rows,rows] and J^T J [B,N,N] (N = H*D) and runs a dense
// Cholesky per trajectory.  J^T J is block-tridiagonal with D x D blocks:
//   diagonal block t    : [t==0] K_s  +  [t<H-1] Phi^T Q^-1 Phi  +  [t>0] Q^-1  +  [t==H-1] K_g
//                         + sum_f (1/sigma_f^2) h_f h_f^T  on the position part (t >= 1)
//                         + delta * I               (plain LM)
//                         | delta * diag(mean_b diag(A^T K A))   (trust region, gpmp2.py:366: a BATCH mean, quirk B10)
//   block (t+1, t)      : -Q^-1 Phi                 (constant)
//   g_t                 : start / GP / goal / collision pieces of A^T K b
// Three kernels:
//   gpmp2_linearize_kernel   thread per (b,t): err_f and h_f = -d err_f / d q  (analytic, collision_grad.cuh)
//   gpmp2_diag_mean_kernel   thread per (t,k): deterministic batch mean of the collision part of diag(A^T K A)
//   gpmp2_solve_kernel       warp per trajectory: block Cholesky (forward sweep), back substitution, update,
//                            cost = b^T K b.  The factorisation runs in fp64 (entries reach 1e10 at the default
//                            sigmas 1e-5, cond ~1e12+): the reference's fp32 dense Cholesky is the less accurate
//                            of the two, so agreement is limited by the reference's own rounding.
#include "collision_grad.cuh"

namespace mpb {

struct LinArgs {
    const float* x;
    int B, H, D, d, M;
    mpb_robot_desc robot;
    RobotLayout rl;
    FieldArgs fields;
    float* err;        // [nf,B,H]
    float* hobs;       // [nf,B,H,d]
    int n_interp;      // extra points per segment (0: none)
    float w[MPB_MAX_INTERP + 1];
};

template <int KIND>
__global__ void __launch_bounds__(128) gpmp2_linearize_kernel(const __grid_constant__ LinArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_fields(a.fields, a.robot, smem);
    if (KIND == MPB_ROBOT_CHAIN) stage_robot(a.robot, a.rl, smem);
    __syncthreads();
    const float point_r = (KIND == MPB_ROBOT_POINT) ? __ldg(a.robot.sphere_r) : 0.f;
    const int d = a.d, nf = a.fields.n_fields;
    const long long n = (long long)a.B * a.H;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % a.H);
        float q[MPB_MAX_DOF];
        const float* xt = a.x + (size_t)i * a.D;
#pragma unroll
        for (int k = 0; k < MPB_MAX_DOF; ++k) q[k] = (k < d) ? __ldg(xt + k) : 0.f;
        for (int f = 0; f < nf; ++f) {
            float g[MPB_MAX_DOF];
            float e = 0.f;
            if (t >= 1) {                               // waypoint 0 carries no collision factor (cost_functions.py:165-169)
                e = waypoint_err_grad<KIND>(smem, a.fields.l[f], a.rl, a.robot.ws_dim, point_r, q, d, g);
                // interpolated collision checking: the Jacobian row of waypoint t collects the gradients of the
                // up-sampled points of both adjacent segments, weighted by d p / d q_t (field_factor.py:44-57)
#pragma unroll 1
                for (int k = 1; k <= a.n_interp; ++k) {
                    const float wk = a.w[k], uk = __fsub_rn(1.f, wk);
                    float qi[MPB_MAX_DOF], gi[MPB_MAX_DOF];
                    if (t < a.H - 1) {                  // segment t -> t+1: p = q_t (1-w) + q_{t+1} w
#pragma unroll
                        for (int j = 0; j < MPB_MAX_DOF; ++j)
                            qi[j] = (j < d) ? __fadd_rn(__fmul_rn(q[j], uk), __fmul_rn(__ldg(xt + a.D + j), wk)) : 0.f;
                        waypoint_err_grad<KIND>(smem, a.fields.l[f], a.rl, a.robot.ws_dim, point_r, qi, d, gi);
#pragma unroll
                        for (int j = 0; j < MPB_MAX_DOF; ++j) g[j] = fmaf(uk, gi[j], g[j]);
                    }
                    // segment t-1 -> t: p = q_{t-1} (1-w) + q_t w
#pragma unroll
                    for (int j = 0; j < MPB_MAX_DOF; ++j)
                        qi[j] = (j < d) ? __fadd_rn(__fmul_rn(__ldg(xt - a.D + j), uk), __fmul_rn(q[j], wk)) : 0.f;
                    waypoint_err_grad<KIND>(smem, a.fields.l[f], a.rl, a.robot.ws_dim, point_r, qi, d, gi);
#pragma unroll
                    for (int j = 0; j < MPB_MAX_DOF; ++j) g[j] = fmaf(wk, gi[j], g[j]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < MPB_MAX_DOF; ++k) g[k] = 0.f;
            }
            a.err[(size_t)f * n + i] = e;
            float* ho = a.hobs + ((size_t)f * n + i) * d;
#pragma unroll
            for (int k = 0; k < MPB_MAX_DOF; ++k)
                if (k < d) ho[k] = -g[k];               // H_obst = -d err / d q  (field_factor.py:54-57)
        }
    }
}

// dm[t*d + k] = (1/B) sum_b sum_f inv_sigma2_f * hobs[f,b,t,k]^2      (fixed summation order)
// One WARP per entry: lane l folds the trajectories b = l, l + 32, ... in ascending order (four independent loads in
// flight), then the 32 partial sums are combined by a fixed butterfly -- deterministic, and 32 x 4 loads in flight instead
// of the one dependent load per trajectory of a thread-per-entry walk (50 us at B = 1024).
__global__ void __launch_bounds__(128) gpmp2_diag_mean_kernel(const float* __restrict__ hobs, double* __restrict__ dm, int B, int H,
                                                              int d, int nf, float w0, float w1, float w2, float w3) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= H * d) return;
    const float w[4] = {w0, w1, w2, w3};
    const size_t stride = (size_t)H * d;
    double acc = 0.0;
    for (int f = 0; f < nf; ++f) {
        const float* hf = hobs + (size_t)f * B * stride + i;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int b = lane;
        for (; b + 96 < B; b += 128) {
            const float v0 = __ldg(hf + (size_t)b * stride), v1 = __ldg(hf + (size_t)(b + 32) * stride);
            const float v2 = __ldg(hf + (size_t)(b + 64) * stride), v3 = __ldg(hf + (size_t)(b + 96) * stride);
            s0 = fma((double)v0, (double)v0, s0); s1 = fma((double)v1, (double)v1, s1);
            s2 = fma((double)v2, (double)v2, s2); s3 = fma((double)v3, (double)v3, s3);
        }
        for (; b < B; b += 32) {
            const double h = (double)__ldg(hf + (size_t)b * stride);
            s0 = fma(h, h, s0);
        }
        const double s = warp_sum((s0 + s1) + (s2 + s3));
        acc = fma((double)w[f], s, acc);
    }
    if (lane == 0) dm[i] = acc / (double)B;
}

struct SolveArgs {
    float* x;                 // [B,H,D] updated in place
    int B, H, D, d, nf;
    mpb_gp_desc gp;
    const float* err;         // [nf,B,H]
    const float* hobs;        // [nf,B,H,d]
    float wcoll[MPB_MAX_FIELDS];
    const double* diag_mean;  // [H*d] collision part of mean_b diag(A^T K A), or NULL (plain LM)
    float delta, step;
    double* ws;               // [B, H, 2*D*D] factor blocks (C_t | W_{t+1})
    float* cost;              // [B] b^T K b, or NULL
    float* dtheta;            // [B,H,D] or NULL
    int warps_per_cta;
};

// Constant part of the 2x2 (x) I_d block structure.  a=q11, b=q12, c=q22, Phi=[[I, dt I],[0, I]].
struct GPConst {
    double a, b, c, dt, ks, kg;
    double ptqp_pp, ptqp_pv, ptqp_vv;    // Phi^T Q^-1 Phi
    double o_pp, o_pv, o_vp, o_vv;       // O = -Q^-1 Phi  (block (t+1,t))
};

__device__ __forceinline__ GPConst make_gpconst(const mpb_gp_desc& g) {
    GPConst c;
    c.a = g.q11; c.b = g.q12; c.c = g.q22; c.dt = g.dt; c.ks = g.k_start; c.kg = g.has_goal ? (double)g.k_goal : 0.0;
    c.ptqp_pp = c.a;
    c.ptqp_pv = c.a * c.dt + c.b;
    c.ptqp_vv = c.a * c.dt * c.dt + 2.0 * c.b * c.dt + c.c;
    c.o_pp = -c.a; c.o_pv = -(c.a * c.dt + c.b);
    c.o_vp = -c.b; c.o_vv = -(c.b * c.dt + c.c);
    return c;
}

// entry (i,j) of the constant part of diagonal block t
__device__ __forceinline__ double diag_const(const GPConst& c, int t, int H, int d, int i, int j) {
    const int bi = i >= d, bj = j >= d, ki = i - bi * d, kj = j - bj * d;
    if (ki != kj) return 0.0;
    double v = 0.0;
    if (t == 0 && bi == bj) v += c.ks;
    if (t < H - 1) v += (bi == 0 && bj == 0) ? c.ptqp_pp : ((bi == 1 && bj == 1) ? c.ptqp_vv : c.ptqp_pv);
    if (t > 0) v += (bi == 0 && bj == 0) ? c.a : ((bi == 1 && bj == 1) ? c.c : c.b);
    if (t == H - 1 && bi == bj) v += c.kg;
    return v;
}

__device__ __forceinline__ double off_const(const GPConst& c, int d, int i, int j) {
    const int bi = i >= d, bj = j >= d, ki = i - bi * d, kj = j - bj * d;
    if (ki != kj) return 0.0;
    return bi == 0 ? (bj == 0 ? c.o_pp : c.o_pv) : (bj == 0 ? c.o_vp : c.o_vv);
}

__global__ void __launch_bounds__(128) gpmp2_solve_kernel(const __grid_constant__ SolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.H, D = a.D, d = a.d, DD = D * D, M = H * D;
    // per-warp shared memory: S[DD] W[DD] Wn[DD] y[M] r[D] gt[D] gn[D] xs[M floats]
    const size_t per_warp = (size_t)(3 * DD + M + 3 * D) * sizeof(double) + (size_t)((M + 1) & ~1) * sizeof(float);
    unsigned char* base = smem_raw + (size_t)warp * per_warp;
    double* S = reinterpret_cast<double*>(base);
    double* W = S + DD;
    double* Wn = W + DD;
    double* y = Wn + DD;
    double* r = y + M;
    double* gt = r + D;         // g_t being assembled
    double* gn = gt + D;        // contribution of GP factor t to g_{t+1}
    float* xs = reinterpret_cast<float*>(gn + D);
    const GPConst gc = make_gpconst(a.gp);
    const size_t nBH = (size_t)a.B * H;

    for (int b = blockIdx.x * a.warps_per_cta + warp; b < a.B; b += gridDim.x * a.warps_per_cta) {
        float* xg = a.x + (size_t)b * M;
        for (int i = lane; i < M; i += 32) xs[i] = xg[i];
        for (int i = lane; i < D; i += 32) gn[i] = 0.0;
        __syncwarp();
        double* wsb = a.ws + (size_t)b * H * 2 * DD;
        double cost_acc = 0.0;                     // lane-partial b^T K b

        for (int t = 0; t < H; ++t) {
            const float* xt = xs + t * D;
            // ---- g_t ------------------------------------------------------------------------------
            for (int i = lane; i < D; i += 32) {
                double g = gn[i];                  // -Q^-1 e_{t-1}, prepared by the previous step
                if (t == 0) {
                    const double e0 = (double)(__ldg(a.gp.start_state + i) - xt[i]);
                    g += gc.ks * e0;
                    cost_acc += gc.ks * e0 * e0;
                }
                if (t == H - 1 && a.gp.has_goal) {
                    const double eg = (double)(__ldg(a.gp.goal_state + i) - xt[i]);
                    g += gc.kg * eg;
                    cost_acc += gc.kg * eg * eg;
                }
                gt[i] = g;
            }
            __syncwarp();
            if (t < H - 1) {
                // GP factor t: e = x_{t+1} - Phi x_t;  g_t += Phi^T Q^-1 e;  g_{t+1} -= Q^-1 e
                for (int k = lane; k < d; k += 32) {
                    const float* xn = xt + D;
                    const double ep = (double)(xn[k] - fmaf(a.gp.dt, xt[d + k], xt[k]));
                    const double ev = (double)(xn[d + k] - xt[d + k]);
                    const double qp = gc.a * ep + gc.b * ev, qv = gc.b * ep + gc.c * ev;
                    gt[k] += qp;
                    gt[d + k] += gc.dt * qp + qv;
                    gn[k] = -qp;
                    gn[d + k] = -qv;
                    cost_acc += ep * qp + ev * qv;
                }
            }
            __syncwarp();
            // collision rows of waypoint t
            if (t >= 1) {
                for (int f = 0; f < a.nf; ++f) {
                    const double e = (double)__ldg(a.err + (size_t)f * nBH + (size_t)b * H + t);
                    const double w = (double)a.wcoll[f];
                    if (lane == 0) cost_acc += w * e * e;
                    if (e != 0.0) {
                        const float* h = a.hobs + ((size_t)f * nBH + (size_t)b * H + t) * d;
                        for (int k = lane; k < d; k += 32) gt[k] += w * (double)__ldg(h + k) * e;
                    }
                }
            }
            __syncwarp();
            // ---- S = diagonal block t - W W^T --------------------------------------------------------
            for (int idx = lane; idx < DD; idx += 32) {
                const int i = idx / D, j = idx - i * D;
                if (j > i) continue;
                double v = diag_const(gc, t, H, d, i, j);
                if (t >= 1 && i < d) {
                    for (int f = 0; f < a.nf; ++f) {
                        const float* h = a.hobs + ((size_t)f * nBH + (size_t)b * H + t) * d;
                        v += (double)a.wcoll[f] * (double)__ldg(h + i) * (double)__ldg(h + j);
                    }
                }
                if (i == j) {
                    if (a.diag_mean) {
                        double dmv = diag_const(gc, t, H, d, i, i);
                        if (i < d) dmv += a.diag_mean[t * d + i];
                        v += (double)a.delta * dmv;
                    } else {
                        v += (double)a.delta;
                    }
                }
                if (t > 0) {
                    const double* wi = W + i * D;
                    const double* wj = W + j * D;
                    double s = 0.0;
                    for (int k = 0; k < D; ++k) s = fma(wi[k], wj[k], s);
                    v -= s;
                }
                S[idx] = v;
            }
            __syncwarp();
            // ---- in-place Cholesky of S (lower) ------------------------------------------------------
            for (int k = 0; k < D; ++k) {
                const double piv = sqrt(S[k * D + k]);
                __syncwarp();
                if (lane == 0) S[k * D + k] = piv;
                for (int i = k + 1 + lane; i < D; i += 32) S[i * D + k] /= piv;
                __syncwarp();
                const int n = D - k - 1;
                for (int idx = lane; idx < n * n; idx += 32) {
                    const int i = k + 1 + idx / n, j = k + 1 + idx % n;
                    if (j <= i) S[i * D + j] -= S[i * D + k] * S[j * D + k];
                }
                __syncwarp();
            }
            // ---- forward substitution: C y_t = g_t - W y_{t-1} --------------------------------------
            for (int i = lane; i < D; i += 32) {
                double v = gt[i];
                if (t > 0) {
                    const double* wi = W + i * D;
                    const double* yp = y + (t - 1) * D;
                    for (int k = 0; k < D; ++k) v -= wi[k] * yp[k];
                }
                r[i] = v;
            }
            __syncwarp();
            for (int k = 0; k < D; ++k) {
                const double yk = r[k] / S[k * D + k];
                __syncwarp();
                if (lane == 0) y[t * D + k] = yk;
                for (int i = k + 1 + lane; i < D; i += 32) r[i] -= S[i * D + k] * yk;
                __syncwarp();
            }
            // ---- W_{t+1} = O C^-T : row i solves C w = O[i,:]^T ---------------------------------------
            if (t < H - 1) {
                for (int i = lane; i < D; i += 32) {
                    double* w = Wn + i * D;
                    for (int j = 0; j < D; ++j) {
                        double v = off_const(gc, d, i, j);
                        for (int k = 0; k < j; ++k) v -= S[j * D + k] * w[k];
                        w[j] = v / S[j * D + j];
                    }
                }
            }
            __syncwarp();
            // ---- keep the factor blocks for the backward sweep ----------------------------------------
            double* wst = wsb + (size_t)t * 2 * DD;
            for (int idx = lane; idx < DD; idx += 32) {
                wst[idx] = S[idx];
                wst[DD + idx] = (t < H - 1) ? Wn[idx] : 0.0;
            }
            __syncwarp();
            double* tmp = W; W = Wn; Wn = tmp;
        }

        // ---- backward sweep: C_t^T dx_t = y_t - W_{t+1}^T dx_{t+1};  x_t += step * dx_t ---------------------
        // dx_{t+1} is kept in gn (fp64); the updated trajectory is written straight to global memory.
        for (int t = H - 1; t >= 0; --t) {
            const double* wst = wsb + (size_t)t * 2 * DD;
            for (int idx = lane; idx < DD; idx += 32) {
                S[idx] = wst[idx];
                if (t < H - 1) Wn[idx] = wst[DD + idx];
            }
            __syncwarp();
            for (int i = lane; i < D; i += 32) {
                double v = y[t * D + i];
                if (t < H - 1)
                    for (int k = 0; k < D; ++k) v -= Wn[k * D + i] * gn[k];
                r[i] = v;
            }
            __syncwarp();
            for (int k = D - 1; k >= 0; --k) {
                const double xk = r[k] / S[k * D + k];
                __syncwarp();
                if (lane == 0) gt[k] = xk;
                for (int i = lane; i < k; i += 32) r[i] -= S[k * D + i] * xk;
                __syncwarp();
            }
            for (int i = lane; i < D; i += 32) {
                const double dx = gt[i];
                gn[i] = dx;
                if (a.dtheta) a.dtheta[(size_t)b * M + t * D + i] = (float)dx;
                xg[t * D + i] = xs[t * D + i] + a.step * (float)dx;
            }
            __syncwarp();
        }
        if (a.cost) {
            const double c = warp_sum(cost_acc);
            if (lane == 0) a.cost[b] = (float)c;
        }
        __syncwarp();
    }
}

// Small state dimensions (point robots: D = 4 or 6): one THREAD per trajectory, every block in registers, all loops
// unrolled.  The warp-cooperative kernel above is laid out for D up to 16 (the Panda); at D = 4 most of its lanes idle
// behind __syncwarp()s and index arithmetic (137 M warp instructions for 1024 trajectories, 0.485 ms at C2).  Same
// formulas in fp64, same elimination order.  Workspace layout (private to this kernel): per (t, slot) a contiguous
// run over the batch -- C_t lower triangle | W_{t+1} | y_t in 2 D^2 slots -- so that the threads of a warp read and
// write consecutive doubles.
template <int D>
__global__ void __launch_bounds__(32) gpmp2_solve_small_kernel(const __grid_constant__ SolveArgs a) {
    constexpr int d = D / 2, DD = D * D, SLOTS = 2 * DD, NL = D * (D + 1) / 2;
    static_assert(NL + DD + D <= SLOTS, "workspace slots");
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int H = a.H, M = H * D, B = a.B;
    const GPConst gc = make_gpconst(a.gp);
    const size_t nBH = (size_t)B * H;
    float* xg = a.x + (size_t)b * M;
    double W[D][D], y_prev[D], gn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) { gn[i] = 0.0; y_prev[i] = 0.0; }
    double cost = 0.0;
    float xt[D], xn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xt[i] = xg[i];
    for (int t = 0; t < H; ++t) {
        if (t < H - 1) {
#pragma unroll
            for (int i = 0; i < D; ++i) xn[i] = xg[(t + 1) * D + i];
        }
        // ---- g_t -----------------------------------------------------------------------------------------
        double gt[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double g = gn[i];
            if (t == 0) {
                const double e0 = (double)(__ldg(a.gp.start_state + i) - xt[i]);
                g += gc.ks * e0;
                cost += gc.ks * e0 * e0;
            }
            if (t == H - 1 && a.gp.has_goal) {
                const double eg = (double)(__ldg(a.gp.goal_state + i) - xt[i]);
                g += gc.kg * eg;
                cost += gc.kg * eg * eg;
            }
            gt[i] = g;
        }
        if (t < H - 1) {
#pragma unroll
            for (int k = 0; k < d; ++k) {
                const double ep = (double)(xn[k] - fmaf(a.gp.dt, xt[d + k], xt[k]));
                const double ev = (double)(xn[d + k] - xt[d + k]);
                const double qp = gc.a * ep + gc.b * ev, qv = gc.b * ep + gc.c * ev;
                gt[k] += qp;
                gt[d + k] += gc.dt * qp + qv;
                gn[k] = -qp;
                gn[d + k] = -qv;
                cost += ep * qp + ev * qv;
            }
        }
        double hsum[d][d];                    // sum_f w_f h_f h_f^T (collision part of the diagonal block)
#pragma unroll
        for (int i = 0; i < d; ++i)
#pragma unroll
            for (int j = 0; j < d; ++j) hsum[i][j] = 0.0;
        if (t >= 1) {
            for (int f = 0; f < a.nf; ++f) {
                const double e = (double)__ldg(a.err + (size_t)f * nBH + (size_t)b * H + t);
                const double w = (double)a.wcoll[f];
                cost += w * e * e;
                const float* h = a.hobs + ((size_t)f * nBH + (size_t)b * H + t) * d;
                double hv[d];
#pragma unroll
                for (int k = 0; k < d; ++k) hv[k] = (double)__ldg(h + k);
                if (e != 0.0) {
#pragma unroll
                    for (int k = 0; k < d; ++k) gt[k] += w * hv[k] * e;
                }
#pragma unroll
                for (int i = 0; i < d; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j) hsum[i][j] += w * hv[i] * hv[j];
            }
        }
        // ---- S = diagonal block t - W W^T (lower triangle), in-place Cholesky ------------------------------
        double S[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double v = diag_const(gc, t, H, d, i, j);
                if (i < d) v += hsum[i][j];
                if (i == j) {
                    if (a.diag_mean) {
                        double dmv = diag_const(gc, t, H, d, i, i);
                        if (i < d) dmv += a.diag_mean[t * d + i];
                        v += (double)a.delta * dmv;
                    } else {
                        v += (double)a.delta;
                    }
                }
                if (t > 0) {
                    double s2 = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) s2 = fma(W[i][k], W[j][k], s2);
                    v -= s2;
                }
                S[i][j] = v;
            }
        }
        // the diagonal keeps 1 / C_kk: every later division by a pivot (26 per step) becomes a multiplication -- fp64
        // divisions are ~30 dependent instructions each and were most of this latency-bound kernel
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double inv = rsqrt(S[k][k]);
            S[k][k] = inv;
#pragma unroll
            for (int i = k + 1; i < D; ++i) S[i][k] *= inv;
#pragma unroll
            for (int i = k + 1; i < D; ++i)
#pragma unroll
                for (int j = k + 1; j <= i; ++j) S[i][j] -= S[i][k] * S[j][k];
        }
        // ---- forward substitution: C y_t = g_t - W y_{t-1} ----------------------------------------------------
        double r[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double v = gt[i];
            if (t > 0) {
#pragma unroll
                for (int k = 0; k < D; ++k) v -= W[i][k] * y_prev[k];
            }
            r[i] = v;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double yk = r[k] * S[k][k];
            y_prev[k] = yk;
#pragma unroll
            for (int i = k + 1; i < D; ++i) r[i] -= S[i][k] * yk;
        }
        // ---- W_{t+1} = O C^-T ------------------------------------------------------------------------------
        if (t < H - 1) {
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    double v = off_const(gc, d, i, j);
#pragma unroll
                    for (int k = 0; k < j; ++k) v -= S[j][k] * W[i][k];     // W[i][k < j] already holds the new row
                    W[i][j] = v * S[j][j];
                }
        }
        // ---- keep C_t, W_{t+1}, y_t for the backward sweep ------------------------------------------------------
        double* wst = a.ws + (size_t)t * SLOTS * B + b;
        int slot = 0;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) wst[(size_t)(slot++) * B] = S[i][j];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) wst[(size_t)(slot++) * B] = (t < H - 1) ? W[i][j] : 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) wst[(size_t)(slot++) * B] = y_prev[i];
#pragma unroll
        for (int i = 0; i < D; ++i) xt[i] = xn[i];
    }
    // ---- backward sweep: C_t^T dx_t = y_t - W_{t+1}^T dx_{t+1};  x_t += step * dx_t -----------------------------------
    double dxn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) dxn[i] = 0.0;
    for (int t = H - 1; t >= 0; --t) {
        const double* wst = a.ws + (size_t)t * SLOTS * B + b;
        double S[D][D], Wn[D][D], r[D];
        int slot = 0;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) S[i][j] = wst[(size_t)(slot++) * B];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) Wn[i][j] = wst[(size_t)(slot++) * B];
#pragma unroll
        for (int i = 0; i < D; ++i) r[i] = wst[(size_t)(slot++) * B];
        if (t < H - 1) {
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int k = 0; k < D; ++k) r[i] -= Wn[k][i] * dxn[k];
        }
#pragma unroll
        for (int k = D - 1; k >= 0; --k) {
            const double xk = r[k] * S[k][k];
            dxn[k] = xk;
#pragma unroll
            for (int i = 0; i < k; ++i) r[i] -= S[k][i] * xk;
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            if (a.dtheta) a.dtheta[(size_t)b * M + t * D + i] = (float)dxn[i];
            xg[t * D + i] = xg[t * D + i] + a.step * (float)dxn[i];
        }
    }
    if (a.cost) a.cost[b] = (float)cost;
}

static size_t solve_smem_per_warp(int H, int D) {
    const int DD = D * D, M = H * D;
    return (size_t)(3 * DD + M + 3 * D) * sizeof(double) + (size_t)((M + 1) & ~1) * sizeof(float);
}

}  // namespace mpb

extern "C" long long mpb_gpmp2_workspace_bytes(int B, int H, int D) {
    if (B < 0 || H < 0 || D < 0) return -1;
    return (long long)B * H * 2 * D * D * (long long)sizeof(double);
}

extern "C" int mpb_gpmp2_linearize(const float* x, int B, int H, const mpb_robot_desc* robot,
                                   const mpb_field_desc* fields, int n_fields, float* err, float* hobs,
                                   double* diag_mean, void* stream) {
    return mpb_gpmp2_linearize_ex(x, B, H, robot, fields, n_fields, err, hobs, diag_mean, 0, nullptr, stream);
}

extern "C" int mpb_gpmp2_linearize_ex(const float* x, int B, int H, const mpb_robot_desc* robot,
                                      const mpb_field_desc* fields, int n_fields, float* err, float* hobs,
                                      double* diag_mean, int n_interp, const float* interp_w_host, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(n_interp >= 0 && n_interp <= MPB_MAX_INTERP && (n_interp == 0 || interp_w_host),
                "mpb_gpmp2_linearize: n_interp must be in [0,%d] with a weight array", MPB_MAX_INTERP);
    MPB_REQUIRE(B >= 0, "mpb_gpmp2_linearize: negative batch size");
    if (B == 0 || n_fields == 0) return MPB_OK;
    MPB_REQUIRE(x && robot && fields && err && hobs, "mpb_gpmp2_linearize: null pointer");
    MPB_REQUIRE(H >= 2 && n_fields >= 0 && n_fields <= MPB_MAX_FIELDS, "mpb_gpmp2_linearize: bad H / n_fields");
    MPB_REQUIRE(robot->kind == MPB_ROBOT_POINT || robot->kind == MPB_ROBOT_CHAIN, "mpb_gpmp2_linearize: unknown robot kind");
    MPB_REQUIRE(robot->q_dim >= 1 && robot->q_dim <= MPB_MAX_DOF, "mpb_gpmp2_linearize: q_dim out of range");
    if (robot->kind == MPB_ROBOT_POINT)
        MPB_REQUIRE(robot->q_dim == robot->ws_dim && robot->sphere_r, "mpb_gpmp2_linearize: point robot needs q_dim == ws_dim");
    else
        MPB_REQUIRE(robot->ws_dim == 3 && robot->fixed_tf && robot->sphere_link && robot->sphere_off && robot->sphere_r,
                    "mpb_gpmp2_linearize: chain robot needs ws_dim 3 and a sphere table");
    LinArgs a{};
    a.x = x; a.B = B; a.H = H; a.d = robot->q_dim; a.D = 2 * a.d; a.M = H * a.D;
    a.robot = *robot;
    a.fields.n_fields = n_fields;
    {
        const char* why = validate_fields(fields, n_fields, *robot);
        MPB_REQUIRE(!why, "mpb_gpmp2_linearize: %s", why);
    }
    for (int i = 0; i < n_fields; ++i) a.fields.f[i] = fields[i];
    a.err = err; a.hobs = hobs;
    a.n_interp = n_interp;
    for (int k = 0; k <= n_interp && n_interp > 0; ++k) {
        MPB_REQUIRE(interp_w_host[k] >= 0.f && interp_w_host[k] < 1.f, "mpb_gpmp2_linearize: interpolation weights must lie in [0,1)");
        a.w[k] = interp_w_host[k];
    }
    unsigned off = layout_fields(a.fields, 0);
    off = layout_robot(*robot, a.rl, off);
    const size_t smem = off;
    MPB_REQUIRE(smem <= 227 * 1024, "mpb_gpmp2_linearize: too many primitives for shared memory");
    const long long n = (long long)B * H;
    const long long blocks = (n + 127) / 128;
    const int grid = (int)(blocks < (long long)sm_count() * 16 ? blocks : (long long)sm_count() * 16);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (robot->kind == MPB_ROBOT_POINT) {
        e = cudaFuncSetAttribute(gpmp2_linearize_kernel<MPB_ROBOT_POINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) gpmp2_linearize_kernel<MPB_ROBOT_POINT><<<grid, 128, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(gpmp2_linearize_kernel<MPB_ROBOT_CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) gpmp2_linearize_kernel<MPB_ROBOT_CHAIN><<<grid, 128, smem, st>>>(a);
    }
    if (e != cudaSuccess) { set_error("mpb_gpmp2_linearize: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    int rc = check_launch("mpb_gpmp2_linearize");
    if (rc) return rc;
    if (diag_mean) {
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < n_fields; ++i) w[i] = fields[i].inv_sigma2;
        const int nd = H * a.d;
        gpmp2_diag_mean_kernel<<<(nd + 3) / 4, 128, 0, st>>>(hobs, diag_mean, B, H, a.d, n_fields, w[0], w[1], w[2], w[3]);
        rc = check_launch("mpb_gpmp2_linearize(diag_mean)");
    }
    return rc;
}

extern "C" int mpb_gpmp2_solve(float* x, int B, int H, int d, const mpb_gp_desc* gp, const float* err,
                               const float* hobs, const float* inv_sigma2, int n_fields, const double* diag_mean,
                               float delta, float step, double* workspace, float* cost, float* dtheta, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(B >= 0, "mpb_gpmp2_solve: negative batch size");
    if (B == 0) return MPB_OK;
    MPB_REQUIRE(x && gp && workspace, "mpb_gpmp2_solve: null x/gp/workspace");
    MPB_REQUIRE(gp->enabled && gp->start_state && (!gp->has_goal || gp->goal_state), "mpb_gpmp2_solve: the GP prior factors are required");
    MPB_REQUIRE(H >= 2 && d >= 1 && d <= MPB_MAX_DOF, "mpb_gpmp2_solve: bad H / d");
    MPB_REQUIRE(n_fields >= 0 && n_fields <= MPB_MAX_FIELDS && (n_fields == 0 || (err && hobs && inv_sigma2)), "mpb_gpmp2_solve: bad fields");
    SolveArgs a{};
    a.x = x; a.B = B; a.H = H; a.d = d; a.D = 2 * d; a.nf = n_fields;
    a.gp = *gp; a.err = err; a.hobs = hobs;
    for (int i = 0; i < n_fields; ++i) a.wcoll[i] = inv_sigma2[i];
    a.diag_mean = diag_mean; a.delta = delta; a.step = step; a.ws = workspace; a.cost = cost; a.dtheta = dtheta;
    if (a.D == 4 || a.D == 6) {             // point robots: thread per trajectory, blocks in registers
        const int grid_s = (B + 31) / 32;
        if (a.D == 4) gpmp2_solve_small_kernel<4><<<grid_s, 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
        else gpmp2_solve_small_kernel<6><<<grid_s, 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
        return check_launch("mpb_gpmp2_solve");
    }
    const size_t per_warp = solve_smem_per_warp(H, a.D);
    int warps = 4;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    MPB_REQUIRE(per_warp * warps <= 227 * 1024, "mpb_gpmp2_solve: H*D too large for shared memory");
    a.warps_per_cta = warps;
    const size_t smem = per_warp * warps;
    cudaError_t e = cudaFuncSetAttribute(gpmp2_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mpb_gpmp2_solve: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    const int blocks = (B + warps - 1) / warps;
    const int grid = blocks < sm_count() * 8 ? blocks : sm_count() * 8;
    gpmp2_solve_kernel<<<grid, warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("mpb_gpmp2_solve");
}
