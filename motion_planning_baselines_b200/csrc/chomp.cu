// K4: the whole CHOMP optimisation loop in one kernel.
//
// Replaces CHOMP._run_optimization / _eval (mp_baselines/planners/chomp.py:127-169): per iteration
//   J    = sum_p cost(x_p) + P * w_prior * sum_{p,j} x_{p,:,j}^T R x_{p,:,j}        (quirk B1: the P factor)
//   g    = dJ/dx   (reference: autograd through FK + SDF; here: analytic, collision_grad.cuh)
//   g    = clamp(g, +-clip);  g[:, 0] = g[:, -1] = 0;   x -= lr * g
// `cost` is a CostComposite of CostCollision terms: weight_f * (1/sigma_f^2) * sum_{t>=1} err_f(x_t)
// (cost_functions.py:85,171-189).  R is the tridiagonal precision of chomp.py:81-101 ([H,H], only the three
// diagonals are read).
//
// Mapping: one warp per trajectory, lanes over waypoints.  The trajectory lives in shared memory for ALL
// iterations (two rows, ping-pong), so global memory sees one read and one write per optimize() call instead of
// one autograd graph per iteration.  Bound: latency / FP32 (SURVEY.md 8d, C2 working set 1 MiB).
#include "collision_grad.cuh"

namespace mpb {

constexpr int kChompWarps = 4;

struct ChompArgs {
    float* x;
    int P, H, D, d, M;
    mpb_robot_desc robot;
    RobotLayout rl;
    FieldArgs fields;
    const float* R;
    float smooth2;        // 2 * P_global * weight_prior_cost
    float lr, clip;
    int n_iters;
    unsigned rows_off;
    int row_stride;
    mpb_extra_cost_desc ex;   // jl_enabled: CostJointLimits term in the cost
};

template <int KIND>
__global__ void __launch_bounds__(kChompWarps * 32) chomp_kernel(const __grid_constant__ ChompArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_fields(a.fields, a.robot, smem);
    if (KIND == MPB_ROBOT_CHAIN) stage_robot(a.robot, a.rl, smem);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.H, D = a.D, d = a.d, M = a.M, nf = a.fields.n_fields;
    float* xa = reinterpret_cast<float*>(smem + a.rows_off) + (size_t)warp * 2 * a.row_stride;
    float* xb = xa + a.row_stride;
    const float point_r = (KIND == MPB_ROBOT_POINT) ? __ldg(a.robot.sphere_r) : 0.f;

    for (int p = blockIdx.x * kChompWarps + warp; p < a.P; p += gridDim.x * kChompWarps) {
        float* xg = a.x + (size_t)p * M;
        for (int i = lane; i < M; i += 32) xa[i] = xg[i];
        __syncwarp();
        for (int it = 0; it < a.n_iters; ++it) {
            for (int t = lane; t < H; t += 32) {
                const float* xt = xa + t * D;
                float* xo = xb + t * D;
                if (t == 0 || t == H - 1) {                 // gradient zeroed at both ends (chomp.py:143-144)
                    for (int k = 0; k < D; ++k) xo[k] = xt[k];
                    continue;
                }
                float gc[MPB_MAX_DOF];
#pragma unroll
                for (int k = 0; k < MPB_MAX_DOF; ++k) gc[k] = 0.f;
                if (nf > 0) {
                    float q[MPB_MAX_DOF];
#pragma unroll
                    for (int k = 0; k < MPB_MAX_DOF; ++k) q[k] = (k < d) ? xt[k] : 0.f;
                    for (int f = 0; f < nf; ++f) {
                        float g1[MPB_MAX_DOF];
                        waypoint_err_grad<KIND>(smem, a.fields.l[f], a.rl, a.robot.ws_dim, point_r, q, d, g1);
                        const float wf = a.fields.l[f].weight * a.fields.l[f].inv_sigma2;
#pragma unroll
                        for (int k = 0; k < MPB_MAX_DOF; ++k) gc[k] = fmaf(wf, g1[k], gc[k]);
                    }
                }
                if (a.ex.jl_enabled) {      // d/dq of w_jl * sum relu(q_min+eps-q)^2 + relu(q-(q_max-eps))^2
#pragma unroll
                    for (int k = 0; k < MPB_MAX_DOF; ++k) {
                        if (k < d) {
                            const float lo = fmaxf(__fsub_rn(__fadd_rn(__ldg(a.ex.q_min + k), a.ex.jl_eps), xt[k]), 0.f);
                            const float hi = fmaxf(__fsub_rn(xt[k], __fsub_rn(__ldg(a.ex.q_max + k), a.ex.jl_eps)), 0.f);
                            gc[k] = fmaf(a.ex.w_jl, 2.f * (hi - lo), gc[k]);
                        }
                    }
                }
                const double r0 = (double)__ldg(a.R + (size_t)t * H + t - 1), r1 = (double)__ldg(a.R + (size_t)t * H + t),
                             r2 = (double)__ldg(a.R + (size_t)t * H + t + 1);
                // position columns: smoothness + collision
#pragma unroll
                for (int k = 0; k < MPB_MAX_DOF; ++k) {
                    if (k < d) {
                        const float rx = (float)(r0 * (double)xt[k - D] + r1 * (double)xt[k] + r2 * (double)xt[k + D]);
                        float g = fmaf(a.smooth2, rx, gc[k]);
                        g = fminf(fmaxf(g, -a.clip), a.clip);
                        xo[k] = xt[k] - a.lr * g;
                    }
                }
                // velocity columns: smoothness only
                for (int k = d; k < D; ++k) {
                    const float rx = (float)(r0 * (double)xt[k - D] + r1 * (double)xt[k] + r2 * (double)xt[k + D]);
                    float g = a.smooth2 * rx;
                    g = fminf(fmaxf(g, -a.clip), a.clip);
                    xo[k] = xt[k] - a.lr * g;
                }
            }
            __syncwarp();
            float* tmp = xa; xa = xb; xb = tmp;
        }
        for (int i = lane; i < M; i += 32) xg[i] = xa[i];
        __syncwarp();
    }
}

template <int KIND>
static cudaError_t launch_chomp(const ChompArgs& a, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(chomp_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    chomp_kernel<KIND><<<grid, kChompWarps * 32, smem, st>>>(a);
    return cudaSuccess;
}

}  // namespace mpb

extern "C" int mpb_chomp_run(float* x, int P, int H, const mpb_robot_desc* robot, const mpb_field_desc* fields,
                             int n_fields, const float* R, float smooth_scale, float lr, float grad_clip, int n_iters,
                             void* stream) {
    return mpb_chomp_run_ex(x, P, H, robot, fields, n_fields, R, smooth_scale, lr, grad_clip, n_iters, nullptr, stream);
}

extern "C" int mpb_chomp_run_ex(float* x, int P, int H, const mpb_robot_desc* robot, const mpb_field_desc* fields,
                                int n_fields, const float* R, float smooth_scale, float lr, float grad_clip, int n_iters,
                                const mpb_extra_cost_desc* extra, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(!extra || !extra->gp_traj_enabled, "mpb_chomp_run: the GP-trajectory term has no fused CHOMP gradient");
    MPB_REQUIRE(!extra || !extra->jl_enabled || (extra->q_min && extra->q_max), "mpb_chomp_run: joint-limit term needs q_min / q_max");
    MPB_REQUIRE(P >= 0 && n_iters >= 0, "mpb_chomp_run: negative P or n_iters");
    if (P == 0 || n_iters == 0) return MPB_OK;
    MPB_REQUIRE(x && robot && R, "mpb_chomp_run: null x/robot/R");
    MPB_REQUIRE(H >= 3, "mpb_chomp_run: need H >= 3 (got %d)", H);
    MPB_REQUIRE(n_fields >= 0 && n_fields <= MPB_MAX_FIELDS && (n_fields == 0 || fields), "mpb_chomp_run: bad fields");
    MPB_REQUIRE(robot->kind == MPB_ROBOT_POINT || robot->kind == MPB_ROBOT_CHAIN, "mpb_chomp_run: unknown robot kind");
    MPB_REQUIRE(robot->q_dim >= 1 && robot->q_dim <= MPB_MAX_DOF, "mpb_chomp_run: q_dim out of range");
    if (robot->kind == MPB_ROBOT_POINT)
        MPB_REQUIRE(robot->q_dim == robot->ws_dim && robot->sphere_r, "mpb_chomp_run: point robot needs q_dim == ws_dim");
    else
        MPB_REQUIRE(robot->ws_dim == 3 && robot->fixed_tf && robot->sphere_link && robot->sphere_off && robot->sphere_r,
                    "mpb_chomp_run: chain robot needs ws_dim 3 and a sphere table");
    ChompArgs a{};
    a.x = x; a.P = P; a.H = H; a.d = robot->q_dim; a.D = 2 * a.d; a.M = H * a.D;
    a.robot = *robot;
    a.fields.n_fields = n_fields;
    {
        const char* why = validate_fields(fields, n_fields, *robot);
        MPB_REQUIRE(!why, "mpb_chomp_run: %s", why);
    }
    for (int i = 0; i < n_fields; ++i) a.fields.f[i] = fields[i];
    if (extra && extra->jl_enabled) a.ex = *extra;
    a.R = R; a.smooth2 = 2.f * smooth_scale; a.lr = lr; a.clip = grad_clip; a.n_iters = n_iters;
    unsigned off = layout_fields(a.fields, 0);
    off = layout_robot(*robot, a.rl, off);
    a.row_stride = (a.M + 3) & ~3;
    a.rows_off = off;
    off += (unsigned)(kChompWarps * 2 * a.row_stride * sizeof(float));
    const size_t smem = off;
    MPB_REQUIRE(smem <= 227 * 1024, "mpb_chomp_run: %zu bytes of shared memory needed", smem);
    const int blocks = (P + kChompWarps - 1) / kChompWarps;
    const int grid = blocks < sm_count() * 8 ? blocks : sm_count() * 8;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = (robot->kind == MPB_ROBOT_POINT) ? launch_chomp<MPB_ROBOT_POINT>(a, grid, smem, st)
                                                     : launch_chomp<MPB_ROBOT_CHAIN>(a, grid, smem, st);
    if (e != cudaSuccess) { set_error("mpb_chomp_run: %s", cudaGetErrorString(e)); return MPB_ECUDA; }
    return check_launch("mpb_chomp_run");
}
