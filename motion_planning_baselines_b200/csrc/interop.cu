// Plug-point kernels: the pieces of the duck-typed `robot` / `field` / `cost` contract that the UNMODIFIED reference
// planners call one at a time (SURVEY.md 8b plug points 1 and 2), plus the batched state-collision query of the
// sample-based planners.
//
//   mpb_fk_spheres / _vjp   robot.fk_map_collision(q_pos)           (cost_functions.py:50-52) and its backward
//   mpb_field_cost          field.compute_cost(q_pos, link_pos)     (costs/factors/field_factor.py:39) and d err / d link_pos
//                           (the reference differentiates it by torch.autograd, field_factor.py:52-57, chomp.py:139)
//   mpb_cost_grad           d/dx of CostComposite.eval -- lets `cost(x).sum().backward()` of the reference CHOMP
//                           (chomp.py:134-139) run on the fused cost object
//   mpb_collision_query     task.compute_collision(q) of the RRT planners (rrt_base.py:100-101), batched, and the
//                           per-waypoint flags behind the examples' trajectory statistics
//                           (examples/pointmass_dense_2d_CHOMP.py:130-133)
//
// None of these is on the fused per-iteration path (K1-K6 never materialise link positions); they exist so that a
// reference planner that still runs its own Python loop can swap its robot / field / cost objects for ours.
// Mapping: one thread per configuration, robot and field tables staged in shared memory; signed distances from
// exact_sdf (collision.cuh), i.e. the same separately rounded operations as the cost kernel's exact pass.
#include "collision_grad.cuh"

namespace mpb {

struct FkArgs {
    const float* q;        // [N,d]
    const float* gin;      // [N,Ns,3]   (vjp)
    float* out;            // [N,Ns,3]   (forward)
    float* gq;             // [N,d]      (vjp)
    long long N;
    mpb_robot_desc robot;
    RobotLayout rl;
};

template <bool VJP>
__global__ void __launch_bounds__(128) fk_spheres_kernel(const __grid_constant__ FkArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_robot(a.robot, a.rl, smem);
    __syncthreads();
    const float4* rsphere = reinterpret_cast<const float4*>(smem + a.rl.sphere);
    const float* rtf = reinterpret_cast<const float*>(smem + a.rl.tf);
    const int* rlend = reinterpret_cast<const int*>(smem + a.rl.link_end);
    const int d = a.robot.q_dim, ns = a.robot.n_spheres;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < a.N; n += (long long)gridDim.x * blockDim.x) {
        const float* qn = a.q + n * d;
        float zx[MPB_MAX_DOF], zy[MPB_MAX_DOF], zz[MPB_MAX_DOF], px[MPB_MAX_DOF], py[MPB_MAX_DOF], pz[MPB_MAX_DOF];
        float Fx[MPB_MAX_DOF], Fy[MPB_MAX_DOF], Fz[MPB_MAX_DOF], Tx[MPB_MAX_DOF], Ty[MPB_MAX_DOF], Tz[MPB_MAX_DOF];
        Frame T;
        frame_identity(T);
        int s_begin = 0;
#pragma unroll
        for (int j = 0; j < MPB_MAX_DOF; ++j) {
            if (VJP) { Fx[j] = Fy[j] = Fz[j] = Tx[j] = Ty[j] = Tz[j] = 0.f; zx[j] = zy[j] = zz[j] = px[j] = py[j] = pz[j] = 0.f; }
            if (j < d) {
                float sn, cs;
                sincosf(__ldg(qn + j), &sn, &cs);
                frame_advance(T, rtf + j * 12, cs, sn);
                if (VJP) { zx[j] = T.r02; zy[j] = T.r12; zz[j] = T.r22; px[j] = T.tx; py[j] = T.ty; pz[j] = T.tz; }
                const int s_end = rlend[j];
                float fx = 0.f, fy = 0.f, fz = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
#pragma unroll 1
                for (int s = s_begin; s < s_end; ++s) {
                    const float4 o = rsphere[s];
                    const float cx = fmaf(T.r00, o.x, fmaf(T.r01, o.y, fmaf(T.r02, o.z, T.tx)));
                    const float cy = fmaf(T.r10, o.x, fmaf(T.r11, o.y, fmaf(T.r12, o.z, T.ty)));
                    const float cz = fmaf(T.r20, o.x, fmaf(T.r21, o.y, fmaf(T.r22, o.z, T.tz)));
                    if (!VJP) {
                        float* o3 = a.out + (n * ns + s) * 3;
                        o3[0] = cx; o3[1] = cy; o3[2] = cz;
                    } else {
                        const float* g3 = a.gin + (n * ns + s) * 3;
                        const float gx = __ldg(g3), gy = __ldg(g3 + 1), gz = __ldg(g3 + 2);
                        fx += gx; fy += gy; fz += gz;                      // force on the link
                        tx += cy * gz - cz * gy;                           // torque c x f about the origin
                        ty += cz * gx - cx * gz;
                        tz += cx * gy - cy * gx;
                    }
                }
                if (VJP) { Fx[j] = fx; Fy[j] = fy; Fz[j] = fz; Tx[j] = tx; Ty[j] = ty; Tz[j] = tz; }
                s_begin = s_end;
            }
        }
        if (VJP) {
            // d c_s / d q_k = z_k x (c_s - p_k) for every joint k at or below the sphere's link
            float aFx = 0.f, aFy = 0.f, aFz = 0.f, aTx = 0.f, aTy = 0.f, aTz = 0.f;
#pragma unroll
            for (int j = MPB_MAX_DOF - 1; j >= 0; --j) {
                if (j < d) {
                    aFx += Fx[j]; aFy += Fy[j]; aFz += Fz[j];
                    aTx += Tx[j]; aTy += Ty[j]; aTz += Tz[j];
                    const float mx = aTx - (py[j] * aFz - pz[j] * aFy);
                    const float my = aTy - (pz[j] * aFx - px[j] * aFz);
                    const float mz = aTz - (px[j] * aFy - py[j] * aFx);
                    a.gq[n * d + j] = zx[j] * mx + zy[j] * my + zz[j] * mz;
                }
            }
        }
    }
}

struct FieldCostArgs {
    const float* link_pos;   // [N,Ns,ws]
    long long N;
    int ns, ws;
    mpb_robot_desc robot;
    FieldArgs fields;        // exactly one field
    float* err;              // [N]
    float* grad;             // [N,Ns,ws] | NULL
};

__global__ void __launch_bounds__(128) field_cost_kernel(const __grid_constant__ FieldCostArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_fields(a.fields, a.robot, smem);
    __syncthreads();
    const FieldLayout& fl = a.fields.l[0];
    const int ns = a.ns, ws = a.ws;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < a.N; n += (long long)gridDim.x * blockDim.x) {
        const float* lp = a.link_pos + n * ns * ws;
        float* g = a.grad ? a.grad + n * ns * ws : nullptr;
        float err = 0.f;
        if (fl.kind == MPB_FIELD_SELF) {
            if (g) for (int i = 0; i < ns * ws; ++i) g[i] = 0.f;
            const ushort2* pr = reinterpret_cast<const ushort2*>(smem + fl.pairs);
            for (int p = 0; p < fl.n_pairs; ++p) {
                const ushort2 ij = pr[p];
                const float* ci = lp + ij.x * 3;
                const float* cj = lp + ij.y * 3;
                const float dx = __fsub_rn(__ldg(ci), __ldg(cj)), dy = __fsub_rn(__ldg(ci + 1), __ldg(cj + 1)),
                            dz = __fsub_rn(__ldg(ci + 2), __ldg(cj + 2));
                const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                const float thr = __fadd_rn(__fadd_rn(__ldg(a.robot.sphere_r + ij.x), __ldg(a.robot.sphere_r + ij.y)), fl.margin);
                if (!(d2 < fmaf(thr * thr, 1.0001f, 1e-6f))) continue;
                const float dist = __fsqrt_rn(d2);
                const float h = __fsub_rn(thr, dist);
                if (h > 0.f) {
                    err = __fadd_rn(err, h);
                    if (g) {
                        const float inv = 1.f / dist;
                        const float ux = dx * inv, uy = dy * inv, uz = dz * inv;
                        g[ij.x * 3] -= ux; g[ij.x * 3 + 1] -= uy; g[ij.x * 3 + 2] -= uz;
                        g[ij.y * 3] += ux; g[ij.y * 3 + 1] += uy; g[ij.y * 3 + 2] += uz;
                    }
                }
            }
        } else {
            for (int s = 0; s < ns; ++s) {
                const float cx = __ldg(lp + s * ws), cy = __ldg(lp + s * ws + 1), cz = (ws == 3) ? __ldg(lp + s * ws + 2) : 0.f;
                const float b = __fadd_rn(__ldg(a.robot.sphere_r + s), fl.margin);
                float gx = 0.f, gy = 0.f, gz = 0.f;
                const float sd = field_sdf_grad(smem, fl, ws, cx, cy, cz, b, &gx, &gy, &gz);
                const float h = __fsub_rn(b, sd);
                const bool on = h > 0.f;
                if (on) err = __fadd_rn(err, h);
                if (g) {
                    g[s * ws] = on ? -gx : 0.f;
                    g[s * ws + 1] = on ? -gy : 0.f;
                    if (ws == 3) g[s * ws + 2] = on ? -gz : 0.f;
                }
            }
        }
        a.err[n] = err;
    }
}

struct QueryArgs {
    const float* q;          // [N,stride] (the first d entries of every row are the joint positions)
    long long N;
    int stride;
    mpb_robot_desc robot;
    RobotLayout rl;
    FieldArgs fields;
    uint8_t* flag;           // [N] | NULL
    float* err;              // [N] | NULL   sum over fields of the raw (unweighted) hinge sums
};

template <int KIND>
__global__ void __launch_bounds__(128) collision_query_kernel(const __grid_constant__ QueryArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_fields(a.fields, a.robot, smem);
    if (KIND == MPB_ROBOT_CHAIN) stage_robot(a.robot, a.rl, smem);
    __syncthreads();
    const float point_r = (KIND == MPB_ROBOT_POINT) ? __ldg(a.robot.sphere_r) : 0.f;
    const int d = a.robot.q_dim;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < a.N; n += (long long)gridDim.x * blockDim.x) {
        float q[MPB_MAX_DOF], g[MPB_MAX_DOF];
#pragma unroll
        for (int k = 0; k < MPB_MAX_DOF; ++k) q[k] = (k < d) ? __ldg(a.q + n * a.stride + k) : 0.f;
        float tot = 0.f;
        bool hit = false;
        for (int f = 0; f < a.fields.n_fields; ++f) {
            const float e = waypoint_err_grad<KIND>(smem, a.fields.l[f], a.rl, a.robot.ws_dim, point_r, q, d, g);
            hit = hit || (e > 0.f);
            tot = __fadd_rn(tot, e);
        }
        if (a.flag) a.flag[n] = hit ? 1 : 0;
        if (a.err) a.err[n] = tot;
    }
}

struct GradArgs {
    const float* x;          // [B,H,D]
    const float* gout;       // [B] | NULL (ones)
    float* grad;             // [B,H,D]
    int B, H, D, d;
    mpb_robot_desc robot;
    RobotLayout rl;
    FieldArgs fields;
    mpb_gp_desc gp;
    mpb_extra_cost_desc ex;
    const float* jl_gsum;    // device scalar | NULL
};

template <int KIND>
__global__ void __launch_bounds__(128) cost_grad_kernel(const __grid_constant__ GradArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_fields(a.fields, a.robot, smem);
    if (KIND == MPB_ROBOT_CHAIN) stage_robot(a.robot, a.rl, smem);
    __syncthreads();
    const float point_r = (KIND == MPB_ROBOT_POINT) ? __ldg(a.robot.sphere_r) : 0.f;
    const int d = a.d, D = a.D, H = a.H, nf = a.fields.n_fields;
    const long long n = (long long)a.B * H;
    const float jl_scale = a.ex.jl_enabled ? a.ex.w_jl * (a.jl_gsum ? __ldg(a.jl_gsum) : 1.f) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % H);
        const long long b = i / H;
        const float go = a.gout ? __ldg(a.gout + b) : 1.f;
        const float* xt = a.x + i * D;
        float q[MPB_MAX_DOF], gp[MPB_MAX_DOF], gv[MPB_MAX_DOF];
#pragma unroll
        for (int k = 0; k < MPB_MAX_DOF; ++k) { q[k] = (k < d) ? __ldg(xt + k) : 0.f; gp[k] = 0.f; gv[k] = 0.f; }
        if (t >= 1) {                                       // waypoint 0 carries no collision term (cost_functions.py:165-169)
            for (int f = 0; f < nf; ++f) {
                float g1[MPB_MAX_DOF];
                waypoint_err_grad<KIND>(smem, a.fields.l[f], a.rl, a.robot.ws_dim, point_r, q, d, g1);
                const float wf = a.fields.l[f].weight * a.fields.l[f].inv_sigma2;
#pragma unroll
                for (int k = 0; k < MPB_MAX_DOF; ++k) gp[k] = fmaf(wf, g1[k], gp[k]);
            }
        }
        // GP factor terms: e_t = x_{t+1} - Phi x_t, cost = sum_t e^T Q^-1 e  ->  with a = q11 e_p + q12 e_v, c = q12 e_p + q22 e_v:
        //   d/dp_t = 2 (a_{t-1} - a_t),  d/dv_t = 2 (c_{t-1} - c_t - dt a_t)
        const float w1 = a.gp.enabled ? a.gp.w_gp : 0.f, w2 = a.ex.gp_traj_enabled ? a.ex.w_gp_traj : 0.f;
        if (w1 != 0.f || w2 != 0.f) {
            const float dt = a.gp.dt;
            const float A11 = w1 * a.gp.q11 + w2 * a.ex.t11, A12 = w1 * a.gp.q12 + w2 * a.ex.t12, A22 = w1 * a.gp.q22 + w2 * a.ex.t22;
#pragma unroll
            for (int k = 0; k < MPB_MAX_DOF; ++k) {
                if (k < d) {
                    const float p0 = q[k], v0 = __ldg(xt + d + k);
                    if (t >= 1) {
                        const float pm = __ldg(xt - D + k), vm = __ldg(xt - D + d + k);
                        const float ep = p0 - fmaf(dt, vm, pm), ev = v0 - vm;
                        gp[k] += 2.f * fmaf(A11, ep, A12 * ev);
                        gv[k] += 2.f * fmaf(A12, ep, A22 * ev);
                    }
                    if (t < H - 1) {
                        const float pn = __ldg(xt + D + k), vn = __ldg(xt + D + d + k);
                        const float ep = pn - fmaf(dt, v0, p0), ev = vn - v0;
                        const float aa = fmaf(A11, ep, A12 * ev), cc = fmaf(A12, ep, A22 * ev);
                        gp[k] -= 2.f * aa;
                        gv[k] -= 2.f * fmaf(dt, aa, cc);
                    }
                }
            }
        }
        if (a.gp.enabled) {
            if (t == 0) {
#pragma unroll
                for (int k = 0; k < MPB_MAX_DOF; ++k) {
                    if (k < d) {
                        gp[k] -= 2.f * a.gp.w_gp * a.gp.k_start * (__ldg(a.gp.start_state + k) - q[k]);
                        gv[k] -= 2.f * a.gp.w_gp * a.gp.k_start * (__ldg(a.gp.start_state + d + k) - __ldg(xt + d + k));
                    }
                }
            }
            if (t == H - 1 && a.gp.has_goal) {
#pragma unroll
                for (int k = 0; k < MPB_MAX_DOF; ++k) {
                    if (k < d) {
                        gp[k] -= 2.f * a.gp.w_goal * a.gp.k_goal * (__ldg(a.gp.goal_state + k) - q[k]);
                        gv[k] -= 2.f * a.gp.w_goal * a.gp.k_goal * (__ldg(a.gp.goal_state + d + k) - __ldg(xt + d + k));
                    }
                }
            }
        }
        float* gr = a.grad + i * D;
#pragma unroll
        for (int k = 0; k < MPB_MAX_DOF; ++k) {
            if (k < d) {
                float gk = go * gp[k];
                if (a.ex.jl_enabled) {      // the joint-limit term is a batch-wide scalar added to every trajectory
                    const float lo = fmaxf(__fsub_rn(__fadd_rn(__ldg(a.ex.q_min + k), a.ex.jl_eps), q[k]), 0.f);
                    const float hi = fmaxf(__fsub_rn(q[k], __fsub_rn(__ldg(a.ex.q_max + k), a.ex.jl_eps)), 0.f);
                    gk = fmaf(jl_scale, 2.f * (hi - lo), gk);
                }
                gr[k] = gk;
                gr[d + k] = go * gv[k];
            }
        }
    }
}

static int check_robot(const mpb_robot_desc* robot, const char* who) {
    MPB_REQUIRE(robot, "%s: null robot", who);
    MPB_REQUIRE(robot->kind == MPB_ROBOT_POINT || robot->kind == MPB_ROBOT_CHAIN, "%s: unknown robot kind", who);
    MPB_REQUIRE(robot->q_dim >= 1 && robot->q_dim <= MPB_MAX_DOF, "%s: q_dim out of range", who);
    if (robot->kind == MPB_ROBOT_POINT)
        MPB_REQUIRE(robot->q_dim == robot->ws_dim && robot->sphere_r, "%s: point robot needs q_dim == ws_dim", who);
    else
        MPB_REQUIRE(robot->ws_dim == 3 && robot->fixed_tf && robot->sphere_link && robot->sphere_off && robot->sphere_r,
                    "%s: chain robot needs ws_dim 3 and a sphere table", who);
    return MPB_OK;
}

static int grid_for(long long n, int block) {
    const long long blocks = (n + block - 1) / block;
    const long long cap = (long long)sm_count() * 16;
    return (int)(blocks < cap ? blocks : cap);
}

template <typename K, typename A>
static int launch(K kernel, const A& a, long long n, size_t smem, void* stream, const char* who) {
    MPB_REQUIRE(smem <= 227 * 1024, "%s: %zu bytes of shared memory needed", who, smem);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return MPB_ECUDA; }
    kernel<<<grid_for(n, 128), 128, smem, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch(who);
}

}  // namespace mpb

extern "C" int mpb_fk_spheres(const float* q, long long N, const mpb_robot_desc* robot, float* link_pos, void* stream) {
    using namespace mpb;
    if (int rc = check_robot(robot, "mpb_fk_spheres")) return rc;
    MPB_REQUIRE(robot->kind == MPB_ROBOT_CHAIN, "mpb_fk_spheres: a point robot's sphere centre IS q (take a view)");
    MPB_REQUIRE(N >= 0, "mpb_fk_spheres: negative N");
    if (N == 0) return MPB_OK;
    MPB_REQUIRE(q && link_pos, "mpb_fk_spheres: null pointer");
    FkArgs a{};
    a.q = q; a.out = link_pos; a.N = N; a.robot = *robot;
    const size_t smem = layout_robot(*robot, a.rl, 0);
    return launch(fk_spheres_kernel<false>, a, N, smem, stream, "mpb_fk_spheres");
}

extern "C" int mpb_fk_spheres_vjp(const float* q, const float* grad_link, long long N, const mpb_robot_desc* robot, float* grad_q,
                                  void* stream) {
    using namespace mpb;
    if (int rc = check_robot(robot, "mpb_fk_spheres_vjp")) return rc;
    MPB_REQUIRE(robot->kind == MPB_ROBOT_CHAIN, "mpb_fk_spheres_vjp: chain robots only");
    MPB_REQUIRE(N >= 0, "mpb_fk_spheres_vjp: negative N");
    if (N == 0) return MPB_OK;
    MPB_REQUIRE(q && grad_link && grad_q, "mpb_fk_spheres_vjp: null pointer");
    FkArgs a{};
    a.q = q; a.gin = grad_link; a.gq = grad_q; a.N = N; a.robot = *robot;
    const size_t smem = layout_robot(*robot, a.rl, 0);
    return launch(fk_spheres_kernel<true>, a, N, smem, stream, "mpb_fk_spheres_vjp");
}

extern "C" int mpb_field_cost(const float* link_pos, long long N, const mpb_robot_desc* robot, const mpb_field_desc* field,
                              float* err, float* grad_link, void* stream) {
    using namespace mpb;
    if (int rc = check_robot(robot, "mpb_field_cost")) return rc;
    MPB_REQUIRE(N >= 0, "mpb_field_cost: negative N");
    if (N == 0) return MPB_OK;
    MPB_REQUIRE(link_pos && field && err, "mpb_field_cost: null pointer");
    FieldCostArgs a{};
    a.link_pos = link_pos; a.N = N; a.ns = robot->n_spheres; a.ws = robot->ws_dim; a.robot = *robot;
    a.err = err; a.grad = grad_link;
    a.fields.n_fields = 1;
    {
        const char* why = validate_fields(field, 1, *robot);
        MPB_REQUIRE(!why, "mpb_field_cost: %s", why);
    }
    a.fields.f[0] = *field;
    const size_t smem = layout_fields(a.fields, 0);
    return launch(field_cost_kernel, a, N, smem, stream, "mpb_field_cost");
}

extern "C" int mpb_collision_query(const float* q, long long N, int row_stride, const mpb_robot_desc* robot,
                                   const mpb_field_desc* fields, int n_fields, uint8_t* in_collision, float* err, void* stream) {
    using namespace mpb;
    if (int rc = check_robot(robot, "mpb_collision_query")) return rc;
    MPB_REQUIRE(N >= 0, "mpb_collision_query: negative N");
    if (N == 0) return MPB_OK;
    MPB_REQUIRE(q && (in_collision || err), "mpb_collision_query: null pointer");
    MPB_REQUIRE(row_stride >= robot->q_dim, "mpb_collision_query: row_stride < q_dim");
    QueryArgs a{};
    a.q = q; a.N = N; a.stride = row_stride; a.robot = *robot; a.flag = in_collision; a.err = err;
    a.fields.n_fields = n_fields;
    {
        const char* why = validate_fields(fields, n_fields, *robot);
        MPB_REQUIRE(!why, "mpb_collision_query: %s", why);
    }
    for (int i = 0; i < n_fields; ++i) a.fields.f[i] = fields[i];
    unsigned off = layout_fields(a.fields, 0);
    off = layout_robot(*robot, a.rl, off);
    return robot->kind == MPB_ROBOT_POINT
               ? launch(collision_query_kernel<MPB_ROBOT_POINT>, a, N, off, stream, "mpb_collision_query")
               : launch(collision_query_kernel<MPB_ROBOT_CHAIN>, a, N, off, stream, "mpb_collision_query");
}

extern "C" int mpb_cost_grad(const float* x, int B, int H, const mpb_robot_desc* robot, const mpb_field_desc* fields,
                             int n_fields, const mpb_gp_desc* gp, const mpb_extra_cost_desc* extra, const float* gout,
                             const float* jl_gsum, float* grad, void* stream) {
    using namespace mpb;
    if (int rc = check_robot(robot, "mpb_cost_grad")) return rc;
    MPB_REQUIRE(B >= 0, "mpb_cost_grad: negative B");
    if (B == 0) return MPB_OK;
    MPB_REQUIRE(x && grad, "mpb_cost_grad: null pointer");
    MPB_REQUIRE(H >= 2, "mpb_cost_grad: need H >= 2 (got %d)", H);
    GradArgs a{};
    a.x = x; a.gout = gout; a.grad = grad; a.B = B; a.H = H; a.d = robot->q_dim; a.D = 2 * a.d; a.robot = *robot;
    a.jl_gsum = jl_gsum;
    if (gp) {
        a.gp = *gp;
        MPB_REQUIRE(!gp->enabled || gp->start_state, "mpb_cost_grad: GP terms need start_state");
        MPB_REQUIRE(!gp->enabled || !gp->has_goal || gp->goal_state, "mpb_cost_grad: has_goal needs goal_state");
    }
    if (extra) {
        a.ex = *extra;
        MPB_REQUIRE(!extra->jl_enabled || (extra->q_min && extra->q_max), "mpb_cost_grad: joint-limit term needs q_min / q_max");
        if (extra->gp_traj_enabled && !(gp && gp->enabled)) {
            MPB_REQUIRE(gp, "mpb_cost_grad: the GP-trajectory term takes dt from gp->dt");
        }
    }
    a.fields.n_fields = n_fields;
    {
        const char* why = validate_fields(fields, n_fields, *robot);
        MPB_REQUIRE(!why, "mpb_cost_grad: %s", why);
    }
    for (int i = 0; i < n_fields; ++i) a.fields.f[i] = fields[i];
    unsigned off = layout_fields(a.fields, 0);
    off = layout_robot(*robot, a.rl, off);
    const long long n = (long long)B * H;
    return robot->kind == MPB_ROBOT_POINT ? launch(cost_grad_kernel<MPB_ROBOT_POINT>, a, n, off, stream, "mpb_cost_grad")
                                          : launch(cost_grad_kernel<MPB_ROBOT_CHAIN>, a, n, off, stream, "mpb_cost_grad");
}
