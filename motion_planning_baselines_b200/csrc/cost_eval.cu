// K2: fused forward kinematics + collision hinge + GP/start/goal cost (+ importance-sampling dot).
//
// Replaces CostComposite.eval over {CostGP, CostGoalPrior, CostCollision...} and everything below it
// (mp_baselines/planners/costs/cost_functions.py:41-53,70-87,171-189,271-289,523-536;
// costs/factors/field_factor.py:17-39; gp_factor.py:52-56; unary_factor.py:22-29) and the second
// half of the IS term of stoch_gpmp.py:239-241.
//
// Mapping: one warp per trajectory, lanes over waypoints (t = lane, lane+32, ...).  The trajectory
// row [H*D] is staged once in shared memory with coalesced 128-bit loads (the GP factor needs the
// neighbouring waypoint, FK needs the joint vector).  Link frames live in registers; sphere centres
// are produced in register blocks of G and run through the branch-free cull pass against obstacle
// primitives broadcast from shared memory (collision.cuh).  Flagged spheres are compacted into a
// per-warp queue and evaluated exactly 32 at a time.  Per-trajectory sums are reduced with warp
// shuffles; nothing but the [B] cost (and optional terms / flags) is written.
// Bound: FP32 issue rate.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "collision.cuh"

namespace mpb {

constexpr int kWarps = 8;
constexpr int kQCap = 64;            // queue entries per warp (>= 2 * 32)

struct CullTable {       // byte offsets into dynamic shared memory
    unsigned a;          // float4[n_fields * n_spheres]
    unsigned b;          // float[n_fields * n_spheres]
};

struct CostArgs {
    const float* x;
    int B, H, D, d, M;
    mpb_robot_desc robot;
    RobotLayout rl;
    FieldArgs fields;
    mpb_gp_desc gp;
    const float* is_vec;
    int S;
    float is_scale;
    float* cost;
    float* terms;
    unsigned char* free_flag;
    unsigned rows_off, queue_off, list_off;   // byte offsets into dynamic shared memory
    int row_stride;                    // floats
    unsigned* sched;                   // [2] dynamic scheduler: next trajectory index, finished CTAs (self-resetting)
    int list_cap;                      // broad-phase list entries per warp (max primitives of one field)
    mpb_extra_cost_desc ex;            // CostGPTrajectory / CostJointLimits (only read by the XF variant)
    float* jl_out;                     // [B] raw joint-limit term per trajectory
    CullTable ctab;                    // packed kernel: link-frame cull table (cost_eval_packed.cuh)
    int x_dm;                          // packed kernel, 7 dofs, H = 64: the rows of x are dof-major (sample_gp_kron_gen_dm.cu)
    int k2_local;                      // packed kernel: sphere-only lists are culled in the link frame (MPB_K2_LOCAL=0: world frame)
};

// Per-warp queue of flagged spheres: structure of arrays in shared memory at byte offset `base`
// (x[kQCap] y[kQCap] z[kQCap] b[kQCap] f[kQCap]); offsets instead of pointers keep it at two registers.
struct WarpQueue {
    unsigned base;
    int n;
};

struct HingeAcc {
    float h[MPB_MAX_FIELDS];
    bool all_zero;
};

// Exact pass over `count` (<= 32) queue entries starting at `first`, one entry per lane.  Deliberately NOT
// inlined: it is the rare path, and keeping one copy keeps the hot loops resident in the instruction cache.
// It returns the lane's (hinge, field index) and the caller accumulates: accumulators handed over by reference would
// live in local memory (lanes past `count` return (0, field 0); adding +0 is exact).
__device__ __noinline__ float2 drain(const FieldArgs& fa, unsigned qbase, int first, int count, int lane) {
    extern __shared__ __align__(16) unsigned char smem[];
    float2 out = make_float2(0.f, __int_as_float(0));
    if (lane < count) {
        const int i = first + lane;
        const float* qb = reinterpret_cast<const float*>(smem + qbase);
        const int f = __float_as_int(qb[4 * kQCap + i]);
        out = make_float2(exact_hinge(smem, fa.l[f], qb[i], qb[kQCap + i], qb[2 * kQCap + i], qb[3 * kQCap + i]), __int_as_float(f));
    }
    __syncwarp();
    return out;
}

__device__ __forceinline__ void acc_add2(HingeAcc& acc, const float2 hf) {
    const int f = __float_as_int(hf.y);
    acc.all_zero = acc.all_zero && (hf.x == 0.f);
#pragma unroll
    for (int k = 0; k < MPB_MAX_FIELDS; ++k) acc.h[k] += (k == f) ? hf.x : 0.f;
}

// Append the lanes whose `pred` is set; drain a full batch of 32 when available.
__device__ __forceinline__ void enqueue(unsigned char* smem, const FieldArgs& fa, WarpQueue& q, bool pred, float cx,
                                        float cy, float cz, float b, int f, int lane, HingeAcc& acc) {
    const unsigned bal = __ballot_sync(MPB_FULL_MASK, pred);
    if (bal == 0u) return;
    if (pred) {
        float* qb = reinterpret_cast<float*>(smem + q.base) + q.n + __popc(bal & ((1u << lane) - 1u));
        qb[0] = cx; qb[kQCap] = cy; qb[2 * kQCap] = cz; qb[3 * kQCap] = b; qb[4 * kQCap] = __int_as_float(f);
    }
    q.n += __popc(bal);
    __syncwarp();
    if (q.n >= 32) {
        q.n -= 32;
        acc_add2(acc, drain(fa, q.base, q.n, 32, lane));
    }
}

__device__ __forceinline__ void acc_add(HingeAcc& acc, int f, float h) {
    acc.all_zero = acc.all_zero && (h == 0.f);
#pragma unroll
    for (int k = 0; k < MPB_MAX_FIELDS; ++k)
        if (k == f) acc.h[k] += h;
}

template <int G, bool XF>
__device__ __forceinline__ void collide_block(unsigned char* smem, const FieldArgs& fa, WarpQueue& q,
                                              const float (&cx)[G], const float (&cy)[G], const float (&cz)[G],
                                              const float (&rad)[G], bool active, int lane, HingeAcc& acc, int ws_dim) {
    for (int f = 0; f < fa.n_fields; ++f) {
        const FieldLayout& fl = fa.l[f];
        float b[G];
#pragma unroll
        for (int k = 0; k < G; ++k) b[k] = __fadd_rn(rad[k], fl.margin);
        if (XF && fl.kind != MPB_FIELD_PRIMITIVES) {
            if (fl.kind == MPB_FIELD_WORKSPACE && active) {
#pragma unroll
                for (int k = 0; k < G; ++k) acc_add(acc, f, workspace_hinge(fl, ws_dim, cx[k], cy[k], cz[k], b[k]));
            }
            continue;                      // a point robot has no self-collision pairs
        }
        unsigned cand = cull_block<G>(smem, fl, cx, cy, cz, b);
        if (!active) cand = 0u;
        const unsigned any = __reduce_or_sync(MPB_FULL_MASK, cand);
        if (any) {
#pragma unroll
            for (int k = 0; k < G; ++k)
                if (any & (1u << k)) enqueue(smem, fa, q, (cand >> k) & 1u, cx[k], cy[k], cz[k], b[k], f, lane, acc);
        }
    }
}

// Self-collision pass of one waypoint per lane (chain robots; rare path, not inlined).  Link frames are rebuilt by a
// nested chain walk -- Ta after joint a, T continued from Ta up to joint b -- with exactly the operation sequence of
// the main loop, so sphere centres are bit-identical to the ones the other fields see.  A link pair is skipped for the
// whole warp when no lane has the two links' bounding spheres within reach; surviving pairs are evaluated exactly
// (single rounded operations in the oracle's order, oracle/fields.py SelfCollisionField).
__device__ __noinline__ void self_collision_pass(const FieldLayout& fl, const RobotLayout& rl, const float* xt, int d,
                                                 bool active, float& hsum, bool& all_zero) {
    extern __shared__ __align__(16) unsigned char smem[];
    const float4* rsphere = reinterpret_cast<const float4*>(smem + rl.sphere);
    const float4* rbound = reinterpret_cast<const float4*>(smem + rl.bound);
    const float* rtf = reinterpret_cast<const float*>(smem + rl.tf);
    const ushort2* pr = reinterpret_cast<const ushort2*>(smem + fl.pairs);
    const ushort2* grp = reinterpret_cast<const ushort2*>(smem + fl.grp);
    const int* lastb = reinterpret_cast<const int*>(smem + fl.lastb);
    Frame Ta;
    frame_identity(Ta);
#pragma unroll 1
    for (int a = 0; a < d - 1; ++a) {
        frame_advance(Ta, rtf + a * 12, xt[a], xt[d + a]);
        const int b_last = lastb[a];
        if (b_last <= a) continue;
        const float4 ba = rbound[a];
        const float ax = fmaf(Ta.r00, ba.x, fmaf(Ta.r01, ba.y, fmaf(Ta.r02, ba.z, Ta.tx)));
        const float ay = fmaf(Ta.r10, ba.x, fmaf(Ta.r11, ba.y, fmaf(Ta.r12, ba.z, Ta.ty)));
        const float az = fmaf(Ta.r20, ba.x, fmaf(Ta.r21, ba.y, fmaf(Ta.r22, ba.z, Ta.tz)));
        Frame T = Ta;
#pragma unroll 1
        for (int b = a + 1; b <= b_last; ++b) {
            frame_advance(T, rtf + b * 12, xt[b], xt[d + b]);
            const ushort2 g = grp[a * MPB_MAX_DOF + b];
            if (g.x == g.y) continue;
            const float4 bb = rbound[b];
            const float ex = fmaf(T.r00, bb.x, fmaf(T.r01, bb.y, fmaf(T.r02, bb.z, T.tx))) - ax;
            const float ey = fmaf(T.r10, bb.x, fmaf(T.r11, bb.y, fmaf(T.r12, bb.z, T.ty))) - ay;
            const float ez = fmaf(T.r20, bb.x, fmaf(T.r21, bb.y, fmaf(T.r22, bb.z, T.tz))) - az;
            const float reach = ba.w + bb.w + fl.margin;
            const bool near = active && fmaf(ez, ez, fmaf(ey, ey, ex * ex)) < fmaf(reach * reach, 1.001f, 1e-6f);
            if (!__any_sync(MPB_FULL_MASK, near)) continue;
#pragma unroll 1
            for (int p = g.x; p < g.y; ++p) {
                const ushort2 ij = pr[p];
                const float4 oi = rsphere[ij.x], oj = rsphere[ij.y];
                const float cix = fmaf(Ta.r00, oi.x, fmaf(Ta.r01, oi.y, fmaf(Ta.r02, oi.z, Ta.tx)));
                const float ciy = fmaf(Ta.r10, oi.x, fmaf(Ta.r11, oi.y, fmaf(Ta.r12, oi.z, Ta.ty)));
                const float ciz = fmaf(Ta.r20, oi.x, fmaf(Ta.r21, oi.y, fmaf(Ta.r22, oi.z, Ta.tz)));
                const float cjx = fmaf(T.r00, oj.x, fmaf(T.r01, oj.y, fmaf(T.r02, oj.z, T.tx)));
                const float cjy = fmaf(T.r10, oj.x, fmaf(T.r11, oj.y, fmaf(T.r12, oj.z, T.ty)));
                const float cjz = fmaf(T.r20, oj.x, fmaf(T.r21, oj.y, fmaf(T.r22, oj.z, T.tz)));
                const float dx = __fsub_rn(cix, cjx), dy = __fsub_rn(ciy, cjy), dz = __fsub_rn(ciz, cjz);
                const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                const float thr = __fadd_rn(__fadd_rn(oi.w, oj.w), fl.margin);
                if (near && d2 < fmaf(thr * thr, 1.0001f, 1e-6f)) {
                    const float h = fmaxf(__fsub_rn(thr, __fsqrt_rn(d2)), 0.f);
                    hsum += h;
                    all_zero = all_zero && (h == 0.f);
                }
            }
        }
    }
}

// Workspace-boundary hinge of the spheres [s_begin, s_end) of one link (chain robots).  Skipped for the whole warp
// when the lanes' bounding spheres all lie inside the workspace shrunk by their radius + margin.
__device__ __noinline__ void workspace_link_pass(const FieldLayout& fl, const RobotLayout& rl, const Frame& T, float bx,
                                                 float by, float bz, float Rm, int s_begin, int s_end, bool active,
                                                 float& hsum, bool& all_zero) {
    extern __shared__ __align__(16) unsigned char smem[];
    const float slack = fmaf(fabsf(Rm), 1e-3f, 1e-5f) + Rm;
    const float lox = warp_min_f(active ? bx : CUDART_INF_F), hix = warp_max_f(active ? bx : -CUDART_INF_F);
    const float loy = warp_min_f(active ? by : CUDART_INF_F), hiy = warp_max_f(active ? by : -CUDART_INF_F);
    const float loz = warp_min_f(active ? bz : CUDART_INF_F), hiz = warp_max_f(active ? bz : -CUDART_INF_F);
    const bool inside = (lox - slack > fl.lo[0]) && (hix + slack < fl.hi[0]) && (loy - slack > fl.lo[1]) &&
                        (hiy + slack < fl.hi[1]) && (loz - slack > fl.lo[2]) && (hiz + slack < fl.hi[2]);
    if (inside) return;
    const float4* rsphere = reinterpret_cast<const float4*>(smem + rl.sphere);
#pragma unroll 1
    for (int s = s_begin; s < s_end; ++s) {
        const float4 o = rsphere[s];
        const float cx = fmaf(T.r00, o.x, fmaf(T.r01, o.y, fmaf(T.r02, o.z, T.tx)));
        const float cy = fmaf(T.r10, o.x, fmaf(T.r11, o.y, fmaf(T.r12, o.z, T.ty)));
        const float cz = fmaf(T.r20, o.x, fmaf(T.r21, o.y, fmaf(T.r22, o.z, T.tz)));
        const float h = workspace_hinge(fl, 3, cx, cy, cz, __fadd_rn(o.w, fl.margin));
        if (active) {
            hsum += h;
            all_zero = all_zero && (h == 0.f);
        }
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Next trajectory for this warp from the grid-wide counter (one atomic per warp and trajectory).
__device__ __forceinline__ int next_traj(unsigned* ctr, int lane) {
    unsigned v = 0;
    if (lane == 0) v = atomicAdd(ctr, 1u);
    return (int)__shfl_sync(MPB_FULL_MASK, v, 0);
}

// Asynchronous copy of one trajectory row into a warp's staging buffer (16-byte cp.async when aligned).
__device__ __forceinline__ void issue_row(const float* xb, float* dst, int M, bool vec_ok, int lane) {
    if (vec_ok) {
        for (int i = lane; i < (M >> 2); i += 32) cp_async16(dst + 4 * i, xb + 4 * i);
    } else {
#pragma unroll 1
        for (int i = lane; i < M; i += 32) dst[i] = __ldg(xb + i);
    }
    cp_async_commit();
}

template <int KIND, int G, bool XF>
__global__ void __launch_bounds__(kWarps * 32, 3) cost_eval_kernel(const __grid_constant__ CostArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];

    stage_fields(a.fields, a.robot, smem);
    if (KIND == MPB_ROBOT_CHAIN) stage_robot(a.robot, a.rl, smem);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* xs = reinterpret_cast<float*>(smem + a.rows_off) + (size_t)warp * 2 * a.row_stride;   // current row
    float* xnext = xs + a.row_stride;                                                             // row being prefetched
    WarpQueue q;
    q.base = a.queue_off + (unsigned)(warp * kQCap * 5 * sizeof(float));
    q.n = 0;
    const int nf = a.fields.n_fields;
    const int H = a.H, D = a.D, d = a.d, M = a.M;
    const bool vec_ok = ((M & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
    const float4* rsphere = reinterpret_cast<const float4*>(smem + a.rl.sphere);
    const float* rtf = reinterpret_cast<const float*>(smem + a.rl.tf);
    unsigned short* lsph = reinterpret_cast<unsigned short*>(smem + a.list_off) + (size_t)warp * 2 * a.list_cap;
    unsigned short* lbox = lsph + a.list_cap;
    const float point_r = (KIND == MPB_ROBOT_POINT) ? __ldg(a.robot.sphere_r) : 0.f;

    // Persistent warps pull trajectories from a grid-wide counter; the row of the NEXT trajectory streams into the
    // second staging buffer (cp.async) while the current one is evaluated.
    int b = next_traj(a.sched, lane);
    if (b < a.B) issue_row(a.x + (size_t)b * M, xs, M, vec_ok, lane);
    while (b < a.B) {
        const int b_next = next_traj(a.sched, lane);
        cp_async_wait_all();
        __syncwarp();
        if (b_next < a.B) issue_row(a.x + (size_t)b_next * M, xnext, M, vec_ok, lane);

        double acc_gp = 0.0, acc_goal = 0.0, acc_is = 0.0, acc_gpt = 0.0, acc_jl = 0.0;
        HingeAcc hacc;
#pragma unroll
        for (int f = 0; f < MPB_MAX_FIELDS; ++f) hacc.h[f] = 0.f;
        hacc.all_zero = true;
        q.n = 0;
        const float* isv = a.is_vec ? a.is_vec + (size_t)(b / a.S) * M : nullptr;

        for (int t0 = 0; t0 < H; t0 += 32) {
            const int t = t0 + lane;
            const bool valid = t < H;
            const int tc = valid ? t : H - 1;
            float* xt = xs + tc * D;

            // ---- start / GP / goal Mahalanobis terms -------------------------------------
            if (a.gp.enabled && valid) {
                if (t == 0) {
                    float c = 0.f;
#pragma unroll 1
                    for (int k = 0; k < D; ++k) {
                        const float e = __ldg(a.gp.start_state + k) - xt[k];
                        c = fmaf(e * a.gp.k_start, e, c);
                    }
                    acc_gp += (double)c;
                }
                if (t < H - 1) {
                    const float* xn = xt + D;
                    float c = 0.f;
#pragma unroll 1
                    for (int k = 0; k < d; ++k) {
                        const float ep = xn[k] - fmaf(a.gp.dt, xt[d + k], xt[k]);
                        const float ev = xn[d + k] - xt[d + k];
                        c = fmaf(fmaf(a.gp.q11, ep, a.gp.q12 * ev), ep, c);
                        c = fmaf(fmaf(a.gp.q12, ep, a.gp.q22 * ev), ev, c);
                    }
                    acc_gp += (double)c;
                }
                if (t == H - 1 && a.gp.has_goal) {
                    float c = 0.f;
#pragma unroll 1
                    for (int k = 0; k < D; ++k) {
                        const float e = __ldg(a.gp.goal_state + k) - xt[k];
                        c = fmaf(e * a.gp.k_goal, e, c);
                    }
                    acc_goal += (double)c;
                }
            }
            // ---- optional extra terms (before FK overwrites the joint angles) ------------
            if (XF && valid) {
                if (a.ex.gp_traj_enabled && t < H - 1) {          // CostGPTrajectory: GP factors with their own sigma
                    const float* xn = xt + D;
                    float c = 0.f;
#pragma unroll 1
                    for (int k = 0; k < d; ++k) {
                        const float ep = xn[k] - fmaf(a.gp.dt, xt[d + k], xt[k]);
                        const float ev = xn[d + k] - xt[d + k];
                        c = fmaf(fmaf(a.ex.t11, ep, a.ex.t12 * ev), ep, c);
                        c = fmaf(fmaf(a.ex.t12, ep, a.ex.t22 * ev), ev, c);
                    }
                    acc_gpt += (double)c;
                }
                if (a.ex.jl_enabled) {                            // CostJointLimits: squared violation, all waypoints
                    float c = 0.f;
#pragma unroll 1
                    for (int k = 0; k < d; ++k) {
                        const float lo = fmaxf(__fsub_rn(__fadd_rn(__ldg(a.ex.q_min + k), a.ex.jl_eps), xt[k]), 0.f);
                        const float hi = fmaxf(__fsub_rn(xt[k], __fsub_rn(__ldg(a.ex.q_max + k), a.ex.jl_eps)), 0.f);
                        c = fmaf(lo, lo, fmaf(hi, hi, c));
                    }
                    acc_jl += (double)c;
                }
            }
            // ---- importance-sampling dot  x . (Sigma^-1 mu_p) ---------------------------
            if (isv && valid) {
                // terms reach 1e7 and cancel: double-float accumulation (error-free TwoProd / TwoSum in fp32; the FP64
                // pipe is slow on this part), rounded into the fp64 running sum once per waypoint
                const float* yv = isv + t * D;
                float hi = 0.f, lo = 0.f;
#pragma unroll 2
                for (int k = 0; k < D; ++k) {
                    const float xv = xt[k], yk = __ldg(yv + k);
                    const float p = __fmul_rn(xv, yk);
                    const float e = fmaf(xv, yk, -p);                     // xv*yk = p + e exactly
                    const float s = __fadd_rn(hi, p);
                    const float z = __fsub_rn(s, hi);
                    lo = __fadd_rn(lo, __fadd_rn(__fadd_rn(__fsub_rn(hi, __fsub_rn(s, z)), __fsub_rn(p, z)), e));
                    hi = s;
                }
                acc_is += (double)hi + (double)lo;
            }
            // ---- collision (warp-synchronous: every lane takes part, results masked) -------
            if (nf > 0) {
                const bool active = valid && t >= 1;    // waypoint 0 is skipped (cost_functions.py:165-169)
                if (KIND == MPB_ROBOT_POINT) {
                    float cx[1], cy[1], cz[1], rad[1];
                    cx[0] = xt[0];
                    cy[0] = xt[1];
                    cz[0] = (a.robot.ws_dim == 3) ? xt[2] : 0.f;
                    rad[0] = point_r;
                    collide_block<1, XF>(smem, a.fields, q, cx, cy, cz, rad, active, lane, hacc, a.robot.ws_dim);
                } else {
                    // every neighbour has consumed this batch of rows: turn (q | qdot) into (cos q | sin q) in place
                    __syncwarp();
                    if (valid) {
#pragma unroll 1
                        for (int j = 0; j < d; ++j) {
                            float sn, cs;
                            sincosf(xt[j], &sn, &cs);
                            xt[j] = cs;
                            xt[d + j] = sn;
                        }
                    }
                    __syncwarp();
                    Frame T;
                    frame_identity(T);
                    int s_begin = 0;
                    const float4* rbound = reinterpret_cast<const float4*>(smem + a.rl.bound);
                    const int* rlend = reinterpret_cast<const int*>(smem + a.rl.link_end);
#pragma unroll 1
                    for (int j = 0; j < d; ++j) {
                        frame_advance(T, rtf + j * 12, xt[j], xt[d + j]);
                        const int s_end = rlend[j];
                        if (s_end == s_begin) continue;
                        const float4 bs = rbound[j];
                        const float bx = fmaf(T.r00, bs.x, fmaf(T.r01, bs.y, fmaf(T.r02, bs.z, T.tx)));
                        const float by = fmaf(T.r10, bs.x, fmaf(T.r11, bs.y, fmaf(T.r12, bs.z, T.ty)));
                        const float bz = fmaf(T.r20, bs.x, fmaf(T.r21, bs.y, fmaf(T.r22, bs.z, T.tz)));
#pragma unroll 1
                        for (int f = 0; f < nf; ++f) {
                            const FieldLayout& fl = a.fields.l[f];
                            if (XF && fl.kind != MPB_FIELD_PRIMITIVES) {
                                if (fl.kind == MPB_FIELD_WORKSPACE) {
                                    float hs = 0.f;
                                    bool az = true;
                                    workspace_link_pass(fl, a.rl, T, bx, by, bz, bs.w + fl.margin, s_begin, s_end, active, hs, az);
                                    hacc.all_zero = hacc.all_zero && az;
#pragma unroll
                                    for (int k = 0; k < MPB_MAX_FIELDS; ++k)
                                        if (k == f) hacc.h[k] += hs;
                                }
                                continue;
                            }
                            int n_ls, n_lb;
                            __syncwarp();
                            broad_phase(smem, fl, bx, by, bz, bs.w + fl.margin, active, lane, lsph, n_ls, lbox, n_lb);
                            if (n_ls + n_lb == 0) continue;
                            __syncwarp();
#pragma unroll 1
                            for (int s0 = s_begin; s0 < s_end; s0 += G) {
                                float cx[G], cy[G], cz[G], bb[G];
#pragma unroll
                                for (int k = 0; k < G; ++k) {
                                    if (s0 + k < s_end) {
                                        const float4 o = rsphere[s0 + k];
                                        cx[k] = fmaf(T.r00, o.x, fmaf(T.r01, o.y, fmaf(T.r02, o.z, T.tx)));
                                        cy[k] = fmaf(T.r10, o.x, fmaf(T.r11, o.y, fmaf(T.r12, o.z, T.ty)));
                                        cz[k] = fmaf(T.r20, o.x, fmaf(T.r21, o.y, fmaf(T.r22, o.z, T.tz)));
                                        bb[k] = __fadd_rn(o.w, fl.margin);
                                    } else {
                                        cx[k] = cy[k] = cz[k] = 1e18f;     // padding slot: never a candidate
                                        bb[k] = 0.f;
                                    }
                                }
                                unsigned cand = cull_list<G>(smem, fl, lsph, n_ls, lbox, n_lb, cx, cy, cz, bb);
                                if (!active) cand = 0u;
                                const unsigned any = __reduce_or_sync(MPB_FULL_MASK, cand);
                                if (any) {
#pragma unroll
                                    for (int k = 0; k < G; ++k)
                                        if (any & (1u << k))
                                            enqueue(smem, a.fields, q, (cand >> k) & 1u, cx[k], cy[k], cz[k], bb[k], f, lane, hacc);
                                }
                            }
                        }
                        s_begin = s_end;
                    }
                    if (XF) {
#pragma unroll 1
                        for (int f = 0; f < nf; ++f) {
                            const FieldLayout& fl = a.fields.l[f];
                            if (fl.kind != MPB_FIELD_SELF || fl.n_pairs == 0) continue;
                            float hs = 0.f;
                            bool az = true;
                            self_collision_pass(fl, a.rl, xt, d, active, hs, az);
                            hacc.all_zero = hacc.all_zero && az;
#pragma unroll
                            for (int k = 0; k < MPB_MAX_FIELDS; ++k)
                                if (k == f) hacc.h[k] += hs;
                        }
                    }
                }
            }
        }
        if (q.n > 0) {
            acc_add2(hacc, drain(a.fields, q.base, 0, q.n, lane));
            q.n = 0;
        }

        // ---- per-trajectory reductions ------------------------------------------------------
        float total = 0.f;
        int term = 0;
        if (a.gp.enabled) {
            const float c0 = a.gp.w_gp * (float)warp_sum(acc_gp);
            total += c0;
            if (a.terms && lane == 0) a.terms[(size_t)term * a.B + b] = c0;
            ++term;
            if (a.gp.has_goal) {
                const float c1 = a.gp.w_goal * (float)warp_sum(acc_goal);
                total += c1;
                if (a.terms && lane == 0) a.terms[(size_t)term * a.B + b] = c1;
                ++term;
            }
        }
#pragma unroll
        for (int f = 0; f < MPB_MAX_FIELDS; ++f) {
            if (f < nf) {
                const float e = (float)warp_sum((double)hacc.h[f]);
                const float c = a.fields.l[f].weight * (a.fields.l[f].inv_sigma2 * e);
                total += c;
                if (a.terms && lane == 0) a.terms[(size_t)term * a.B + b] = c;
                ++term;
            }
        }
        if (XF && a.ex.gp_traj_enabled) {
            const float c = a.ex.w_gp_traj * (float)warp_sum(acc_gpt);
            total += c;
            if (a.terms && lane == 0) a.terms[(size_t)term * a.B + b] = c;
            ++term;
        }
        if (XF && a.ex.jl_enabled) {
            const float c = (float)warp_sum(acc_jl);
            if (a.jl_out && lane == 0) a.jl_out[b] = c;
        }
        if (isv) total += a.is_scale * (float)warp_sum(acc_is);
        const int all_free = __all_sync(MPB_FULL_MASK, hacc.all_zero);
        if (lane == 0) {
            a.cost[b] = total;
            if (a.free_flag) a.free_flag[b] = (unsigned char)(all_free ? 1 : 0);
        }
        __syncwarp();
        float* t_ = xs; xs = xnext; xnext = t_;
        b = b_next;
    }
    // the last CTA to finish re-arms the scheduler for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(a.sched + 1, 1u);
        if (done == gridDim.x - 1) {
            a.sched[0] = 0u;
            a.sched[1] = 0u;
            __threadfence();
        }
    }
}

template <int KIND, int G, bool XF>
static cudaError_t launch(const CostArgs& a, int blocks_needed, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(cost_eval_kernel<KIND, G, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cost_eval_kernel<KIND, G, XF>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    // persistent grid: exactly one wave of resident CTAs; trajectories are handed out dynamically
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cost_eval_kernel<KIND, G, XF>, kWarps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int resident = sm_count() * per_sm;
    const int grid = blocks_needed < resident ? blocks_needed : resident;
    cost_eval_kernel<KIND, G, XF><<<grid, kWarps * 32, smem, st>>>(a);
    return cudaSuccess;
}

}  // namespace mpb

#include "cost_eval_packed.cuh"

namespace mpb {

template <int DOF, int NW, int MINB, bool BOXES, bool SPH, bool DM = false>
static cudaError_t launch_chain2b(const CostArgs& a, size_t smem, cudaStream_t st) {
    auto kern = cost_eval_chain2_kernel<DOF, NW, MINB, BOXES, SPH, DM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int resident = sm_count() * per_sm, blocks_needed = (a.B + NW - 1) / NW;
    const int grid = blocks_needed < resident ? blocks_needed : resident;
    return launch_pdl(kern, dim3(grid), dim3(NW * 32), smem, st, a);
}

template <int DOF, int NW, int MINB>
static cudaError_t launch_chain2(const CostArgs& a, size_t smem, cudaStream_t st) {
    bool boxes = !a.k2_local;                 // MPB_K2_LOCAL=0 (A/B, tests): the instance with the world-frame cull
    bool spheres = !a.k2_local;
    for (int i = 0; i < a.fields.n_fields; ++i) {
        boxes = boxes || a.fields.f[i].n_boxes > 0;
        spheres = spheres || a.fields.f[i].n_spheres > 0;
    }
    if (!boxes) return launch_chain2b<DOF, NW, MINB, false, true>(a, smem, st);
    return spheres ? launch_chain2b<DOF, NW, MINB, true, true>(a, smem, st) : launch_chain2b<DOF, NW, MINB, true, false>(a, smem, st);
}

// MPB_COST_EVAL=generic forces the one-waypoint-per-lane kernel (A/B timing and the cross-check in the tests).
static bool packed_allowed() {
    const char* v = getenv("MPB_COST_EVAL");
    return !(v && strcmp(v, "generic") == 0);
}

// Launch shape of the packed kernel: warps per CTA x resident CTAs per SM (register cap = 65536 / threads per SM).
// MPB_K2_CFG=<warps>x<ctas> selects one of the compiled experiment shapes (7-dof chains only).
static int k2_cfg() {
    const char* v = getenv("MPB_K2_CFG");
    if (!v) return 0;
    int w = 0, c = 0;
    if (sscanf(v, "%dx%d", &w, &c) != 2) return 0;
    return w * 10 + c;
}

}  // namespace mpb

extern "C" int mpb_cost_eval(const float* x, int B, int H, const mpb_robot_desc* robot,
                             const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp,
                             const float* is_vec, int samples_per_particle, float is_scale,
                             float* cost, float* terms, uint8_t* free_flag, void* stream) {
    return mpb_cost_eval_ex(x, B, H, robot, fields, n_fields, gp, is_vec, samples_per_particle, is_scale, cost, terms,
                            free_flag, nullptr, nullptr, stream);
}

static int cost_eval_impl(const float* x, int B, int H, const mpb_robot_desc* robot,
                          const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp,
                          const float* is_vec, int samples_per_particle, float is_scale,
                          float* cost, float* terms, uint8_t* free_flag,
                          const mpb_extra_cost_desc* extra, float* jl_per_traj, int x_dm, void* stream);

extern "C" int mpb_cost_eval_ex(const float* x, int B, int H, const mpb_robot_desc* robot,
                                const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp,
                                const float* is_vec, int samples_per_particle, float is_scale,
                                float* cost, float* terms, uint8_t* free_flag,
                                const mpb_extra_cost_desc* extra, float* jl_per_traj, void* stream) {
    return cost_eval_impl(x, B, H, robot, fields, n_fields, gp, is_vec, samples_per_particle, is_scale, cost, terms, free_flag, extra,
                          jl_per_traj, 0, stream);
}

extern "C" int mpb_cost_eval_dm_supported(const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields, int H) {
    if (!robot || robot->kind != MPB_ROBOT_CHAIN || robot->q_dim != 7 || H != 64 || !mpb::packed_allowed()) return 0;
    if (getenv("MPB_K2_CFG") || getenv("MPB_K2_LOCAL")) return 0;          // experiment shapes exist in the natural layout only
    for (int i = 0; i < n_fields; ++i)
        if (!fields || fields[i].kind != MPB_FIELD_PRIMITIVES) return 0;
    return 1;
}

extern "C" int mpb_cost_eval_dm(const float* x_dm, int B, int H, const mpb_robot_desc* robot,
                                const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp,
                                const float* is_vec, int samples_per_particle, float is_scale,
                                float* cost, float* terms, uint8_t* free_flag, void* stream) {
    MPB_REQUIRE(mpb_cost_eval_dm_supported(robot, fields, n_fields, H),
                "mpb_cost_eval_dm: dof-major rows need a 7-dof chain, H = 64 and primitive fields only");
    return cost_eval_impl(x_dm, B, H, robot, fields, n_fields, gp, is_vec, samples_per_particle, is_scale, cost, terms, free_flag,
                          nullptr, nullptr, 1, stream);
}

static int cost_eval_impl(const float* x, int B, int H, const mpb_robot_desc* robot,
                          const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp,
                          const float* is_vec, int samples_per_particle, float is_scale,
                          float* cost, float* terms, uint8_t* free_flag,
                          const mpb_extra_cost_desc* extra, float* jl_per_traj, int x_dm, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(B >= 0, "mpb_cost_eval: negative batch size %d", B);
    if (B == 0) return MPB_OK;
    MPB_REQUIRE(x && cost && robot, "mpb_cost_eval: null x/cost/robot");
    MPB_REQUIRE(H >= 2, "mpb_cost_eval: need B >= 0 and H >= 2 (got B=%d H=%d)", B, H);
    MPB_REQUIRE(n_fields >= 0 && n_fields <= MPB_MAX_FIELDS, "mpb_cost_eval: n_fields=%d not in [0,%d]", n_fields, MPB_MAX_FIELDS);
    MPB_REQUIRE(robot->kind == MPB_ROBOT_POINT || robot->kind == MPB_ROBOT_CHAIN, "mpb_cost_eval: unknown robot kind %d", robot->kind);
    MPB_REQUIRE(robot->ws_dim == 2 || robot->ws_dim == 3, "mpb_cost_eval: ws_dim must be 2 or 3");
    if (robot->kind == MPB_ROBOT_POINT)
        MPB_REQUIRE(robot->q_dim == robot->ws_dim && robot->n_spheres == 1 && robot->sphere_r, "mpb_cost_eval: point robot needs q_dim == ws_dim, one sphere");
    else
        MPB_REQUIRE(robot->q_dim >= 1 && robot->q_dim <= MPB_MAX_DOF && robot->ws_dim == 3 && robot->n_spheres >= 1 &&
                    robot->fixed_tf && robot->sphere_link && robot->sphere_off && robot->sphere_r,
                    "mpb_cost_eval: chain robot needs 1..%d joints, ws_dim 3 and a sphere table", MPB_MAX_DOF);
    MPB_REQUIRE(!is_vec || samples_per_particle >= 1, "mpb_cost_eval: samples_per_particle must be >= 1 with is_vec");

    CostArgs a{};
    a.x = x; a.B = B; a.H = H; a.d = robot->q_dim; a.D = 2 * robot->q_dim; a.M = H * a.D;
    a.robot = *robot;
    a.fields.n_fields = n_fields;
    {
        const char* why = validate_fields(fields, n_fields, *robot);
        MPB_REQUIRE(!why, "mpb_cost_eval: %s", why);
    }
    for (int i = 0; i < n_fields; ++i) a.fields.f[i] = fields[i];
    if (gp && gp->enabled) {
        MPB_REQUIRE(gp->start_state && (!gp->has_goal || gp->goal_state), "mpb_cost_eval: gp start/goal state is null");
        a.gp = *gp;
    } else {
        a.gp.enabled = 0;
    }
    bool has_extra_terms = false;
    if (extra && (extra->gp_traj_enabled || extra->jl_enabled)) {
        MPB_REQUIRE(!extra->jl_enabled || (extra->q_min && extra->q_max && jl_per_traj),
                    "mpb_cost_eval: joint-limit term needs q_min, q_max and jl_per_traj");
        a.ex = *extra;
        a.jl_out = jl_per_traj;
        if (extra->gp_traj_enabled && !a.gp.enabled) a.gp.dt = gp ? gp->dt : 0.f;     // Phi needs dt
        MPB_REQUIRE(!extra->gp_traj_enabled || gp, "mpb_cost_eval: the GP-trajectory term takes dt from the gp descriptor");
        has_extra_terms = true;
    }
    a.is_vec = is_vec; a.S = is_vec ? samples_per_particle : 1; a.is_scale = is_scale;
    a.cost = cost; a.terms = terms; a.free_flag = free_flag;

    unsigned off = layout_fields(a.fields, 0);
    off = layout_robot(*robot, a.rl, off);
    // serial chains against primitive fields take the packed two-waypoints-per-lane kernel (cost_eval_packed.cuh)
    bool packed = robot->kind == MPB_ROBOT_CHAIN && !a.fields.has_extra && !has_extra_terms && robot->q_dim >= 2 &&
                  robot->q_dim <= 8 && packed_allowed();
    int cfg = 82;
    if (packed && robot->q_dim == 7) {
        // measured at C4 (profiles/r02_k2_cfg_sweep.txt): without box code the kernel fits 96 registers, and 2 CTAs x 10 warps
        // (5 warps per scheduler) beat 2 x 8 at 116; the mixed instance (boxes and spheres) spills at 96 and keeps 8 x 2
        bool boxes = false, spheres = false;
        for (int i = 0; i < n_fields; ++i) {
            boxes = boxes || fields[i].n_boxes > 0;
            spheres = spheres || fields[i].n_spheres > 0;
        }
        const char* kl = getenv("MPB_K2_LOCAL");
        if (kl && kl[0] == '0') boxes = spheres = true;
        cfg = k2_cfg();
        // box-only instance (C5, table + shelf; profiles/tools/c5_cfg.py): 6 x 3 at 96 registers is the best of 8x2 / 10x2 / 6x3
        if (cfg == 0) cfg = boxes ? (spheres ? 82 : 63) : 102;
    }
    int nw = packed ? cfg / 10 : kWarps;
    if (nw != 4 && nw != 6 && nw != 8 && nw != 10 && nw != 12) nw = 8;
    a.row_stride = (a.M + 3) & ~3;
    a.rows_off = off;
    off += (unsigned)(nw * 2 * a.row_stride * sizeof(float));       // current + prefetched row per warp
    a.queue_off = off;
    // generic: 5 words per entry; packed: + primitive masks, and each warp's per-lane hinge sums behind its queue
    off += packed ? (unsigned)nw * kQ2Stride : (unsigned)(nw * kQCap * 5 * sizeof(float));
    a.list_cap = 8;
    for (int i = 0; i < n_fields; ++i) {
        if (fields[i].kind != MPB_FIELD_PRIMITIVES) continue;
        const int m = fields[i].n_spheres > fields[i].n_boxes ? fields[i].n_spheres : fields[i].n_boxes;
        if (m > a.list_cap) a.list_cap = (m + 7) & ~7;
    }
    a.list_off = off;
    // packed: per-warp primitive DATA lists (52 bytes per entry); generic: index lists
    off += packed ? (unsigned)(nw * a.list_cap * 52) : (unsigned)(nw * 2 * a.list_cap * sizeof(unsigned short));
    if (packed) {
        const unsigned nt = (unsigned)(n_fields * robot->n_spheres);
        off = (off + 15u) & ~15u;
        a.ctab.a = off; off += nt * 16u;
        a.ctab.b = off; off += (nt * 4u + 15u) & ~15u;
    }
    const size_t smem = off;
    MPB_REQUIRE(smem <= 227 * 1024, "mpb_cost_eval: %zu bytes of shared memory needed (H*D too large or too many primitives)", smem);

    MPB_REQUIRE(B <= (1 << 30), "mpb_cost_eval: batch too large (%d)", B);
    {
        const char* v = getenv("MPB_K2_LOCAL");
        a.k2_local = !(v && v[0] == '0');
    }
    a.sched = sched_slot();
    if (!a.sched) { set_error("mpb_cost_eval: could not allocate the scheduler counters"); return MPB_ECUDA; }
    a.x_dm = x_dm;
    MPB_REQUIRE(!x_dm || (packed && robot->q_dim == 7 && H == 64 && (cfg == 102 || cfg == 82 || cfg == 63)),
                "mpb_cost_eval_dm: configuration has no dof-major instance");
    const int blocks_needed = (B + kWarps - 1) / kWarps;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (packed) {
        switch (robot->q_dim) {
            case 2: e = launch_chain2<2, 8, 2>(a, smem, st); break;
            case 3: e = launch_chain2<3, 8, 2>(a, smem, st); break;
            case 4: e = launch_chain2<4, 8, 2>(a, smem, st); break;
            case 5: e = launch_chain2<5, 8, 2>(a, smem, st); break;
            case 6: e = launch_chain2<6, 8, 2>(a, smem, st); break;
            case 8: e = launch_chain2<8, 8, 2>(a, smem, st); break;
            default:
                if (x_dm) {
                    // dof-major rows: the three default instances only (cfg was chosen from the fields above)
                    e = cfg == 102 ? launch_chain2b<7, 10, 2, false, true, true>(a, smem, st)
                        : cfg == 63 ? launch_chain2b<7, 6, 3, true, false, true>(a, smem, st)
                                    : launch_chain2b<7, 8, 2, true, true, true>(a, smem, st);
                    break;
                }
                switch (cfg) {
                    case 83: e = launch_chain2<7, 8, 3>(a, smem, st); break;
                    case 45: e = launch_chain2<7, 4, 5>(a, smem, st); break;
                    case 46: e = launch_chain2<7, 4, 6>(a, smem, st); break;
                    case 63: e = launch_chain2<7, 6, 3>(a, smem, st); break;
                    case 102: e = launch_chain2<7, 10, 2>(a, smem, st); break;
                    case 122: e = launch_chain2<7, 12, 2>(a, smem, st); break;
                    default: e = launch_chain2<7, 8, 2>(a, smem, st); break;
                }
        }
    } else if (a.fields.has_extra || has_extra_terms)   // self-collision / workspace fields or extra terms: the variant that carries them
        e = (robot->kind == MPB_ROBOT_POINT) ? launch<MPB_ROBOT_POINT, 1, true>(a, blocks_needed, smem, st)
                                             : launch<MPB_ROBOT_CHAIN, 4, true>(a, blocks_needed, smem, st);
    else
        e = (robot->kind == MPB_ROBOT_POINT) ? launch<MPB_ROBOT_POINT, 1, false>(a, blocks_needed, smem, st)
                                             : launch<MPB_ROBOT_CHAIN, 4, false>(a, blocks_needed, smem, st);
    if (e != cudaSuccess) {
        set_error("mpb_cost_eval: launch configuration failed: %s", cudaGetErrorString(e));
        return MPB_ECUDA;
    }
    return check_launch("mpb_cost_eval");
}

namespace mpb {
// out[b,j] = sum_t x[b,t,j] * (R[t,t-1] x[b,t-1,j] + R[t,t] x[b,t,j] + R[t,t+1] x[b,t+1,j])
__global__ void smoothness_cost_kernel(const float* __restrict__ x, const float* __restrict__ R, float* __restrict__ out,
                                       int B, int H, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * D) return;
    const int b = (int)(i / D), j = (int)(i % D);
    const float* xb = x + (size_t)b * H * D + j;
    double acc = 0.0;
    for (int t = 0; t < H; ++t) {
        double r = (double)__ldg(R + (size_t)t * H + t) * (double)xb[(size_t)t * D];
        if (t > 0) r += (double)__ldg(R + (size_t)t * H + t - 1) * (double)xb[(size_t)(t - 1) * D];
        if (t < H - 1) r += (double)__ldg(R + (size_t)t * H + t + 1) * (double)xb[(size_t)(t + 1) * D];
        acc += (double)xb[(size_t)t * D] * r;
    }
    out[i] = (float)acc;
}
}  // namespace mpb

extern "C" int mpb_smoothness_cost(const float* x, const float* R, float* out, int B, int H, int D, void* stream) {
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && H >= 1 && D >= 1, "mpb_smoothness_cost: bad sizes");
    if (B == 0) return MPB_OK;
    MPB_REQUIRE(x && R && out, "mpb_smoothness_cost: null pointer");
    const long long n = (long long)B * D;
    smoothness_cost_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, R, out, B, H, D);
    return check_launch("mpb_smoothness_cost");
}
