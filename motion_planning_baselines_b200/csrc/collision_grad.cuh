// Collision error of ONE waypoint against ONE field together with its analytic gradient w.r.t. the joint
// positions -- shared by the CHOMP (chomp.cu) and GPMP2 (gpmp2.cu) kernels.
//
// The reference obtains this gradient by autograd through FK + SDF + hinge
// (mp_baselines/planners/chomp.py:139; costs/factors/field_factor.py:41-57).  Here it is closed form:
//   err      = sum_s relu(b_s - sdf(c_s)),             b_s = radius_s + cutoff_margin
//   d err/dc = -grad sdf(c_s)  where b_s - sdf > 0     (relu'(0) = 0 like torch)
//   d c_s/dq_k = z_k x (c_s - p_k)   for every joint k at or below the link carrying sphere s
// so with per-link force F_j = sum_s f_s and torque tau_j = sum_s c_s x f_s (f_s = d err/dc_s):
//   d err/dq_k = z_k . ( sum_{j>=k} tau_j  -  p_k x sum_{j>=k} F_j ).
// Signed distances come from exact_sdf<true> (collision.cuh): separately rounded IEEE operations in the
// oracle's order, so err is bit-identical to the cost kernel's hinge for the same sphere centre.
// Field kinds: PRIMITIVES and WORKSPACE share the per-sphere path (field_sdf_grad); SELF walks the sphere-pair
// list with a nested chain walk (frames of link a and link b) and feeds the same force / torque accumulators.
#pragma once
#include "collision.cuh"

namespace mpb {

// Signed distance + unit gradient of one sphere centre against a PRIMITIVES or WORKSPACE field.
__device__ __forceinline__ float field_sdf_grad(const unsigned char* smem, const FieldLayout& fl, int ws_dim, float cx,
                                                float cy, float cz, float b, float* gx, float* gy, float* gz) {
    if (fl.kind == MPB_FIELD_WORKSPACE) return workspace_sdf<true>(fl, ws_dim, cx, cy, cz, gx, gy, gz);
    return exact_sdf<true>(smem, fl, cx, cy, cz, b, gx, gy, gz);
}

template <int KIND>
__device__ __forceinline__ float waypoint_err_grad(const unsigned char* smem, const FieldLayout& fl, const RobotLayout& rl,
                                                   int ws_dim, float point_r, const float (&q)[MPB_MAX_DOF], int d,
                                                   float (&gq)[MPB_MAX_DOF]) {
#pragma unroll
    for (int k = 0; k < MPB_MAX_DOF; ++k) gq[k] = 0.f;
    float err = 0.f;
    if (KIND == MPB_ROBOT_POINT) {
        if (fl.kind == MPB_FIELD_SELF) return 0.f;      // a point has no self-collision pairs
        const float cx = q[0], cy = q[1], cz = (ws_dim == 3) ? q[2] : 0.f;
        const float b = __fadd_rn(point_r, fl.margin);
        float gx, gy, gz;
        const float sd = field_sdf_grad(smem, fl, ws_dim, cx, cy, cz, b, &gx, &gy, &gz);
        const float h = __fsub_rn(b, sd);
        if (h > 0.f) {
            err = h;
            gq[0] = -gx;
            gq[1] = -gy;
            if (ws_dim == 3) gq[2] = -gz;
        }
        return err;
    }
    const float4* rsphere = reinterpret_cast<const float4*>(smem + rl.sphere);
    const float* rtf = reinterpret_cast<const float*>(smem + rl.tf);
    const int* rlend = reinterpret_cast<const int*>(smem + rl.link_end);
    const bool self = fl.kind == MPB_FIELD_SELF;
    float zx[MPB_MAX_DOF], zy[MPB_MAX_DOF], zz[MPB_MAX_DOF], px[MPB_MAX_DOF], py[MPB_MAX_DOF], pz[MPB_MAX_DOF];
    float Fx[MPB_MAX_DOF], Fy[MPB_MAX_DOF], Fz[MPB_MAX_DOF], Tx[MPB_MAX_DOF], Ty[MPB_MAX_DOF], Tz[MPB_MAX_DOF];
    float cs[MPB_MAX_DOF], sn[MPB_MAX_DOF];
    Frame T;
    frame_identity(T);
    int s_begin = 0;
#pragma unroll
    for (int j = 0; j < MPB_MAX_DOF; ++j) {
        Fx[j] = Fy[j] = Fz[j] = Tx[j] = Ty[j] = Tz[j] = 0.f;
        zx[j] = zy[j] = zz[j] = px[j] = py[j] = pz[j] = 0.f;
        cs[j] = 1.f; sn[j] = 0.f;
        if (j < d) {
            sincosf(q[j], &sn[j], &cs[j]);
            frame_advance(T, rtf + j * 12, cs[j], sn[j]);
            zx[j] = T.r02; zy[j] = T.r12; zz[j] = T.r22;
            px[j] = T.tx; py[j] = T.ty; pz[j] = T.tz;
            const int s_end = rlend[j];
            float fx = 0.f, fy = 0.f, fz = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
            if (!self) {
#pragma unroll 1
                for (int s = s_begin; s < s_end; ++s) {
                    const float4 o = rsphere[s];
                    const float cx = fmaf(T.r00, o.x, fmaf(T.r01, o.y, fmaf(T.r02, o.z, T.tx)));
                    const float cy = fmaf(T.r10, o.x, fmaf(T.r11, o.y, fmaf(T.r12, o.z, T.ty)));
                    const float cz = fmaf(T.r20, o.x, fmaf(T.r21, o.y, fmaf(T.r22, o.z, T.tz)));
                    const float b = __fadd_rn(o.w, fl.margin);
                    float gx, gy, gz;
                    const float sd = field_sdf_grad(smem, fl, 3, cx, cy, cz, b, &gx, &gy, &gz);
                    const float h = __fsub_rn(b, sd);
                    if (h > 0.f) {
                        err = __fadd_rn(err, h);
                        fx -= gx; fy -= gy; fz -= gz;              // f = -grad sdf
                        tx -= cy * gz - cz * gy;                   // tau += c x f
                        ty -= cz * gx - cx * gz;
                        tz -= cx * gy - cy * gx;
                    }
                }
            }
            Fx[j] = fx; Fy[j] = fy; Fz[j] = fz; Tx[j] = tx; Ty[j] = ty; Tz[j] = tz;
            s_begin = s_end;
        }
    }
    if (self && fl.n_pairs > 0) {
        // err = sum_pairs relu(thr - ||c_i - c_j||): force u = (c_i - c_j)/dist ... on sphere i it is d err/d c_i = -u,
        // on sphere j it is +u; accumulated as per-link force / torque like the object fields above.
        const ushort2* pr = reinterpret_cast<const ushort2*>(smem + fl.pairs);
        const ushort2* grp = reinterpret_cast<const ushort2*>(smem + fl.grp);
        const int* lastb = reinterpret_cast<const int*>(smem + fl.lastb);
        const float4* rbound = reinterpret_cast<const float4*>(smem + rl.bound);
        Frame Ta;
        frame_identity(Ta);
#pragma unroll 1
        for (int a = 0; a < d - 1; ++a) {
            frame_advance(Ta, rtf + a * 12, cs[a], sn[a]);
            const int b_last = lastb[a];
            if (b_last <= a) continue;
            const float4 ba = rbound[a];
            const float ax = fmaf(Ta.r00, ba.x, fmaf(Ta.r01, ba.y, fmaf(Ta.r02, ba.z, Ta.tx)));
            const float ay = fmaf(Ta.r10, ba.x, fmaf(Ta.r11, ba.y, fmaf(Ta.r12, ba.z, Ta.ty)));
            const float az = fmaf(Ta.r20, ba.x, fmaf(Ta.r21, ba.y, fmaf(Ta.r22, ba.z, Ta.tz)));
            Frame Tb = Ta;
#pragma unroll 1
            for (int b = a + 1; b <= b_last; ++b) {
                frame_advance(Tb, rtf + b * 12, cs[b], sn[b]);
                const ushort2 g = grp[a * MPB_MAX_DOF + b];
                if (g.x == g.y) continue;
                const float4 bb = rbound[b];
                const float ex = fmaf(Tb.r00, bb.x, fmaf(Tb.r01, bb.y, fmaf(Tb.r02, bb.z, Tb.tx))) - ax;
                const float ey = fmaf(Tb.r10, bb.x, fmaf(Tb.r11, bb.y, fmaf(Tb.r12, bb.z, Tb.ty))) - ay;
                const float ez = fmaf(Tb.r20, bb.x, fmaf(Tb.r21, bb.y, fmaf(Tb.r22, bb.z, Tb.tz))) - az;
                const float reach = ba.w + bb.w + fl.margin;
                if (!(fmaf(ez, ez, fmaf(ey, ey, ex * ex)) < fmaf(reach * reach, 1.001f, 1e-6f))) continue;
                float fax = 0.f, fay = 0.f, faz = 0.f, tax = 0.f, tay = 0.f, taz = 0.f;
                float fbx = 0.f, fby = 0.f, fbz = 0.f, tbx = 0.f, tby = 0.f, tbz = 0.f;
#pragma unroll 1
                for (int p = g.x; p < g.y; ++p) {
                    const ushort2 ij = pr[p];
                    const float4 oi = rsphere[ij.x], oj = rsphere[ij.y];
                    const float cix = fmaf(Ta.r00, oi.x, fmaf(Ta.r01, oi.y, fmaf(Ta.r02, oi.z, Ta.tx)));
                    const float ciy = fmaf(Ta.r10, oi.x, fmaf(Ta.r11, oi.y, fmaf(Ta.r12, oi.z, Ta.ty)));
                    const float ciz = fmaf(Ta.r20, oi.x, fmaf(Ta.r21, oi.y, fmaf(Ta.r22, oi.z, Ta.tz)));
                    const float cjx = fmaf(Tb.r00, oj.x, fmaf(Tb.r01, oj.y, fmaf(Tb.r02, oj.z, Tb.tx)));
                    const float cjy = fmaf(Tb.r10, oj.x, fmaf(Tb.r11, oj.y, fmaf(Tb.r12, oj.z, Tb.ty)));
                    const float cjz = fmaf(Tb.r20, oj.x, fmaf(Tb.r21, oj.y, fmaf(Tb.r22, oj.z, Tb.tz)));
                    const float dx = __fsub_rn(cix, cjx), dy = __fsub_rn(ciy, cjy), dz = __fsub_rn(ciz, cjz);
                    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    const float thr = __fadd_rn(__fadd_rn(oi.w, oj.w), fl.margin);
                    if (!(d2 < fmaf(thr * thr, 1.0001f, 1e-6f))) continue;
                    const float dist = __fsqrt_rn(d2);
                    const float h = __fsub_rn(thr, dist);
                    if (h > 0.f) {
                        err = __fadd_rn(err, h);
                        const float inv = 1.f / dist;
                        const float ux = dx * inv, uy = dy * inv, uz = dz * inv;
                        fax -= ux; fay -= uy; faz -= uz;                                  // d err / d c_i = -u
                        tax -= ciy * uz - ciz * uy; tay -= ciz * ux - cix * uz; taz -= cix * uy - ciy * ux;
                        fbx += ux; fby += uy; fbz += uz;                                  // d err / d c_j = +u
                        tbx += cjy * uz - cjz * uy; tby += cjz * ux - cjx * uz; tbz += cjx * uy - cjy * ux;
                    }
                }
#pragma unroll
                for (int j = 0; j < MPB_MAX_DOF; ++j) {
                    if (j == a) { Fx[j] += fax; Fy[j] += fay; Fz[j] += faz; Tx[j] += tax; Ty[j] += tay; Tz[j] += taz; }
                    if (j == b) { Fx[j] += fbx; Fy[j] += fby; Fz[j] += fbz; Tx[j] += tbx; Ty[j] += tby; Tz[j] += tbz; }
                }
            }
        }
    }
    float aFx = 0.f, aFy = 0.f, aFz = 0.f, aTx = 0.f, aTy = 0.f, aTz = 0.f;
#pragma unroll
    for (int j = MPB_MAX_DOF - 1; j >= 0; --j) {
        if (j < d) {
            aFx += Fx[j]; aFy += Fy[j]; aFz += Fz[j];
            aTx += Tx[j]; aTy += Ty[j]; aTz += Tz[j];
            const float mx = aTx - (py[j] * aFz - pz[j] * aFy);
            const float my = aTy - (pz[j] * aFx - px[j] * aFz);
            const float mz = aTz - (px[j] * aFy - py[j] * aFx);
            gq[j] = zx[j] * mx + zy[j] * my + zz[j] * mz;
        }
    }
    return err;
}

}  // namespace mpb
