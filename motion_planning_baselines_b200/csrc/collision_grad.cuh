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
#pragma once
#include "collision.cuh"

namespace mpb {

template <int KIND>
__device__ __forceinline__ float waypoint_err_grad(const unsigned char* smem, const FieldLayout& fl, const RobotLayout& rl,
                                                   int ws_dim, float point_r, const float (&q)[MPB_MAX_DOF], int d,
                                                   float (&gq)[MPB_MAX_DOF]) {
#pragma unroll
    for (int k = 0; k < MPB_MAX_DOF; ++k) gq[k] = 0.f;
    float err = 0.f;
    if (KIND == MPB_ROBOT_POINT) {
        const float cx = q[0], cy = q[1], cz = (ws_dim == 3) ? q[2] : 0.f;
        const float b = __fadd_rn(point_r, fl.margin);
        float gx, gy, gz;
        const float sd = exact_sdf<true>(smem, fl, cx, cy, cz, b, &gx, &gy, &gz);
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
    float zx[MPB_MAX_DOF], zy[MPB_MAX_DOF], zz[MPB_MAX_DOF], px[MPB_MAX_DOF], py[MPB_MAX_DOF], pz[MPB_MAX_DOF];
    float Fx[MPB_MAX_DOF], Fy[MPB_MAX_DOF], Fz[MPB_MAX_DOF], Tx[MPB_MAX_DOF], Ty[MPB_MAX_DOF], Tz[MPB_MAX_DOF];
    Frame T;
    frame_identity(T);
    int s_begin = 0;
#pragma unroll
    for (int j = 0; j < MPB_MAX_DOF; ++j) {
        Fx[j] = Fy[j] = Fz[j] = Tx[j] = Ty[j] = Tz[j] = 0.f;
        zx[j] = zy[j] = zz[j] = px[j] = py[j] = pz[j] = 0.f;
        if (j < d) {
            float sn, cs;
            sincosf(q[j], &sn, &cs);
            frame_advance(T, rtf + j * 12, cs, sn);
            zx[j] = T.r02; zy[j] = T.r12; zz[j] = T.r22;
            px[j] = T.tx; py[j] = T.ty; pz[j] = T.tz;
            const int s_end = rlend[j];
            float fx = 0.f, fy = 0.f, fz = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
#pragma unroll 1
            for (int s = s_begin; s < s_end; ++s) {
                const float4 o = rsphere[s];
                const float cx = fmaf(T.r00, o.x, fmaf(T.r01, o.y, fmaf(T.r02, o.z, T.tx)));
                const float cy = fmaf(T.r10, o.x, fmaf(T.r11, o.y, fmaf(T.r12, o.z, T.ty)));
                const float cz = fmaf(T.r20, o.x, fmaf(T.r21, o.y, fmaf(T.r22, o.z, T.tz)));
                const float b = __fadd_rn(o.w, fl.margin);
                float gx, gy, gz;
                const float sd = exact_sdf<true>(smem, fl, cx, cy, cz, b, &gx, &gy, &gz);
                const float h = __fsub_rn(b, sd);
                if (h > 0.f) {
                    err = __fadd_rn(err, h);
                    fx -= gx; fy -= gy; fz -= gz;              // f = -grad sdf
                    tx -= cy * gz - cz * gy;                   // tau += c x f
                    ty -= cz * gx - cx * gz;
                    tz -= cx * gy - cy * gx;
                }
            }
            Fx[j] = fx; Fy[j] = fy; Fz[j] = fz; Tx[j] = tx; Ty[j] = ty; Tz[j] = tz;
            s_begin = s_end;
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
