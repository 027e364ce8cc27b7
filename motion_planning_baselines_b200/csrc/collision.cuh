// Robot-sphere vs obstacle-primitive collision hinge, shared by the cost, CHOMP and GPMP2 kernels.
//
// Specification (oracle/fields.py; the reference's own SDF lives in the absent torch_robotics):
//   sphere_sdf(x) = ||x - c|| - r ;  box_sdf(x) = ||max(q,0)|| + min(max_k q_k, 0), q = |x - c| - h
//   hinge_s = relu( (radius_s + margin) - min_primitives sdf(centre_s) )
//
// Two passes per robot sphere:
//   1. cull pass  -- branch-free FMA arithmetic, 8 issue slots per sphere/sphere pair and 9 per
//      sphere/box pair, obstacle primitives broadcast from shared memory and amortised over a
//      register block of G robot spheres.  It only decides "could the hinge be non-zero?"
//      with a conservative slack, so its rounding never reaches the result.
//   2. exact pass -- only for robot spheres some lane of the warp flagged: every operation is a
//      separately rounded IEEE op in the oracle's order (__fmul_rn/__fadd_rn/__fsqrt_rn cannot be
//      contracted into FMAs), which makes hinge values and collision-free flags bit-identical to
//      the oracle for the same sphere centre.
#pragma once
#include "mpb_common.cuh"

namespace mpb {

struct FieldSmem {
    const float4* sph;    // cx, cy, cz, r
    const float2* sphx;   // -r^2, -2r            (cull pass)
    const float4* boxc;   // cx, cy, cz, 0
    const float4* boxh;   // hx, hy, hz, 0        (hz = +inf in 2-D)
    int n_sph, n_box;
    float margin;
    float weight, inv_sigma2;
};

// Bytes of shared memory needed to stage the obstacle primitives of `n` fields.
inline size_t field_smem_bytes(const mpb_field_desc* f, int n) {
    size_t b = 0;
    for (int i = 0; i < n; ++i) b += (size_t)f[i].n_spheres * 24 + (size_t)f[i].n_boxes * 32;
    return (b + 15) & ~(size_t)15;
}

struct FieldArgs {              // by-value kernel argument
    int n_fields;
    mpb_field_desc f[MPB_MAX_FIELDS];
};

// Cooperative staging by the whole CTA.  `base` must be 16-byte aligned.  Caller syncs afterwards.
__device__ __forceinline__ void stage_fields(const FieldArgs& fa, unsigned char* base, FieldSmem* out) {
    unsigned char* p = base;
    for (int i = 0; i < fa.n_fields; ++i) {
        const mpb_field_desc& d = fa.f[i];
        float4* sph = reinterpret_cast<float4*>(p);
        p += (size_t)d.n_spheres * 16;
        float4* boxc = reinterpret_cast<float4*>(p);
        p += (size_t)d.n_boxes * 16;
        float4* boxh = reinterpret_cast<float4*>(p);
        p += (size_t)d.n_boxes * 16;
        float2* sphx = reinterpret_cast<float2*>(p);
        p += (size_t)d.n_spheres * 8;
        for (int o = threadIdx.x; o < d.n_spheres; o += blockDim.x) {
            float4 s = reinterpret_cast<const float4*>(d.spheres)[o];
            sph[o] = s;
            sphx[o] = make_float2(-s.w * s.w, -2.f * s.w);
        }
        for (int o = threadIdx.x; o < d.n_boxes; o += blockDim.x) {
            boxc[o] = reinterpret_cast<const float4*>(d.boxes)[2 * o];
            boxh[o] = reinterpret_cast<const float4*>(d.boxes)[2 * o + 1];
        }
        if (threadIdx.x == 0) {
            FieldSmem fs;
            fs.sph = sph; fs.sphx = sphx; fs.boxc = boxc; fs.boxh = boxh;
            fs.n_sph = d.n_spheres; fs.n_box = d.n_boxes;
            fs.margin = d.cutoff_margin; fs.weight = d.weight; fs.inv_sigma2 = d.inv_sigma2;
            out[i] = fs;
        }
    }
}

// ---- pass 1: conservative candidate test for a register block of G sphere centres -------------
template <int G>
__device__ __forceinline__ unsigned cull_block(const FieldSmem& f, const float (&cx)[G], const float (&cy)[G],
                                               const float (&cz)[G], const float (&b)[G]) {
    float ms[G], mb[G];
#pragma unroll
    for (int k = 0; k < G; ++k) { ms[k] = CUDART_INF_F; mb[k] = CUDART_INF_F; }
#pragma unroll 2
    for (int o = 0; o < f.n_sph; ++o) {
        const float4 s = f.sph[o];
        const float2 e = f.sphx[o];
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const float dx = cx[k] - s.x, dy = cy[k] - s.y, dz = cz[k] - s.z;
            float a = fmaf(dx, dx, e.x);            // d^2 - r^2 - 2 r b  <  b^2   <=>  d < r + b
            a = fmaf(dy, dy, a);
            a = fmaf(dz, dz, a);
            a = fmaf(e.y, b[k], a);
            ms[k] = fminf(ms[k], a);
        }
    }
#pragma unroll 2
    for (int o = 0; o < f.n_box; ++o) {
        const float4 c = f.boxc[o];
        const float4 h = f.boxh[o];
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const float qx = fabsf(cx[k] - c.x) - h.x;
            const float qy = fabsf(cy[k] - c.y) - h.y;
            const float qz = fabsf(cz[k] - c.z) - h.z;
            mb[k] = fminf(mb[k], fmaxf(fmaxf(qx, qy), qz));     // box_sdf >= max_k q_k
        }
    }
    unsigned cand = 0;
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const bool c = (ms[k] < fmaf(b[k] * b[k], 1.0001f, 1e-6f)) || (mb[k] < fmaf(fabsf(b[k]), 1e-5f, b[k] + 1e-6f));
        cand |= (c ? 1u : 0u) << k;
    }
    return cand;
}

// ---- pass 2: exact signed distance in the oracle's operation order -----------------------------
// Returns min over the primitives that can possibly be closer than b (all others have sdf >= b and
// cannot change relu(b - min sdf)).  If GRAD, also returns the unit gradient of the active primitive.
template <bool GRAD>
__device__ __forceinline__ float exact_sdf(const FieldSmem& f, float cx, float cy, float cz, float b,
                                           float* gx, float* gy, float* gz) {
    float best = CUDART_INF_F;
    float bgx = 0.f, bgy = 0.f, bgz = 0.f;
    for (int o = 0; o < f.n_sph; ++o) {
        const float4 s = f.sph[o];
        const float dx = __fsub_rn(cx, s.x), dy = __fsub_rn(cy, s.y), dz = __fsub_rn(cz, s.z);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const float t = b + s.w;
        if (d2 <= fmaf(t * t, 1.0001f, 1e-6f)) {
            const float dist = __fsqrt_rn(d2);
            const float sd = __fsub_rn(dist, s.w);
            if (sd < best) {
                best = sd;
                if (GRAD) { const float inv = 1.f / dist; bgx = dx * inv; bgy = dy * inv; bgz = dz * inv; }
            }
        }
    }
    for (int o = 0; o < f.n_box; ++o) {
        const float4 c = f.boxc[o];
        const float4 h = f.boxh[o];
        const float dx = __fsub_rn(cx, c.x), dy = __fsub_rn(cy, c.y), dz = __fsub_rn(cz, c.z);
        const float qx = __fsub_rn(fabsf(dx), h.x), qy = __fsub_rn(fabsf(dy), h.y), qz = __fsub_rn(fabsf(dz), h.z);
        const float m = fmaxf(fmaxf(qx, qy), qz);
        if (m < fmaf(fabsf(b), 1e-5f, b + 1e-6f)) {
            const float px = fmaxf(qx, 0.f), py = fmaxf(qy, 0.f), pz = fmaxf(qz, 0.f);
            const float o2 = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
            const float outside = __fsqrt_rn(o2);
            const float sd = __fadd_rn(outside, fminf(m, 0.f));
            if (sd < best) {
                best = sd;
                if (GRAD) {
                    if (m > 0.f) {          // outside: gradient of ||max(q,0)||
                        const float inv = 1.f / outside;
                        bgx = copysignf(px * inv, dx); bgy = copysignf(py * inv, dy); bgz = copysignf(pz * inv, dz);
                    } else {                // inside: gradient of max_k q_k (first maximal axis, as torch.max)
                        bgx = bgy = bgz = 0.f;
                        if (qx >= qy && qx >= qz) bgx = copysignf(1.f, dx);
                        else if (qy >= qz) bgy = copysignf(1.f, dy);
                        else bgz = copysignf(1.f, dz);
                    }
                }
            }
        }
    }
    if (GRAD) { *gx = bgx; *gy = bgy; *gz = bgz; }
    return best;
}

// hinge of one robot sphere against one field, exact arithmetic.
__device__ __forceinline__ float exact_hinge(const FieldSmem& f, float cx, float cy, float cz, float b) {
    const float sd = exact_sdf<false>(f, cx, cy, cz, b, nullptr, nullptr, nullptr);
    return fmaxf(__fsub_rn(b, sd), 0.f);
}

// ---- robot tables in shared memory ---------------------------------------------------------------
struct RobotSmem {
    const float* fixed_tf;     // [dof*12]
    const float4* sphere;      // [n] ox, oy, oz, radius
    const int* link;           // [n] ascending
    int n_spheres, dof;
};

inline size_t robot_smem_bytes(const mpb_robot_desc& r) {
    if (r.kind != MPB_ROBOT_CHAIN) return 0;
    size_t b = (size_t)r.n_spheres * 16 + (size_t)r.q_dim * 48 + (size_t)r.n_spheres * 4;
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ void stage_robot(const mpb_robot_desc& r, unsigned char* base, RobotSmem* out) {
    float4* sp = reinterpret_cast<float4*>(base);
    float* tf = reinterpret_cast<float*>(base + (size_t)r.n_spheres * 16);
    int* lk = reinterpret_cast<int*>(base + (size_t)r.n_spheres * 16 + (size_t)r.q_dim * 48);
    for (int s = threadIdx.x; s < r.n_spheres; s += blockDim.x) {
        sp[s] = make_float4(r.sphere_off[3 * s], r.sphere_off[3 * s + 1], r.sphere_off[3 * s + 2], r.sphere_r[s]);
        lk[s] = r.sphere_link[s];
    }
    for (int i = threadIdx.x; i < r.q_dim * 12; i += blockDim.x) tf[i] = r.fixed_tf[i];
    if (threadIdx.x == 0) {
        RobotSmem rs;
        rs.fixed_tf = tf; rs.sphere = sp; rs.link = lk; rs.n_spheres = r.n_spheres; rs.dof = r.q_dim;
        *out = rs;
    }
}

// One step of the serial chain:  T <- T * F_j * Rz(q_j)   (oracle/robots.py SerialChainRobot.link_frames)
struct Frame {
    float r00, r01, r02, r10, r11, r12, r20, r21, r22, tx, ty, tz;
};

__device__ __forceinline__ void frame_identity(Frame& T) {
    T.r00 = 1.f; T.r01 = 0.f; T.r02 = 0.f; T.r10 = 0.f; T.r11 = 1.f; T.r12 = 0.f;
    T.r20 = 0.f; T.r21 = 0.f; T.r22 = 1.f; T.tx = T.ty = T.tz = 0.f;
}

__device__ __forceinline__ void frame_advance(Frame& T, const float* F, float q) {
    // translation: t += R * Ft
    const float ftx = F[3], fty = F[7], ftz = F[11];
    T.tx = fmaf(T.r00, ftx, fmaf(T.r01, fty, fmaf(T.r02, ftz, T.tx)));
    T.ty = fmaf(T.r10, ftx, fmaf(T.r11, fty, fmaf(T.r12, ftz, T.ty)));
    T.tz = fmaf(T.r20, ftx, fmaf(T.r21, fty, fmaf(T.r22, ftz, T.tz)));
    // rotation: R <- R * Fr
    const float a00 = fmaf(T.r00, F[0], fmaf(T.r01, F[4], T.r02 * F[8]));
    const float a01 = fmaf(T.r00, F[1], fmaf(T.r01, F[5], T.r02 * F[9]));
    const float a02 = fmaf(T.r00, F[2], fmaf(T.r01, F[6], T.r02 * F[10]));
    const float a10 = fmaf(T.r10, F[0], fmaf(T.r11, F[4], T.r12 * F[8]));
    const float a11 = fmaf(T.r10, F[1], fmaf(T.r11, F[5], T.r12 * F[9]));
    const float a12 = fmaf(T.r10, F[2], fmaf(T.r11, F[6], T.r12 * F[10]));
    const float a20 = fmaf(T.r20, F[0], fmaf(T.r21, F[4], T.r22 * F[8]));
    const float a21 = fmaf(T.r20, F[1], fmaf(T.r21, F[5], T.r22 * F[9]));
    const float a22 = fmaf(T.r20, F[2], fmaf(T.r21, F[6], T.r22 * F[10]));
    // joint rotation about local z
    float sn, cs;
    sincosf(q, &sn, &cs);
    T.r00 = fmaf(a00, cs, a01 * sn); T.r01 = fmaf(a01, cs, -a00 * sn); T.r02 = a02;
    T.r10 = fmaf(a10, cs, a11 * sn); T.r11 = fmaf(a11, cs, -a10 * sn); T.r12 = a12;
    T.r20 = fmaf(a20, cs, a21 * sn); T.r21 = fmaf(a21, cs, -a20 * sn); T.r22 = a22;
}

}  // namespace mpb
