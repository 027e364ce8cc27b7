// Robot-sphere vs obstacle-primitive collision hinge, shared by the cost, CHOMP and GPMP2 kernels.
//
// Specification (oracle/fields.py; the reference's own SDF lives in the absent torch_robotics):
//   sphere_sdf(x) = ||x - c|| - r ;  box_sdf(x) = ||max(q,0)|| + min(max_k q_k, 0), q = |x - c| - h
//   hinge_s = relu( (radius_s + margin) - min_primitives sdf(centre_s) )
//
// Two passes per robot sphere:
//   1. cull pass  -- branch-free FMA arithmetic, 8 issue slots per sphere/sphere pair and 9 per
//      sphere/box pair, obstacle primitives broadcast from shared memory (LDS.128 + LDS.64) and
//      amortised over a register block of G robot spheres.  It only decides "could the hinge be
//      non-zero?" with a conservative slack, so its rounding never reaches the result.
//   2. exact pass -- only for flagged spheres: every operation is a separately rounded IEEE op in the
//      oracle's order (__fmul_rn/__fadd_rn/__fsqrt_rn cannot be contracted into FMAs), which makes
//      hinge values and collision-free flags bit-identical to the oracle for the same sphere centre.
//      Flagged spheres of a warp are compacted into a shared-memory queue and drained 32 at a time,
//      so the (expensive, divergent) exact pass always runs with full lanes.
#pragma once
#include "mpb_common.cuh"

namespace mpb {

// Byte offsets (from the dynamic shared-memory base) of one staged field + its scalars.
struct FieldLayout {
    unsigned sph, sphx, boxc, boxh;   // float4[n_sph], float2[n_sph], float4[n_box], float4[n_box]
    int n_sph, n_box;
    float margin, weight, inv_sigma2;
    int kind;                         // MPB_FIELD_*
    unsigned pairs, grp, lastb;       // SELF: ushort2[n_pairs]; ushort2[MAX_DOF*MAX_DOF] pair range of link pair (a,b);
    int n_pairs;                      //       int[MAX_DOF] largest b with a non-empty range for first link a (0: none)
    float lo[3], hi[3];               // WORKSPACE
};

struct FieldArgs {              // by-value kernel argument
    int n_fields;
    int has_extra;              // any field that is not MPB_FIELD_PRIMITIVES
    mpb_field_desc f[MPB_MAX_FIELDS];
    FieldLayout l[MPB_MAX_FIELDS];
};

// Host-side validation shared by every entry point that takes fields.  Returns nullptr or a message.
inline const char* validate_fields(const mpb_field_desc* fields, int n_fields, const mpb_robot_desc& robot) {
    if (n_fields < 0 || n_fields > MPB_MAX_FIELDS) return "n_fields out of range";
    if (n_fields > 0 && !fields) return "fields is null";
    for (int i = 0; i < n_fields; ++i) {
        const mpb_field_desc& d = fields[i];
        if (d.kind == MPB_FIELD_PRIMITIVES) {
            if (d.n_spheres < 0 || d.n_boxes < 0 || (d.n_spheres > 0 && !d.spheres) || (d.n_boxes > 0 && !d.boxes))
                return "a primitive field has inconsistent sphere / box arrays";
            if (d.n_spheres >= 65536 || d.n_boxes >= 65536) return "too many primitives in one field";
        } else if (d.kind == MPB_FIELD_SELF) {
            if (d.n_pairs < 0 || d.n_pairs > MPB_MAX_SELF_PAIRS || (d.n_pairs > 0 && !d.pairs))
                return "a self-collision field needs 0..MPB_MAX_SELF_PAIRS sphere pairs";
            if (d.n_pairs > 0 && robot.kind != MPB_ROBOT_CHAIN) return "self-collision fields need a chain robot";
            if (robot.n_spheres >= 65536) return "too many robot spheres for a self-collision field";
        } else if (d.kind == MPB_FIELD_WORKSPACE) {
            for (int k = 0; k < robot.ws_dim; ++k)
                if (!(d.ws_min[k] < d.ws_max[k])) return "workspace field needs ws_min < ws_max on every axis";
        } else {
            return "unknown field kind";
        }
    }
    return nullptr;
}

// Fills fa.l[] starting at byte offset `base` (16-byte aligned); returns the end offset.
inline unsigned layout_fields(FieldArgs& fa, unsigned base) {
    unsigned p = base;
    fa.has_extra = 0;
    for (int i = 0; i < fa.n_fields; ++i) {
        const mpb_field_desc& d = fa.f[i];
        FieldLayout& l = fa.l[i];
        l.kind = d.kind;
        if (d.kind != MPB_FIELD_PRIMITIVES) fa.has_extra = 1;
        const int ns = d.kind == MPB_FIELD_PRIMITIVES ? d.n_spheres : 0, nb = d.kind == MPB_FIELD_PRIMITIVES ? d.n_boxes : 0;
        l.sph = p;  p += (unsigned)ns * 16;
        l.boxc = p; p += (unsigned)nb * 16;
        l.boxh = p; p += (unsigned)nb * 16;
        l.sphx = p; p += (unsigned)ns * 8;
        p = (p + 15u) & ~15u;
        l.n_sph = ns; l.n_box = nb;
        l.margin = d.cutoff_margin; l.weight = d.weight; l.inv_sigma2 = d.inv_sigma2;
        l.n_pairs = d.kind == MPB_FIELD_SELF ? d.n_pairs : 0;
        l.pairs = l.grp = l.lastb = p;
        if (l.n_pairs > 0) {
            l.pairs = p; p += (unsigned)l.n_pairs * 4;
            l.grp = p;   p += MPB_MAX_DOF * MPB_MAX_DOF * 4;
            l.lastb = p; p += MPB_MAX_DOF * 4;
            p = (p + 15u) & ~15u;
        }
        for (int k = 0; k < 3; ++k) { l.lo[k] = d.ws_min[k]; l.hi[k] = d.ws_max[k]; }
    }
    return p;
}

// Cooperative staging by the whole CTA (caller syncs afterwards).  Every thread of the CTA must call it.
__device__ __forceinline__ void stage_fields(const FieldArgs& fa, const mpb_robot_desc& robot, unsigned char* smem) {
    for (int i = 0; i < fa.n_fields; ++i) {
        const mpb_field_desc& d = fa.f[i];
        const FieldLayout& l = fa.l[i];
        float4* sph = reinterpret_cast<float4*>(smem + l.sph);
        float2* sphx = reinterpret_cast<float2*>(smem + l.sphx);
        float4* boxc = reinterpret_cast<float4*>(smem + l.boxc);
        float4* boxh = reinterpret_cast<float4*>(smem + l.boxh);
        for (int o = threadIdx.x; o < l.n_sph; o += blockDim.x) {
            const float4 s = reinterpret_cast<const float4*>(d.spheres)[o];
            sph[o] = s;
            sphx[o] = make_float2(-s.w * s.w, -2.f * s.w);
        }
        for (int o = threadIdx.x; o < l.n_box; o += blockDim.x) {
            boxc[o] = reinterpret_cast<const float4*>(d.boxes)[2 * o];
            boxh[o] = reinterpret_cast<const float4*>(d.boxes)[2 * o + 1];
        }
        if (l.n_pairs > 0) {
            // pair list + the [first,last) range of every link pair (the list is sorted by link pair)
            ushort2* pr = reinterpret_cast<ushort2*>(smem + l.pairs);
            ushort2* grp = reinterpret_cast<ushort2*>(smem + l.grp);
            int* lastb = reinterpret_cast<int*>(smem + l.lastb);
            for (int k = threadIdx.x; k < MPB_MAX_DOF * MPB_MAX_DOF; k += blockDim.x) grp[k] = make_ushort2(0, 0);
            for (int k = threadIdx.x; k < MPB_MAX_DOF; k += blockDim.x) lastb[k] = 0;
            __syncthreads();
            for (int p = threadIdx.x; p < l.n_pairs; p += blockDim.x) {
                const int si = d.pairs[2 * p], sj = d.pairs[2 * p + 1];
                pr[p] = make_ushort2((unsigned short)si, (unsigned short)sj);
                const int key = robot.sphere_link[si] * MPB_MAX_DOF + robot.sphere_link[sj];
                int prev = -1, next = -1;
                if (p > 0) prev = robot.sphere_link[d.pairs[2 * p - 2]] * MPB_MAX_DOF + robot.sphere_link[d.pairs[2 * p - 1]];
                if (p + 1 < l.n_pairs) next = robot.sphere_link[d.pairs[2 * p + 2]] * MPB_MAX_DOF + robot.sphere_link[d.pairs[2 * p + 3]];
                if (key != prev) grp[key].x = (unsigned short)p;
                if (key != next) {
                    grp[key].y = (unsigned short)(p + 1);
                    atomicMax(lastb + key / MPB_MAX_DOF, key % MPB_MAX_DOF);
                }
            }
        }
    }
}

// WORKSPACE field: signed distance to the nearest wall, first minimal entry in the oracle's order
// (x-lo_x, y-lo_y, [z-lo_z,] hi_x-x, hi_y-y [, hi_z-z]); every operation is a single rounded subtraction.
template <bool GRAD>
__device__ __forceinline__ float workspace_sdf(const FieldLayout& f, int ws_dim, float cx, float cy, float cz, float* gx,
                                               float* gy, float* gz) {
    float best = __fsub_rn(cx, f.lo[0]);
    int arg = 0;
    float v = __fsub_rn(cy, f.lo[1]);
    if (v < best) { best = v; arg = 1; }
    if (ws_dim == 3) { v = __fsub_rn(cz, f.lo[2]); if (v < best) { best = v; arg = 2; } }
    v = __fsub_rn(f.hi[0], cx); if (v < best) { best = v; arg = 3; }
    v = __fsub_rn(f.hi[1], cy); if (v < best) { best = v; arg = 4; }
    if (ws_dim == 3) { v = __fsub_rn(f.hi[2], cz); if (v < best) { best = v; arg = 5; } }
    if (GRAD) {
        *gx = arg == 0 ? 1.f : (arg == 3 ? -1.f : 0.f);
        *gy = arg == 1 ? 1.f : (arg == 4 ? -1.f : 0.f);
        *gz = arg == 2 ? 1.f : (arg == 5 ? -1.f : 0.f);
    }
    return best;
}

__device__ __forceinline__ float workspace_hinge(const FieldLayout& f, int ws_dim, float cx, float cy, float cz, float b) {
    return fmaxf(__fsub_rn(b, workspace_sdf<false>(f, ws_dim, cx, cy, cz, nullptr, nullptr, nullptr)), 0.f);
}

// ---- pass 1: conservative candidate test for a register block of G sphere centres -------------
// `smem` must be the kernel's extern __shared__ array so that the loads compile to LDS.
template <int G>
__device__ __forceinline__ unsigned cull_block(const unsigned char* smem, const FieldLayout& f, const float (&cx)[G],
                                               const float (&cy)[G], const float (&cz)[G], const float (&b)[G]) {
    float ms[G], mb[G];
#pragma unroll
    for (int k = 0; k < G; ++k) { ms[k] = CUDART_INF_F; mb[k] = CUDART_INF_F; }
    const float4* sph = reinterpret_cast<const float4*>(smem + f.sph);
    const float2* sphx = reinterpret_cast<const float2*>(smem + f.sphx);
#pragma unroll 2
    for (int o = 0; o < f.n_sph; ++o) {
        const float4 s = sph[o];
        const float2 e = sphx[o];
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
    const float4* boxc = reinterpret_cast<const float4*>(smem + f.boxc);
    const float4* boxh = reinterpret_cast<const float4*>(smem + f.boxh);
#pragma unroll 2
    for (int o = 0; o < f.n_box; ++o) {
        const float4 c = boxc[o];
        const float4 h = boxh[o];
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

// Same test restricted to the primitives named in an index list (the output of the per-link broad phase).
template <int G>
__device__ __forceinline__ unsigned cull_list(const unsigned char* smem, const FieldLayout& f, const unsigned short* ls,
                                              int n_ls, const unsigned short* lb, int n_lb, const float (&cx)[G],
                                              const float (&cy)[G], const float (&cz)[G], const float (&b)[G]) {
    float ms[G], mb[G];
#pragma unroll
    for (int k = 0; k < G; ++k) { ms[k] = CUDART_INF_F; mb[k] = CUDART_INF_F; }
    const float4* sph = reinterpret_cast<const float4*>(smem + f.sph);
    const float2* sphx = reinterpret_cast<const float2*>(smem + f.sphx);
#pragma unroll 1
    for (int i = 0; i < n_ls; ++i) {
        const int o = ls[i];
        const float4 s = sph[o];
        const float2 e = sphx[o];
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const float dx = cx[k] - s.x, dy = cy[k] - s.y, dz = cz[k] - s.z;
            float a = fmaf(dx, dx, e.x);
            a = fmaf(dy, dy, a);
            a = fmaf(dz, dz, a);
            a = fmaf(e.y, b[k], a);
            ms[k] = fminf(ms[k], a);
        }
    }
    const float4* boxc = reinterpret_cast<const float4*>(smem + f.boxc);
    const float4* boxh = reinterpret_cast<const float4*>(smem + f.boxh);
#pragma unroll 1
    for (int i = 0; i < n_lb; ++i) {
        const int o = lb[i];
        const float4 c = boxc[o];
        const float4 h = boxh[o];
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const float qx = fabsf(cx[k] - c.x) - h.x;
            const float qy = fabsf(cy[k] - c.y) - h.y;
            const float qz = fabsf(cz[k] - c.z) - h.z;
            mb[k] = fminf(mb[k], fmaxf(fmaxf(qx, qy), qz));
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

// Order-preserving float <-> uint map so that REDUX.MIN/MAX (integer only) can reduce floats in one instruction.
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ float warp_min_f(float f) { return ord2f(__reduce_min_sync(MPB_FULL_MASK, f2ord(f))); }
__device__ __forceinline__ float warp_max_f(float f) { return ord2f(__reduce_max_sync(MPB_FULL_MASK, f2ord(f))); }

// Warp-cooperative broad phase for one link: which primitives of field f can touch ANY collision sphere of
// the link for ANY lane (= waypoint) of the warp?
//   1. the lanes' bounding-sphere centres (bx,by,bz) are reduced to one axis-aligned box [lo,hi] (6 REDUX);
//   2. lane o tests primitive o against that box inflated by Rm = link bounding radius + field margin, so a
//      whole chunk of 32 primitives costs one pass; survivors are compacted into index lists by ballot.
// Conservative: a robot sphere s of the link with a non-zero hinge against primitive o has
// |c_s - o| < r_s + margin + r_o, hence dist(bc, o) < R + margin + r_o (triangle inequality; box SDFs are
// 1-Lipschitz) and bc lies in [lo,hi], so dist([lo,hi], o) < R + margin + r_o: a primitive rejected here
// contributes exactly 0 for every sphere of the link at every waypoint of the warp.
// The caller issues __syncwarp() before reading the lists.
__device__ __forceinline__ void broad_phase(const unsigned char* smem, const FieldLayout& f, float bx, float by, float bz,
                                            float Rm, bool active, int lane, unsigned short* ls, int& n_ls,
                                            unsigned short* lb, int& n_lb) {
    const float lox = warp_min_f(active ? bx : CUDART_INF_F), hix = warp_max_f(active ? bx : -CUDART_INF_F);
    const float loy = warp_min_f(active ? by : CUDART_INF_F), hiy = warp_max_f(active ? by : -CUDART_INF_F);
    const float loz = warp_min_f(active ? bz : CUDART_INF_F), hiz = warp_max_f(active ? bz : -CUDART_INF_F);
    const unsigned lt = (1u << lane) - 1u;
    const float4* sph = reinterpret_cast<const float4*>(smem + f.sph);
    n_ls = 0;
#pragma unroll 1
    for (int o0 = 0; o0 < f.n_sph; o0 += 32) {
        const int o = o0 + lane;
        bool near = false;
        if (o < f.n_sph) {
            const float4 s = sph[o];
            const float dx = fmaxf(fmaxf(lox - s.x, s.x - hix), 0.f);
            const float dy = fmaxf(fmaxf(loy - s.y, s.y - hiy), 0.f);
            const float dz = fmaxf(fmaxf(loz - s.z, s.z - hiz), 0.f);
            const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float t = Rm + s.w;
            near = d2 < fmaf(t * t, 1.001f, 1e-5f);
        }
        const unsigned m = __ballot_sync(MPB_FULL_MASK, near);
        if (near) ls[n_ls + __popc(m & lt)] = (unsigned short)o;
        n_ls += __popc(m);
    }
    const float4* boxc = reinterpret_cast<const float4*>(smem + f.boxc);
    const float4* boxh = reinterpret_cast<const float4*>(smem + f.boxh);
    n_lb = 0;
    const float tb = fmaf(fabsf(Rm), 1e-3f, Rm + 1e-5f);
#pragma unroll 1
    for (int o0 = 0; o0 < f.n_box; o0 += 32) {
        const int o = o0 + lane;
        bool near = false;
        if (o < f.n_box) {
            const float4 c = boxc[o];
            const float4 h = boxh[o];
            const float gx = fmaxf(lox - (c.x + h.x), (c.x - h.x) - hix);     // gap between the two boxes per axis
            const float gy = fmaxf(loy - (c.y + h.y), (c.y - h.y) - hiy);
            const float gz = fmaxf(loz - (c.z + h.z), (c.z - h.z) - hiz);
            near = fmaxf(fmaxf(gx, gy), gz) < tb;
        }
        const unsigned m = __ballot_sync(MPB_FULL_MASK, near);
        if (near) lb[n_lb + __popc(m & lt)] = (unsigned short)o;
        n_lb += __popc(m);
    }
}

// ---- pass 2: exact signed distance in the oracle's operation order -----------------------------
// Returns min over the primitives that can possibly be closer than b (all others have sdf >= b and
// cannot change relu(b - min sdf)).  If GRAD, also returns the unit gradient of the active primitive.
template <bool GRAD>
__device__ __forceinline__ float exact_sdf(const unsigned char* smem, const FieldLayout& f, float cx, float cy, float cz,
                                           float b, float* gx, float* gy, float* gz) {
    float best = CUDART_INF_F;
    float bgx = 0.f, bgy = 0.f, bgz = 0.f;
    const float4* sph = reinterpret_cast<const float4*>(smem + f.sph);
    for (int o = 0; o < f.n_sph; ++o) {
        const float4 s = sph[o];
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
    const float4* boxc = reinterpret_cast<const float4*>(smem + f.boxc);
    const float4* boxh = reinterpret_cast<const float4*>(smem + f.boxh);
    for (int o = 0; o < f.n_box; ++o) {
        const float4 c = boxc[o];
        const float4 h = boxh[o];
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
__device__ __forceinline__ float exact_hinge(const unsigned char* smem, const FieldLayout& f, float cx, float cy, float cz,
                                             float b) {
    const float sd = exact_sdf<false>(smem, f, cx, cy, cz, b, nullptr, nullptr, nullptr);
    return fmaxf(__fsub_rn(b, sd), 0.f);
}

// ---- robot tables in shared memory ---------------------------------------------------------------
struct RobotLayout {
    unsigned sphere;      // float4[n]  ox, oy, oz, radius
    unsigned tf;          // float[dof*12]
    unsigned link;        // int[n]     ascending joint index
    unsigned bound;       // float4[dof] bounding sphere of each link's collision spheres (link frame): mx,my,mz,R
    unsigned link_end;    // int[dof]   one past the last sphere of each link
    unsigned pat;         // int[dof]   rotation pattern of the fixed transform (MPB_TF_*), see frame2_advance
    int n_spheres, dof;
};

// Rotation part of a joint's fixed transform: general, identity, or a quarter turn about x (URDF rpy = (+-pi/2, 0, 0),
// six of the Panda's seven joints).  For the special patterns R * Fr is a signed column permutation: no arithmetic.
enum { MPB_TF_GENERAL = 0, MPB_TF_IDENTITY = 1, MPB_TF_RX_PLUS = 2, MPB_TF_RX_MINUS = 3 };

inline unsigned layout_robot(const mpb_robot_desc& r, RobotLayout& l, unsigned base) {
    l.n_spheres = r.n_spheres; l.dof = r.q_dim;
    if (r.kind != MPB_ROBOT_CHAIN) { l.sphere = l.tf = l.link = l.bound = l.link_end = l.pat = base; return base; }
    unsigned p = base;
    l.sphere = p; p += (unsigned)r.n_spheres * 16;
    l.tf = p;     p += (unsigned)r.q_dim * 48;
    l.bound = p;  p += (unsigned)r.q_dim * 16;
    l.link = p;   p += (unsigned)r.n_spheres * 4;
    l.link_end = p; p += (unsigned)r.q_dim * 4;
    l.pat = p;    p += (unsigned)r.q_dim * 4;
    return (p + 15u) & ~15u;
}

// Called by ALL threads of the CTA (contains a barrier).  Phase A copies the raw tables to shared memory in one global
// round trip; phase B derives the rotation patterns and the per-link bounding spheres from shared memory (deriving them
// from global memory cost three to four more dependent round trips at the start of every CTA).
__device__ __forceinline__ void stage_robot(const mpb_robot_desc& r, const RobotLayout& l, unsigned char* smem) {
    float4* sp = reinterpret_cast<float4*>(smem + l.sphere);
    float* tf = reinterpret_cast<float*>(smem + l.tf);
    int* lk = reinterpret_cast<int*>(smem + l.link);
    for (int s = threadIdx.x; s < r.n_spheres; s += blockDim.x) {
        sp[s] = make_float4(r.sphere_off[3 * s], r.sphere_off[3 * s + 1], r.sphere_off[3 * s + 2], r.sphere_r[s]);
        lk[s] = r.sphere_link[s];
    }
    for (int i = threadIdx.x; i < r.q_dim * 12; i += blockDim.x) tf[i] = r.fixed_tf[i];
    __syncthreads();
    for (int j = threadIdx.x; j < r.q_dim; j += blockDim.x) {
        const float* F = tf + j * 12;               // row-major 3x4
        int pat = MPB_TF_GENERAL;
        if (F[0] == 1.f && F[1] == 0.f && F[2] == 0.f && F[4] == 0.f && F[8] == 0.f) {
            if (F[5] == 1.f && F[6] == 0.f && F[9] == 0.f && F[10] == 1.f) pat = MPB_TF_IDENTITY;
            else if (F[5] == 0.f && F[6] == -1.f && F[9] == 1.f && F[10] == 0.f) pat = MPB_TF_RX_PLUS;
            else if (F[5] == 0.f && F[6] == 1.f && F[9] == -1.f && F[10] == 0.f) pat = MPB_TF_RX_MINUS;
        }
        reinterpret_cast<int*>(smem + l.pat)[j] = pat;
    }
    // per-link bounding spheres: warp w handles links w, w+nwarps, ... with lanes over the sphere table
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    for (int j = wid; j < r.q_dim; j += nwarp) {
        float mx = 0.f, my = 0.f, mz = 0.f;
        int n = 0, end = 0;
        for (int s = lane; s < r.n_spheres; s += 32) {
            const int ls = lk[s];
            const float4 o = sp[s];
            if (ls == j) { mx += o.x; my += o.y; mz += o.z; ++n; }
            if (ls <= j) end = max(end, s + 1);
        }
        mx = warp_sum(mx); my = warp_sum(my); mz = warp_sum(mz);
        n = __reduce_add_sync(MPB_FULL_MASK, n);
        end = __reduce_max_sync(MPB_FULL_MASK, end);
        float R = 0.f;
        if (n > 0) {
            mx /= n; my /= n; mz /= n;
            for (int s = lane; s < r.n_spheres; s += 32) {
                if (lk[s] != j) continue;
                const float4 o = sp[s];
                const float dx = o.x - mx, dy = o.y - my, dz = o.z - mz;
                R = fmaxf(R, sqrtf(dx * dx + dy * dy + dz * dz) + o.w);
            }
            R = warp_max(R);
        }
        if (lane == 0) {
            reinterpret_cast<float4*>(smem + l.bound)[j] = make_float4(mx, my, mz, fmaf(R, 1.0001f, 1e-6f));
            reinterpret_cast<int*>(smem + l.link_end)[j] = end;
        }
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

// F: 12 floats (row-major 3x4) in shared memory; (cs, sn) = cos / sin of the joint angle.
__device__ __forceinline__ void frame_advance(Frame& T, const float* F, float cs, float sn) {
    const float4 f0 = *reinterpret_cast<const float4*>(F);
    const float4 f1 = *reinterpret_cast<const float4*>(F + 4);
    const float4 f2 = *reinterpret_cast<const float4*>(F + 8);
    // translation: t += R * Ft
    T.tx = fmaf(T.r00, f0.w, fmaf(T.r01, f1.w, fmaf(T.r02, f2.w, T.tx)));
    T.ty = fmaf(T.r10, f0.w, fmaf(T.r11, f1.w, fmaf(T.r12, f2.w, T.ty)));
    T.tz = fmaf(T.r20, f0.w, fmaf(T.r21, f1.w, fmaf(T.r22, f2.w, T.tz)));
    // rotation: R <- R * Fr
    const float a00 = fmaf(T.r00, f0.x, fmaf(T.r01, f1.x, T.r02 * f2.x));
    const float a01 = fmaf(T.r00, f0.y, fmaf(T.r01, f1.y, T.r02 * f2.y));
    const float a02 = fmaf(T.r00, f0.z, fmaf(T.r01, f1.z, T.r02 * f2.z));
    const float a10 = fmaf(T.r10, f0.x, fmaf(T.r11, f1.x, T.r12 * f2.x));
    const float a11 = fmaf(T.r10, f0.y, fmaf(T.r11, f1.y, T.r12 * f2.y));
    const float a12 = fmaf(T.r10, f0.z, fmaf(T.r11, f1.z, T.r12 * f2.z));
    const float a20 = fmaf(T.r20, f0.x, fmaf(T.r21, f1.x, T.r22 * f2.x));
    const float a21 = fmaf(T.r20, f0.y, fmaf(T.r21, f1.y, T.r22 * f2.y));
    const float a22 = fmaf(T.r20, f0.z, fmaf(T.r21, f1.z, T.r22 * f2.z));
    // joint rotation about local z
    T.r00 = fmaf(a00, cs, a01 * sn); T.r01 = fmaf(a01, cs, -a00 * sn); T.r02 = a02;
    T.r10 = fmaf(a10, cs, a11 * sn); T.r11 = fmaf(a11, cs, -a10 * sn); T.r12 = a12;
    T.r20 = fmaf(a20, cs, a21 * sn); T.r21 = fmaf(a21, cs, -a20 * sn); T.r22 = a22;
}

}  // namespace mpb
