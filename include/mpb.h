/*
 * mpb.h -- C ABI of libmpb_b200.so: the B200 (sm_100a) batched trajectory
 * cost-and-update hot path of mp_baselines.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MPB_E* code otherwise;
 *     mpb_last_error() returns a thread-local message for the last failure.
 *   - all data pointers are DEVICE pointers on the current CUDA device unless
 *     the parameter name ends in _host; descriptor structs themselves live on the host.
 *   - fp32, row-major, contiguous.  The caller owns every data buffer.  The only
 *     state the library keeps is one 512-byte pool of work-scheduler counters per
 *     device (64 slots x 2 words, handed out round-robin to the persistent cost
 *     kernel and re-armed by the kernel itself): created by mpb_init() -- or lazily
 *     by the first launch, which is NOT capturable into a CUDA graph -- so call
 *     mpb_init() once per device before capturing, and keep fewer than 64 cost
 *     launches in flight concurrently per device.  Everything else is stream-ordered
 *     and re-entrant (one planner <-> one stream).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *
 * Symbols used below: P particles, S samples per particle, H waypoints, d dof,
 * D = 2d state width (position | velocity), M = H*D, B = number of trajectories.
 *
 * Each entry point names the reference code it replaces (paths relative to the
 * reference repo root, anindex/motion_planning_baselines @ 8a50c3c).
 */
#ifndef MPB_H_
#define MPB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPB_OK 0
#define MPB_EINVAL (-1)     /* bad argument (shape, null pointer, unsupported size) */
#define MPB_ECUDA (-2)      /* a CUDA runtime call or kernel launch failed */
#define MPB_EUNSUPPORTED (-3)

#define MPB_MAX_FIELDS 4
#define MPB_MAX_DOF 8
#define MPB_ROBOT_POINT 0   /* q is the workspace position (2-D / 3-D point mass) */
#define MPB_ROBOT_CHAIN 1   /* serial chain of revolute-z joints with a sphere table */

/* Replaces the duck-typed `robot` the reference passes around
 * (mp_baselines/planners/costs/cost_functions.py:21,50-52: q_dim, get_position,
 * get_velocity, fk_map_collision). */
typedef struct mpb_robot_desc {
    int32_t kind;               /* MPB_ROBOT_POINT | MPB_ROBOT_CHAIN */
    int32_t q_dim;              /* d: 2|3 for a point, <= MPB_MAX_DOF for a chain */
    int32_t ws_dim;             /* 2 | 3 */
    int32_t n_spheres;          /* collision spheres on the robot (1 for a point) */
    const float* fixed_tf;      /* [q_dim,3,4] parent->joint transforms      (chain) */
    const int32_t* sphere_link; /* [n_spheres] ascending joint index of the carrying frame (chain) */
    const float* sphere_off;    /* [n_spheres,3] centre in that frame        (chain) */
    const float* sphere_r;      /* [n_spheres] radii */
} mpb_robot_desc;

/* Replaces one entry of `collision_fields` (mp_baselines/planners/gpmp2.py:72-79), i.e. the
 * object whose compute_cost() FieldFactor calls (costs/factors/field_factor.py:39).  The reference's
 * task.get_collision_fields() returns up to three kinds (examples/panda_spheres_GPMP.py:41-57):
 *   MPB_FIELD_PRIMITIVES  objects: a union of sphere and axis-aligned box primitives,
 *                         err = sum over robot spheres s of relu(r_s + cutoff_margin - sdf(c_s))
 *   MPB_FIELD_SELF        self-collision (chain robots): a list of robot-sphere pairs (i,j) on different links,
 *                         err = sum over pairs of relu(r_i + r_j + cutoff_margin - ||c_i - c_j||)
 *   MPB_FIELD_WORKSPACE   workspace boundaries: sdf(c) = min over axes of min(c - ws_min, ws_max - c),
 *                         err = sum over robot spheres of relu(r_s + cutoff_margin - sdf(c_s)) */
#define MPB_FIELD_PRIMITIVES 0
#define MPB_FIELD_SELF 1
#define MPB_FIELD_WORKSPACE 2
#define MPB_MAX_SELF_PAIRS 4096
typedef struct mpb_field_desc {
    int32_t n_spheres;
    int32_t n_boxes;
    const float* spheres;       /* [n_spheres,4] cx,cy,cz,r          (cz = 0 in 2-D) */
    const float* boxes;         /* [n_boxes,8]  cx,cy,cz,0,hx,hy,hz,0 (hz = +inf in 2-D) */
    float cutoff_margin;
    float weight;               /* CostComposite weight * 1/sigma_coll^2 is applied as
                                   weight * (inv_sigma2 * sum_t err)  (cost_functions.py:85,185-186) */
    float inv_sigma2;
    int32_t kind;               /* MPB_FIELD_* (0 = primitives, the layout of the first seven members) */
    int32_t n_pairs;            /* SELF: number of sphere pairs, <= MPB_MAX_SELF_PAIRS */
    const int32_t* pairs;       /* SELF: [n_pairs,2] robot-sphere indices with link(i) < link(j), sorted by
                                   (link(i), link(j)) so that the pairs of one link pair are contiguous */
    float ws_min[3];            /* WORKSPACE: lower / upper corner (the third entry is ignored in 2-D) */
    float ws_max[3];
} mpb_field_desc;

/* Replaces CostGP + CostGoalPrior parameters (cost_functions.py:234-289,488-536;
 * gp_factor.py:34-50; unary_factor.py:19).  All scalars are the fp32 values the reference
 * stores in its K / Q_inv matrices. */
typedef struct mpb_gp_desc {
    int32_t enabled;            /* 0: no start/GP/goal terms (STOMP / CHOMP / MPPI cost objects) */
    int32_t has_goal;
    float dt;
    float k_start;              /* 1/sigma_start^2 */
    float k_goal;               /* 1/sigma_goal_prior^2 */
    float q11, q12, q22;        /* Q^-1 = [[q11 I, q12 I],[q12 I, q22 I]] */
    float w_gp, w_goal;         /* CostComposite weights (1.0 in build_gpmp2_cost_composite) */
    const float* start_state;   /* [D] start with zero velocity (gpmp2.py:50) */
    const float* goal_state;    /* [D] goal with zero velocity  (gpmp2.py:60-61), or NULL */
} mpb_gp_desc;

/* Replaces the cost terms a caller may append through `extra_costs` (gpmp2.py:82-83) or put into its own
 * CostComposite: CostGPTrajectory (cost_functions.py:317-357) and CostJointLimits (cost_functions.py:393-429). */
typedef struct mpb_extra_cost_desc {
    int32_t gp_traj_enabled;    /* adds w_gp_traj * sum_t e_t^T Q^-1 e_t, e_t = x_{t+1} - Phi x_t (no start term) */
    float t11, t12, t22;        /* Q^-1 = [[t11 I, t12 I],[t12 I, t22 I]] for this term's own sigma_gp */
    float w_gp_traj;
    int32_t jl_enabled;         /* per trajectory: sum_{t,j} relu(q_min_j + eps - q)^2 + relu(q - (q_max_j - eps))^2 */
    float jl_eps;
    float w_jl;                 /* used by mpb_chomp_run only (gradient weight); mpb_cost_eval returns the raw term */
    const float* q_min;         /* [d] */
    const float* q_max;         /* [d] */
} mpb_extra_cost_desc;

/* In-kernel Gaussian noise (csrc/philox.cuh).  Replaces the draws the reference makes inside
 * torch.distributions.MultivariateNormal.rsample (mp_priors_multi.py:253-256, stomp.py:102, priors/gaussian.py:295):
 * element e (flat index) of the VIRTUAL GLOBAL noise tensor of the whole job is output e % 4 of
 * Philox4x32-10(counter = (e / 4, offset), key = seed) pushed through Box-Muller -- independent of how particles or
 * samples are sharded over GPUs.  Global tensor per layout (local extents n0..n3 of the call in brackets):
 *   MPB_NOISE_SPM    [S_glob, P_glob, M]      (Stoch-GPMP / MultiMPPrior.sample;  local [S,P,M])
 *   MPB_NOISE_STOMP  [S_glob, D, P_glob, H]   (STOMP.sample;                       local [S,D,P,H])
 *   MPB_NOISE_MPPI   [C, N_glob, T]           (ControlTrajectoryGaussian.sample;   local [C,N,T], N_glob = P_global,
 *                                              first local sample = s_offset)
 *   MPB_NOISE_SPMD   [S_glob, P_glob, dof, 2H] (the tcgen05 structured sampler mpb_sample_gp_kron_gen: dof-major inside a
 *                                              row, so one Philox call yields four consecutive k of one dof; the dump is
 *                                              written in the local [S,P,M] order, m = n * dof + j, with n3 = dof)
 * The caller advances `offset` by one per draw (per optimize() iteration). */
#define MPB_NOISE_SPM 0
#define MPB_NOISE_SPMD 3
#define MPB_NOISE_STOMP 1
#define MPB_NOISE_MPPI 2
typedef struct mpb_noise_desc {
    uint64_t seed;          /* Philox key */
    uint64_t offset;        /* draw counter */
    int64_t s_offset;       /* global index of this call's first sample   (sample split over ranks; else 0) */
    int64_t p_offset;       /* global index of this call's first particle (particle sharding over ranks; else 0) */
    int64_t P_global;       /* particles (MPPI: control samples) of the whole job, >= offset + local count */
} mpb_noise_desc;
/* out[local layout] = the normals a kernel called with this descriptor consumes (debug / replay entry, and the
 * generator in front of the samplers that have no fused variant).  n3 is ignored for 3-D layouts. */
int mpb_philox_normal(const mpb_noise_desc* noise, int layout, float* out, int n0, int n1, int n2, int n3, void* stream);

/* FP32 FFMA micro-benchmark (csrc/microbench.cu): launches 2*64*iters flop per thread on *threads_out threads; the
 * caller times it with CUDA events.  bench.py uses it as the measured FP32 roofline denominator. */
int mpb_bench_fp32_peak(float* scratch, int iters, long long* threads_out, void* stream);

const char* mpb_last_error(void);
int mpb_version(void);
/* Allocates and zeroes the work-scheduler slots of the CURRENT device (see "Conventions"; no counterpart in the
 * reference, which has no native state).  Idempotent, synchronous, not capturable.  Also the recovery call after a
 * failed launch: it re-arms every slot. */
int mpb_init(void);
/* sizeof(mpb_robot_desc | mpb_field_desc | mpb_gp_desc | mpb_extra_cost_desc | mpb_noise_desc) for which = 0 | 1 | 2 | 3 | 4: lets a foreign-language binding
 * verify its struct layout before the first call. */
int mpb_sizeof_desc(int which);

/* x[p,s,:] = mu[p,:] + L @ eps[s,p,:]
 * Replaces MultiMPPrior.sample (costs/factors/mp_priors_multi.py:253-256) =
 * torch MultivariateNormal.rsample with scale_tril L; L is lower-triangular [M,M].
 * eps is laid out as torch draws it: [S,P,M]. */
int mpb_sample_gp(const float* L, const float* mu, const float* eps, float* x,
                  int P, int S, int M, void* stream);

/* Tensor-core variant of mpb_sample_gp (tcgen05 3xTF32, fp32 accumulation in TMEM; csrc/sample_gp_tc.cu).
 * The factor is pre-split once with mpb_split_tf32 into L_hi (upper 11 mantissa bits) and L_lo (next 11 bits);
 * eps is split on the fly.  Requires M % 16 == 0, M >= 32 and 16-byte aligned pointers
 * (mpb_sample_gp_tc_supported() != 0); same result as mpb_sample_gp to ~1e-7 absolute. */
int mpb_split_tf32(const float* src, float* hi, float* lo, long long n, void* stream);
int mpb_sample_gp_tc_supported(int P, int S, int M);
int mpb_sample_gp_tc(const float* L_hi, const float* L_lo, const float* mu, const float* eps, float* x,
                     int P, int S, int M, void* stream);

/* Structured FP32 variant of mpb_sample_gp for a factor that decouples over the degrees of freedom
 * (csrc/sample_gp_kron.cu).  The reference's prior precision (mp_priors_multi.py:213-251 with the identity-scaled
 * K_s, K_g, Q_c of unary_factor.py:19 / gp_factor.py:23-26) couples only entries of the same dof, and
 * torch's factorisation keeps exact zeros exact, so L @ eps is `dof` independent [2H,2H] triangular mat-vecs.
 *   mpb_sample_gp_kron_pack : checks bit-exactly that every dropped entry of L [M,M] (M = 2*H*dof, state order
 *                             (t,[pos|vel],j)) is 0.0f (*structured = 1, host int) and writes the per-dof blocks
 *                             k-major: LkT [dof][2H][2H], LkT[j][2t'+b][2t+a] = L[(t,a,j),(t',b,j)].  Synchronises
 *                             `stream` (one-off setup).
 *   mpb_sample_gp_kron      : same contract and layouts as mpb_sample_gp; identical to the dense FP32 sum in
 *                             ascending k.  Shapes: mpb_sample_gp_kron_supported(H, dof) != 0; 16-byte aligned pointers. */
int mpb_sample_gp_kron_supported(int H, int dof);
int mpb_sample_gp_kron_pack(const float* L, float* LkT, int H, int dof, int* structured, void* stream);
int mpb_sample_gp_kron(const float* LkT, const float* mu, const float* eps, float* x,
                       int P, int S, int H, int dof, void* stream);
/* Tensor-core variant (warp-level m16n8k16 MMA, two-term fp16 split of both operands = 22 significant bits, fp32
 * accumulation): same contract as mpb_sample_gp_kron; agrees with it to ~2e-6 of the noise amplitude.
 *   mpb_sample_gp_kron_tc_prepare : LkT (from mpb_sample_gp_kron_pack) -> LkF, the per-dof power-of-two scaled,
 *                                   fragment-ordered fp16 hi/lo operand; LkF must hold mpb_sample_gp_kron_tc_bytes(H, dof)
 *                                   bytes, 16-byte aligned.  One-off setup, stream-ordered. */
long long mpb_sample_gp_kron_tc_bytes(int H, int dof);
int mpb_sample_gp_kron_tc_prepare(const float* LkT, void* LkF, int H, int dof, void* stream);
/* mpb_sample_gp_kron_tc_rng: the same sampler drawing its own noise (layout MPB_NOISE_SPM) slot by slot while it fills
 * its shared-memory tile -- no eps tensor, no generator launch; bit-identical to mpb_philox_normal + mpb_sample_gp_kron_tc. */
int mpb_sample_gp_kron_tc_rng(const void* LkF, const float* mu, const mpb_noise_desc* noise, float* x,
                              int P, int S, int H, int dof, void* stream);
int mpb_sample_gp_kron_tc(const void* LkF, const float* mu, const float* eps, float* x,
                          int P, int S, int H, int dof, void* stream);

/* Blackwell path of the structured sampler with the noise drawn in the kernel (csrc/sample_gp_kron_gen.cu): the factor is
 * the M = 128 operand of tcgen05.mma.kind::f16 (two-term fp16 split, fp32 accumulation in tensor memory), a tile is 64
 * samples, warp-specialised producers write Philox / Box-Muller noise (layout MPB_NOISE_SPMD) straight into the operand
 * tiles, the factor streams through shared memory by bulk-async (TMA) copies and finished rows leave by bulk-async
 * stores.  Replaces MultiMPPrior.sample (mp_priors_multi.py:253-256) INCLUDING torch's noise draw.
 *   mpb_sample_gp_kron_gen_prepare : LkT (from mpb_sample_gp_kron_pack) -> Limg, mpb_sample_gp_kron_gen_bytes(H, dof) bytes,
 *                                    16-byte aligned; one-off setup, stream-ordered.
 * Same numbers as mpb_philox_normal(MPB_NOISE_SPMD) + mpb_sample_gp_kron_tc up to the fp32 accumulation order.
 * Supported: H = 64 (2H = 128 rows = the MMA's M) with 2..7 dofs. */
int mpb_sample_gp_kron_gen_supported(int H, int dof);
long long mpb_sample_gp_kron_gen_bytes(int H, int dof);
int mpb_sample_gp_kron_gen_prepare(const float* LkT, void* Limg, int H, int dof, void* stream);
int mpb_sample_gp_kron_gen(const void* Limg, const float* mu, const mpb_noise_desc* noise, float* x, int P, int S, int H,
                           int dof, void* stream);
/* The same launch with one more warp per CTA that computes y[p] = Sigma_inv @ mu[p] (the vector of the importance-sampling
 * term, stoch_gpmp.py:239-241) while the tiles run: bit-identical to mpb_prior_matvec_dof, without its launch.  Sigma_inv
 * must have the per-dof structure mpb_prior_dof_structured verifies; Sigma_inv and y are both NULL (plain sampler) or
 * both given.  mu_copy (optional, needs y): a copy of mu written by the same warp -- the planner's pre-update means. */
int mpb_sample_gp_kron_gen_mv(const void* Limg, const float* mu, const mpb_noise_desc* noise, float* x, int P, int S, int H,
                              int dof, const float* Sigma_inv, float* y, float* mu_copy, void* stream);

/* tcgen05 variant of the structured sampler (csrc/sample_gp_tc.cu, sample_gp_kron_umma_kernel): TMA -> per-dof
 * gather + 3xTF32 split into tensor memory -> one M128 x N32 tcgen05.mma chain per dof with TMEM accumulators -> dofs
 * interleaved back in the epilogue.  Same contract as mpb_sample_gp_kron; agrees with it to ~2e-6 of the noise amplitude.
 *   mpb_sample_gp_kron_umma_prepare : LkT (from mpb_sample_gp_kron_pack) -> Lp, the TF32 hi / lo factor tiles in the
 *                                     order the kernel's TMA boxes read them; Lp holds mpb_sample_gp_kron_umma_floats(H, dof)
 *                                     floats, 16-byte aligned.  One-off setup, stream-ordered. */
int mpb_sample_gp_kron_umma_supported(int H, int dof);
long long mpb_sample_gp_kron_umma_floats(int H, int dof);
int mpb_sample_gp_kron_umma_prepare(const float* LkT, float* Lp, int H, int dof, void* stream);
int mpb_sample_gp_kron_umma(const float* Lp, const float* mu, const float* eps, float* x,
                            int P, int S, int H, int dof, void* stream);

/* STOMP noise: x[p,s,h,j] = mu[p,h,j] + (h==0||h==H-1 ? 0 : sum_k L_R[h,k] eps[s,j,p,k])
 * Replaces STOMP.sample (mp_baselines/planners/stomp.py:97-108); eps is [S,D,P,H]. */
int mpb_sample_stomp(const float* L_R, const float* mu, const float* eps, float* x,
                     int P, int S, int H, int D, void* stream);
/* Same with the noise drawn in the kernel (layout MPB_NOISE_STOMP): bit-identical to mpb_philox_normal + mpb_sample_stomp. */
int mpb_sample_stomp_rng(const float* L_R, const float* mu, const mpb_noise_desc* noise, float* x,
                         int P, int S, int H, int D, void* stream);

/* y[p,:] = Sigma_inv @ mu[p,:] for a banded (block-tridiagonal) Sigma_inv [M,M] with
 * half bandwidth half_bw (= 2D-1).  Accumulated in fp64.  First half of the
 * importance-sampling term of StochGPMP._get_costs (stoch_gpmp.py:239-241). */
int mpb_prior_matvec(const float* Sigma_inv, const float* mu, float* y,
                     int P, int M, int half_bw, void* stream);

/* Structured variant for a precision that couples only entries of the same dof (the reference's A^T Q^-1 A,
 * mp_priors_multi.py:213-251: at most 7 non-zeros per row, at columns i + m*dof, m = -3..3, in the state order
 * (t,[pos|vel],j)).  mpb_prior_dof_structured checks the zero pattern of Sigma_inv [M,M], M = 2*H*dof, bit-exactly
 * (*structured = 1, host int; synchronises `stream`, one-off setup); mpb_prior_matvec_dof is then bit-identical to
 * mpb_prior_matvec with half_bw = 2D-1. */
int mpb_prior_dof_structured(const float* Sigma_inv, int H, int dof, int* structured, void* stream);
int mpb_prior_matvec_dof(const float* Sigma_inv, const float* mu, float* y, int P, int H, int dof, void* stream);

/* Fused FK + collision + GP/goal cost (+ importance-sampling term) of B trajectories.
 * Replaces CostComposite.eval over {CostGP, CostGoalPrior, CostCollision...}
 * (cost_functions.py:70-87,171-189,271-289,523-536), Cost.get_q_pos_vel_and_fk_map (:41-53),
 * FieldFactor.get_error (field_factor.py:17-39) and, when is_vec != NULL, the second half of
 * the IS term: cost[b] += is_scale * dot(x[b], is_vec[b / samples_per_particle]).
 *   x          [B,H,D]
 *   cost       [B]               total
 *   terms      [n_terms,B]|NULL  individual weighted terms: (start+GP), goal (if has_goal), one per field
 *   free_flag  [B]|NULL          1 iff every collision hinge term of the trajectory is exactly 0
 */
int mpb_cost_eval(const float* x, int B, int H,
                  const mpb_robot_desc* robot,
                  const mpb_field_desc* fields, int n_fields,
                  const mpb_gp_desc* gp,
                  const float* is_vec, int samples_per_particle, float is_scale,
                  float* cost, float* terms, uint8_t* free_flag, void* stream);
/* Same with the optional extra terms: the CostGPTrajectory term is added to cost[] (and appended to `terms` after the
 * field rows); the CostJointLimits term is written UNWEIGHTED per trajectory to jl_per_traj [B] and NOT added to
 * cost[]: the reference sums it over the whole batch into one scalar (cost_functions.py:411-424, `.sum(-1)` of a flat
 * list) which its composite then adds to every trajectory -- the caller reduces jl_per_traj with mpb_sum_f64. */
int mpb_cost_eval_ex(const float* x, int B, int H,
                     const mpb_robot_desc* robot,
                     const mpb_field_desc* fields, int n_fields,
                     const mpb_gp_desc* gp,
                     const float* is_vec, int samples_per_particle, float is_scale,
                     float* cost, float* terms, uint8_t* free_flag,
                     const mpb_extra_cost_desc* extra, float* jl_per_traj, void* stream);

/* out[b,j] = x[b,:,j]^T R x[b,:,j] for a tridiagonal R [H,H] (only the three diagonals are read), accumulated in
 * fp64.  Replaces CostSmoothnessCHOMP.eval (cost_functions.py:371-387) with R = CHOMP._get_R_mat (chomp.py:81-101). */
int mpb_smoothness_cost(const float* x, const float* R, float* out, int B, int H, int D, void* stream);

/* w = softmax(-cost/temp) over samples; g = sum_s w (x_s - mu); mu += step * (SigmaR @ g if
 * SigmaR else g).  Replaces StochGPMP._update_distribution (stoch_gpmp.py:267-279) and
 * STOMP._update_distribution (stomp.py:199-220; SigmaR = inverse(R) [H,H]).
 *   cost [P,S], x [P,S,H,D], mu [P,H,D] (in place), weights [P,S] out, grad [P,H,D] out|NULL */
int mpb_softmax_update(const float* cost, const float* x, float* mu, float* weights, float* grad,
                       float temp, float step, const float* SigmaR,
                       int P, int S, int H, int D, void* stream);
/* The same update that also writes the new means to mu_copy [P, H*D] (optional): the clone StochGPMP.optimize returns
 * (stoch_gpmp.py:309) without a device copy of its own. */
int mpb_softmax_update_ex(const float* cost, const float* x, float* mu, float* weights, float* grad,
                          float temp, float step, const float* SigmaR, float* mu_copy, int P, int S, int H, int D,
                          void* stream);

/* One fused Stoch-GPMP iteration = sample_gp -> prior_matvec -> cost_eval(+IS) -> softmax_update.
 * Replaces the body of StochGPMP.optimize (stoch_gpmp.py:291-299).
 * L_split: NULL (FP32 SIMT sampler) or the [2,M,M] output of mpb_split_tf32 (tensor-core sampler).
 * workspace: x [P,S,H,D], cost [P,S], weights [P,S], is_vec [P,M]; free_flag [P*S] may be NULL.
 * mpb_stoch_gpmp_iter_kron: same, sampling through mpb_sample_gp_kron_umma (tc_kind 2, L_kron_tc = Lp),
 * mpb_sample_gp_kron_tc (tc_kind 1, L_kron_tc = LkF) or mpb_sample_gp_kron (tc_kind 0) with the packed factor; sigma_inv_structured != 0 (mpb_prior_dof_structured) selects mpb_prior_matvec_dof. */
int mpb_stoch_gpmp_iter_kron(const float* L_kron, const void* L_kron_tc, int tc_kind, const float* Sigma_inv, int sigma_inv_structured, const float* eps,
                             float* mu, float* x, float* cost, float* weights, float* is_vec,
                             uint8_t* free_flag,
                             int P, int S, int H,
                             const mpb_robot_desc* robot,
                             const mpb_field_desc* fields, int n_fields,
                             const mpb_gp_desc* gp,
                             float temp, float step, void* stream);
/* mpb_stoch_gpmp_iter_kron_rng: the iteration as the reference runs it -- optimize() takes no noise (stoch_gpmp.py:281-309),
 * the draw happens inside the sampler: K1 = mpb_sample_gp_kron_tc_rng (tc_kind 1 operand LkF). */
int mpb_stoch_gpmp_iter_kron_rng(const void* L_kron_tc, const float* Sigma_inv, int sigma_inv_structured,
                                 const mpb_noise_desc* noise, float* mu, float* x, float* cost, float* weights,
                                 float* is_vec, uint8_t* free_flag, int P, int S, int H, const mpb_robot_desc* robot,
                                 const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp, float temp, float step,
                                 void* stream);
/* ... and on the Blackwell sampler mpb_sample_gp_kron_gen (operand from mpb_sample_gp_kron_gen_prepare, noise layout
 * MPB_NOISE_SPMD): the default iteration for the shapes that sampler supports. */
int mpb_stoch_gpmp_iter_kron_gen(const void* L_kron_gen, const float* Sigma_inv, int sigma_inv_structured,
                                 const mpb_noise_desc* noise, float* mu, float* x, float* cost, float* weights,
                                 float* is_vec, uint8_t* free_flag, int P, int S, int H, const mpb_robot_desc* robot,
                                 const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp, float temp, float step,
                                 void* stream);
/* ... with mu_prev [P, H*D] (the pre-update means, written by K1's mat-vec warp) and mu_out [P, H*D] (the updated means,
 * written by K3) as side outputs, either may be NULL: the two copies of the particle means the reference API implies
 * (`_recent_*_particles`, the clone optimize() returns: stoch_gpmp.py:300-309) cost no launches of their own. */
int mpb_stoch_gpmp_iter_kron_gen_ex(const void* L_kron_gen, const float* Sigma_inv, int sigma_inv_structured,
                                    const mpb_noise_desc* noise, float* mu, float* x, float* cost, float* weights,
                                    float* is_vec, uint8_t* free_flag, float* mu_prev, float* mu_out, int P, int S, int H,
                                    const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields,
                                    const mpb_gp_desc* gp, float temp, float step, void* stream);
/* ---- dof-major sample rows (the fused Stoch-GPMP iteration of the 7-dof arm at H = 64) --------------------------------
 * The reference keeps a sampled trajectory as [H][2 dof] (mp_priors_multi.py:253-256 returns [S, P, H, 2 dof]).  Between the
 * kernels of ONE fused iteration nothing but our own code reads the samples, and in the order [dof][2H] (column
 * 2H j + 2 h + pv; pv = 0 position, 1 velocity) the dofs of the structured factor stay independent all the way to global
 * memory: the sampler's work unit becomes (64 samples, one dof), its factor is loaded once per CTA and dof, and its time
 * halves.  Values are bit-identical to the natural-layout entry points; only the memory order differs.
 *
 * mpb_sample_gp_kron_gen_dm       MultiMPPrior.sample incl. the noise draw (mp_priors_multi.py:253-256), as
 *                                 mpb_sample_gp_kron_gen_mv but writing x_dm [P, S, dof, 2H]; Sigma_inv / y / mu_copy optional
 * mpb_cost_eval_dm(_supported)    mpb_cost_eval (cost_functions.py:41-53,171-189; stoch_gpmp.py:235-245) on dof-major rows;
 *                                 is_vec, start / goal states stay in the reference layout.  Supported: 7-dof chain, H = 64,
 *                                 primitive fields (the packed kernel's default instances)
 * mpb_softmax_update_dm           mpb_softmax_update_ex (stoch_gpmp.py:267-279) reading dof-major rows; mu / grad / mu_copy in
 *                                 the reference layout
 * mpb_traj_from/to_dof_major      the permutation itself, [B, dof, 2H] <-> [B, H, 2 dof] (state_samples accessors, tests)
 * mpb_stoch_gpmp_iter_kron_gen_dm the three kernels of mpb_stoch_gpmp_iter_kron_gen_ex on x_dm (stoch_gpmp.py:235-279) */
int mpb_sample_gp_kron_gen_dm(const void* Limg, const float* mu, const mpb_noise_desc* nd, float* x_dm, int P, int S, int H, int dof,
                              const float* Sigma_inv, float* y, float* mu_copy, void* stream);
int mpb_cost_eval_dm_supported(const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields, int H);
int mpb_cost_eval_dm(const float* x_dm, int B, int H, const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields,
                     const mpb_gp_desc* gp, const float* is_vec, int samples_per_particle, float is_scale, float* cost,
                     float* terms, uint8_t* free_flag, void* stream);
int mpb_softmax_update_dm(const float* cost, const float* x_dm, float* mu, float* weights, float* grad, float temp, float step,
                          float* mu_copy, int P, int S, int H, int D, void* stream);
int mpb_traj_from_dof_major(const float* x_dm, float* x, long long B, int H, int dof, void* stream);
int mpb_traj_to_dof_major(const float* x, float* x_dm, long long B, int H, int dof, void* stream);
int mpb_stoch_gpmp_iter_kron_gen_dm(const void* L_kron_gen, const float* Sigma_inv, const mpb_noise_desc* noise, float* mu,
                                    float* x_dm, float* cost, float* weights, float* is_vec, uint8_t* free_flag, float* mu_prev,
                                    float* mu_out, int P, int S, int H, const mpb_robot_desc* robot,
                                    const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp, float temp, float step,
                                    void* stream);
/* STOMP, n_iters iterations from one call (stomp.py:137-160): mpb_sample_stomp_rng (draw counter noise->offset + it),
 * mpb_cost_eval, mpb_softmax_update with Sigma_R -- for the launch-latency-bound small configurations.  The caller
 * advances its draw counter by n_iters. */
int mpb_stomp_run(const float* L_R, const float* SigmaR, const mpb_noise_desc* noise, float* mu, float* x, float* cost,
                  float* weights, int P, int S, int H, const mpb_robot_desc* robot, const mpb_field_desc* fields,
                  int n_fields, const mpb_gp_desc* gp, float temp, float lr, int n_iters, void* stream);
int mpb_stoch_gpmp_iter(const float* L, const float* L_split, const float* Sigma_inv, const float* eps,
                        float* mu, float* x, float* cost, float* weights, float* is_vec,
                        uint8_t* free_flag,
                        int P, int S, int H,
                        const mpb_robot_desc* robot,
                        const mpb_field_desc* fields, int n_fields,
                        const mpb_gp_desc* gp,
                        float temp, float step, void* stream);

/* The whole CHOMP loop (n_iters gradient steps) in one launch, x [P,H,D] updated in place.
 * Replaces CHOMP._run_optimization / _eval (mp_baselines/planners/chomp.py:127-169):
 *   g = d/dx [ sum_f weight_f * inv_sigma2_f * sum_{t>=1} err_f(x_t) ] + smooth_scale * 2 R x
 *   g = clamp(g, +-grad_clip); g[:,0] = g[:,H-1] = 0; x -= lr * g
 * The collision gradient is analytic (the reference uses autograd through FK + SDF).  R [H,H] is the
 * tridiagonal precision of chomp.py:81-101.  smooth_scale = P_global * weight_prior_cost: the reference adds
 * the smoothness cost summed over ALL particles to every particle (quirk B1), so when particles are sharded
 * over GPUs the caller passes the global particle count. */
int mpb_chomp_run(float* x, int P, int H,
                  const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields,
                  const float* R, float smooth_scale, float lr, float grad_clip, int n_iters, void* stream);
/* Same with a CostJointLimits term in the cost (extra->jl_enabled; gradient w_jl * d/dq of the term above). */
int mpb_chomp_run_ex(float* x, int P, int H,
                     const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields,
                     const float* R, float smooth_scale, float lr, float grad_clip, int n_iters,
                     const mpb_extra_cost_desc* extra, void* stream);

/* GPMP2 step, part 1: collision errors and Jacobians of every waypoint.
 * Replaces FieldFactor.get_error(calc_jacobian=True) (costs/factors/field_factor.py:41-57) as used by
 * CostCollision.get_linear_system (cost_functions.py:191-231):
 *   err  [n_fields,B,H]     err[f,b,t] = sum_s relu(r_s + margin - sdf(c_s)),  0 at t = 0 (no collision factor there)
 *   hobs [n_fields,B,H,d]   H_obst = -d err / d q_t
 *   diag_mean [H*d] (fp64) | NULL: (1/B) sum_b sum_f inv_sigma2_f * hobs^2 -- the collision part of
 *   mean_batch diag(A^T K A) that the trust-region variant needs (gpmp2.py:366); a deterministic reduction. */
int mpb_gpmp2_linearize(const float* x, int B, int H,
                        const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields,
                        float* err, float* hobs, double* diag_mean, void* stream);
/* Same with interpolated collision checking (CostComposite.get_linear_system(n_interpolated_points=n),
 * cost_functions.py:115-119; field_factor.py:44-57): err stays the error AT the support points, while
 *   hobs[f,b,t,:] = - d/dq_t  sum_i err_f(p_i),   p = the trajectory linearly up-sampled in joint space with n extra
 * points per segment, first point dropped.  interp_w_host [n+1] (HOST array) holds the interpolation weights
 * w_0 = 0 < w_1 < ... < w_n < 1 of the points inside a segment: p = q_t (1 - w_k) + q_{t+1} w_k. */
#define MPB_MAX_INTERP 32
int mpb_gpmp2_linearize_ex(const float* x, int B, int H,
                           const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields,
                           float* err, float* hobs, double* diag_mean,
                           int n_interp, const float* interp_w_host, void* stream);

/* GPMP2 step, part 2: assemble the block-tridiagonal normal equations (never materialising A, K, J^T J),
 * block-Cholesky solve, update x += step * dtheta.  Replaces CostGP/CostGoalPrior.get_linear_system
 * (cost_functions.py:291-314,538-554), GPMP2._get_grad_terms + get_torch_solve('cholesky') + the update of
 * _step (gpmp2.py:333-368,451-452) and _get_costs (gpmp2.py:493-495).
 *   inv_sigma2 [n_fields] HOST array of 1/sigma_coll^2
 *   diag_mean  NULL -> J^T J = A^T K A + delta I;  else trust region: + delta * diag(mean_b diag(A^T K A))
 *   workspace  mpb_gpmp2_workspace_bytes(B,H,D) bytes (fp64 factor blocks)
 *   cost [B]|NULL  b^T K b at the linearisation point;  dtheta [B,H,D]|NULL the step before scaling */
long long mpb_gpmp2_workspace_bytes(int B, int H, int D);
int mpb_gpmp2_solve(float* x, int B, int H, int d, const mpb_gp_desc* gp,
                    const float* err, const float* hobs, const float* inv_sigma2_host, int n_fields,
                    const double* diag_mean, float delta, float step,
                    double* workspace, float* cost, float* dtheta, void* stream);

/* ---- importance-weight update split over CTAs / GPUs through packed partial records (csrc/softmax_split.cu) ----
 * Same maths as mpb_softmax_update for one problem with 10^3..10^6 samples (BASELINE.json configs[4]) or one problem
 * whose samples are sharded over GPUs; also MPPI.update_controller + _save_best (mppi.py:72-86,164-169).
 * Record = [m, Z, cmin, argmin (int32 bits), v[H*Dw]] with m = max(-cost/temp), Z = sum exp(-cost/temp - m),
 * v = sum exp(-cost/temp - m) (x - mu); mpb_softmax_record_len(H,Dw) floats.
 *   partial: cost [P,S], x [P,S,H,Dfull] of which columns [c0,c0+Dw) of every waypoint are the updated variables,
 *            mu [P,H,Dw] -> rec [n_chunks,P,REC]; sample_offset = global index of local sample 0 (argmin bookkeeping).
 *   combine: rec [R,P,REC] (R = n_chunks x number of ranks after an all-gather along the first axis), merged in
 *            fixed order r = 0..R-1 -> mu updated in place (mu += step * (SigmaR @ g | g)), grad [P,H,Dw]|NULL,
 *            lse [P,2] = (m*, Z*)|NULL, best_cost [P]|NULL, best_idx [P]|NULL (first occurrence of the minimum).
 *   weights: w[p,s] = exp(-cost/temp - m*) / Z*. */
int mpb_softmax_record_len(int H, int Dw);
int mpb_softmax_partial(const float* cost, const float* x, const float* mu, float* rec, float temp,
                        int P, int S, int H, int Dfull, int c0, int Dw, int n_chunks, long long sample_offset, void* stream);
int mpb_softmax_combine(const float* rec, int R, float* mu, float* grad, float* lse, float* best_cost, int32_t* best_idx,
                        float step, const float* SigmaR, int P, int H, int Dw, void* stream);
int mpb_softmax_weights(const float* cost, const float* lse, float* weights, float temp, int P, int S, void* stream);
/* out[0] = sum of v[0..n) accumulated in fp64 in a fixed order (scratch: 1024 doubles). */
int mpb_sum_f64(const float* v, long long n, double* out, double* scratch, void* stream);

/* MPPI control sampling + rollout + quadratic cost + importance-sampling dot (csrc/mppi.cu).
 * Replaces ControlTrajectoryGaussian.sample (priors/gaussian.py:276-298), MPPI.get_state_trajectories_rollout
 * (mppi.py:190-210) with PointParticleDynamics.dynamics (dynamics/point.py:102-140; velocity control, deterministic),
 * the quadratic part of traj_cost (point.py:198-225) and the V @ Cov_i^-1 @ mean_i dots of mppi.py:125-128.
 *   L_ctrl, Cov_inv [C,T,T]; mean [T,C]; eps [C,N,T] (one block per control dimension, as torch draws them)
 *   xu   [N,T,sd+C]  (state | control) rows = the `full_traj` the reference hands to the obstacle cost
 *   quad [N]         pos + ctrl + terminal cost;  isv [N,C] the per-dimension IS dots */
int mpb_mppi_rollout(const float* L_ctrl, const float* Cov_inv, const float* mean, const float* eps,
                     const float* state0, const float* goal, const float* ctrl_min, const float* ctrl_max,
                     float* xu, float* quad, float* isv, int N, int T, int C, int sd,
                     float dt, float discount, float w_pos, float w_ctrl, float w_posT, void* stream);
/* mpb_mppi_rollout_ex: `mean_sample` [T,C]|NULL = the mean the controls are SAMPLED around when it differs from the
 * mean of the IS term: the reference samples from ctrl_dist, whose loc is refreshed only by update_ctrl_dist()
 * (mppi.py:68-70,86), so after pop()/shift() (mppi.py:171-178) it still holds the unshifted mean while the IS term and
 * the update use the shifted self._mean.  eps == NULL: the noise is drawn in the kernel (`noise`, layout MPB_NOISE_MPPI;
 * bit-identical to mpb_philox_normal + eps). */
int mpb_mppi_rollout_ex(const float* L_ctrl, const float* Cov_inv, const float* mean, const float* mean_sample,
                        const float* eps, const mpb_noise_desc* noise, const float* state0, const float* goal,
                        const float* ctrl_min, const float* ctrl_max, float* xu, float* quad, float* isv,
                        int N, int T, int C, int sd, float dt, float discount, float w_pos, float w_ctrl, float w_posT,
                        void* stream);
/* mpb_mppi_rollout_opt: flags = MPB_MPPI_SHARED_FACTOR promises that the C factors L_ctrl[i] are identical (the reference's
 * `cov_prior_type='const_ctrl'` with one control_std: gaussian.py:271-333), so only L_ctrl[0] is staged in shared memory
 * (17 KiB instead of 119 KiB at T = 64, C = 7) and twice as many warps are resident.  Same results bit for bit. */
#define MPB_MPPI_SHARED_FACTOR 1
int mpb_mppi_rollout_opt(const float* L_ctrl, const float* Cov_inv, const float* mean, const float* mean_sample,
                         const float* eps, const mpb_noise_desc* noise, const float* state0, const float* goal,
                         const float* ctrl_min, const float* ctrl_max, float* xu, float* quad, float* isv, int N,
                         int T, int C, int sd, float dt, float discount, float w_pos, float w_ctrl, float w_posT,
                         int flags, void* stream);
/* cost[n] = (quad[n] + energy[0]) + temp*isv[n,0] + temp*isv[n,1] + ...; energy (device fp64 scalar | NULL) is the
 * obstacle cost SUMMED OVER THE BATCH, which the reference adds to every sample (point.py:196, quirk B2). */
int mpb_mppi_finalize(const float* quad, const float* isv, const double* energy, float temp, float* cost,
                      int N, int C, void* stream);

/* ---- plug points of the UNMODIFIED reference planners (csrc/interop.cu) ----------------------------------------
 * A reference planner that keeps its own Python loop calls its duck-typed robot / field / cost objects one at a
 * time; these entry points back those methods, so the objects of this library drop into it as they are.
 *
 * robot.fk_map_collision(q_pos) (cost_functions.py:50-52): link_pos[n,s,:] = centre of collision sphere s at q[n,:].
 * Chain robots only (a point robot's single sphere centre is q itself).  q [N,d] -> link_pos [N,Ns,3]. */
int mpb_fk_spheres(const float* q, long long N, const mpb_robot_desc* robot, float* link_pos, void* stream);
/* Its backward for torch.autograd (the reference differentiates through FK, field_factor.py:52-57, chomp.py:139):
 * grad_q[n,k] = sum_s grad_link[n,s,:] . d link_pos[n,s,:] / d q[n,k]. */
int mpb_fk_spheres_vjp(const float* q, const float* grad_link, long long N, const mpb_robot_desc* robot, float* grad_q,
                       void* stream);
/* field.compute_cost(q_pos, link_pos) (costs/factors/field_factor.py:39): err[n] = the field's hinge sum at the sphere
 * centres link_pos[n] ([N,Ns,ws_dim]; radii and self-collision pair indices refer to robot's sphere table);
 * grad_link [N,Ns,ws_dim]|NULL = d err[n] / d link_pos[n] (relu'(0) = 0 as in torch). */
int mpb_field_cost(const float* link_pos, long long N, const mpb_robot_desc* robot, const mpb_field_desc* field,
                   float* err, float* grad_link, void* stream);
/* grad[b,t,:] = gout[b] * d cost[b] / d x[b,t,:] for the cost mpb_cost_eval_ex computes (fields, start/GP/goal terms,
 * the CostGPTrajectory term) + w_jl * jl_gsum[0] * d jl[b] / d x[b,t,:] for the batch-summed joint-limit term.
 * Backs `cost(x).sum().backward()` of the reference CHOMP loop (chomp.py:134-139) on the fused cost object; the
 * collision part is analytic.  gout [B]|NULL (ones); jl_gsum device scalar|NULL (1); gp->dt is read even when
 * gp->enabled == 0 and the GP-trajectory term is on. */
int mpb_cost_grad(const float* x, int B, int H, const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields,
                  const mpb_gp_desc* gp, const mpb_extra_cost_desc* extra, const float* gout, const float* jl_gsum,
                  float* grad, void* stream);
/* task.compute_collision(q) of the sample-based planners (rrt_base.py:100-101), batched over N configurations, and
 * the per-waypoint flags behind the examples' trajectory statistics (examples/pointmass_dense_2d_CHOMP.py:130-133):
 * row n of q starts at q + n*row_stride (row_stride = d for a list of configurations, D for a [B,H,D] trajectory
 * tensor viewed as B*H states).  in_collision [N]|NULL: 1 iff any hinge of any field is > 0;
 * err [N]|NULL: sum over fields of the unweighted hinge sums. */
int mpb_collision_query(const float* q, long long N, int row_stride, const mpb_robot_desc* robot,
                        const mpb_field_desc* fields, int n_fields, uint8_t* in_collision, float* err, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPB_H_ */
