// C-ABI glue of libmpb_b200.so: error reporting, device queries and the fused iteration drivers.
#include <cstdlib>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "mpb_common.cuh"

namespace mpb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return MPB_ECUDA;
    }
    return MPB_OK;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

// Per-device pool of scheduler slots (the library's only persistent allocation: 64 x 8 bytes per device).
namespace {
constexpr int kMaxDev = 64, kSlots = 64;
std::mutex g_pool_mtx;
unsigned* g_pool[kMaxDev] = {};
std::atomic<unsigned> g_slot_seq{0};

// Allocates (first call per device) and zeroes the pool of the current device.  Not capturable: call mpb_init()
// before capturing a CUDA graph that contains mpb_cost_eval.
unsigned* pool_of_current_device(bool rezero) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
    std::lock_guard<std::mutex> lk(g_pool_mtx);
    if (!g_pool[dev]) {
        unsigned* p = nullptr;
        if (cudaMalloc(&p, kSlots * 2 * sizeof(unsigned)) != cudaSuccess) return nullptr;
        if (cudaMemset(p, 0, kSlots * 2 * sizeof(unsigned)) != cudaSuccess) { cudaFree(p); return nullptr; }
        g_pool[dev] = p;
    } else if (rezero) {
        if (cudaMemset(g_pool[dev], 0, kSlots * 2 * sizeof(unsigned)) != cudaSuccess) return nullptr;
    }
    return g_pool[dev];
}
}  // namespace

unsigned* sched_slot() {
    unsigned* pool = pool_of_current_device(false);
    return pool ? pool + 2 * (g_slot_seq.fetch_add(1u) % kSlots) : nullptr;
}

bool pdl_enabled() {
    static const bool on = [] { const char* v = getenv("MPB_PDL"); return !(v && v[0] == '0'); }();
    return on;
}

int init_current_device() { return pool_of_current_device(true) ? MPB_OK : MPB_ECUDA; }

}  // namespace mpb

extern "C" const char* mpb_last_error(void) { return mpb::g_err; }
extern "C" int mpb_version(void) { return 200; }
extern "C" int mpb_init(void) {
    if (mpb::init_current_device() != MPB_OK) {
        mpb::set_error("mpb_init: could not allocate / zero the scheduler slots of the current device");
        return MPB_ECUDA;
    }
    return MPB_OK;
}
extern "C" int mpb_sizeof_desc(int which) {
    switch (which) {
        case 0: return (int)sizeof(mpb_robot_desc);
        case 1: return (int)sizeof(mpb_field_desc);
        case 2: return (int)sizeof(mpb_gp_desc);
        case 3: return (int)sizeof(mpb_extra_cost_desc);
        case 4: return (int)sizeof(mpb_noise_desc);
        default: return -1;
    }
}

// One Stoch-GPMP iteration (mp_baselines/planners/stoch_gpmp.py:291-299): K1 -> matvec -> K2 -> K3,
// all enqueued on `stream` with no host synchronisation.  The prior factor L never changes, so the
// reference's per-iteration re-factorisation (mp_priors_multi.py:120-123) has no counterpart here.
extern "C" int mpb_stoch_gpmp_iter(const float* L, const float* L_split, const float* Sigma_inv, const float* eps, float* mu, float* x,
                                   float* cost, float* weights, float* is_vec, uint8_t* free_flag, int P, int S,
                                   int H, const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields,
                                   const mpb_gp_desc* gp, float temp, float step, void* stream) {
    MPB_REQUIRE(robot, "mpb_stoch_gpmp_iter: robot is null");
    const int D = 2 * robot->q_dim, M = H * D;
    int rc = L_split ? mpb_sample_gp_tc(L_split, L_split + (size_t)M * M, mu, eps, x, P, S, M, stream)
                     : mpb_sample_gp(L, mu, eps, x, P, S, M, stream);
    if (rc) return rc;
    rc = mpb_prior_matvec(Sigma_inv, mu, is_vec, P, M, 2 * D - 1, stream);
    if (rc) return rc;
    rc = mpb_cost_eval(x, P * S, H, robot, fields, n_fields, gp, is_vec, S, temp, cost, nullptr, free_flag, stream);
    if (rc) return rc;
    return mpb_softmax_update(cost, x, mu, weights, nullptr, temp, step, nullptr, P, S, H, D, stream);
}

extern "C" int mpb_stoch_gpmp_iter_kron(const float* L_kron, const void* L_kron_tc, int tc_kind, const float* Sigma_inv, int sigma_inv_structured,
                                        const float* eps, float* mu, float* x,
                                        float* cost, float* weights, float* is_vec, uint8_t* free_flag, int P, int S,
                                        int H, const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields,
                                        const mpb_gp_desc* gp, float temp, float step, void* stream) {
    MPB_REQUIRE(robot, "mpb_stoch_gpmp_iter_kron: robot is null");
    const int D = 2 * robot->q_dim, M = H * D;
    int rc = (tc_kind == 2 && L_kron_tc) ? mpb_sample_gp_kron_umma(static_cast<const float*>(L_kron_tc), mu, eps, x, P, S, H, robot->q_dim, stream)
             : (tc_kind == 1 && L_kron_tc) ? mpb_sample_gp_kron_tc(L_kron_tc, mu, eps, x, P, S, H, robot->q_dim, stream)
                                           : mpb_sample_gp_kron(L_kron, mu, eps, x, P, S, H, robot->q_dim, stream);
    if (rc) return rc;
    rc = sigma_inv_structured ? mpb_prior_matvec_dof(Sigma_inv, mu, is_vec, P, H, robot->q_dim, stream)
                              : mpb_prior_matvec(Sigma_inv, mu, is_vec, P, M, 2 * D - 1, stream);
    if (rc) return rc;
    rc = mpb_cost_eval(x, P * S, H, robot, fields, n_fields, gp, is_vec, S, temp, cost, nullptr, free_flag, stream);
    if (rc) return rc;
    return mpb_softmax_update(cost, x, mu, weights, nullptr, temp, step, nullptr, P, S, H, D, stream);
}

// The same iteration drawing its own noise inside K1 (mpb_sample_gp_kron_tc_rng): the way the reference is called --
// optimize() takes no noise argument (stoch_gpmp.py:281-309), the draw happens inside MultiMPPrior.sample.
extern "C" int mpb_stoch_gpmp_iter_kron_rng(const void* L_kron_tc, const float* Sigma_inv, int sigma_inv_structured,
                                            const mpb_noise_desc* noise, float* mu, float* x, float* cost, float* weights,
                                            float* is_vec, uint8_t* free_flag, int P, int S, int H, const mpb_robot_desc* robot,
                                            const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp, float temp,
                                            float step, void* stream) {
    MPB_REQUIRE(robot && L_kron_tc && noise, "mpb_stoch_gpmp_iter_kron_rng: null robot / factor / noise descriptor");
    const int D = 2 * robot->q_dim, M = H * D;
    int rc = mpb_sample_gp_kron_tc_rng(L_kron_tc, mu, noise, x, P, S, H, robot->q_dim, stream);
    if (rc) return rc;
    rc = sigma_inv_structured ? mpb_prior_matvec_dof(Sigma_inv, mu, is_vec, P, H, robot->q_dim, stream)
                              : mpb_prior_matvec(Sigma_inv, mu, is_vec, P, M, 2 * D - 1, stream);
    if (rc) return rc;
    rc = mpb_cost_eval(x, P * S, H, robot, fields, n_fields, gp, is_vec, S, temp, cost, nullptr, free_flag, stream);
    if (rc) return rc;
    return mpb_softmax_update(cost, x, mu, weights, nullptr, temp, step, nullptr, P, S, H, D, stream);
}

// The same iteration on the Blackwell sampler (mpb_sample_gp_kron_gen: tcgen05 + bulk-async copies, noise layout
// MPB_NOISE_SPMD drawn in the kernel) -- the default for the shapes it supports.
extern "C" int mpb_stoch_gpmp_iter_kron_gen(const void* L_kron_gen, const float* Sigma_inv, int sigma_inv_structured,
                                            const mpb_noise_desc* noise, float* mu, float* x, float* cost, float* weights,
                                            float* is_vec, uint8_t* free_flag, int P, int S, int H, const mpb_robot_desc* robot,
                                            const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp, float temp,
                                            float step, void* stream) {
    return mpb_stoch_gpmp_iter_kron_gen_ex(L_kron_gen, Sigma_inv, sigma_inv_structured, noise, mu, x, cost, weights, is_vec, free_flag,
                                           nullptr, nullptr, P, S, H, robot, fields, n_fields, gp, temp, step, stream);
}

// ... with the two copies of the particle means the reference API implies written by the kernels that touch the means
// anyway: mu_prev (the pre-update means, `_recent_*_particles` of stoch_gpmp.py:300-306) by K1's mat-vec warp, mu_out (the
// clone optimize() returns, stoch_gpmp.py:309) by K3 -- instead of two extra device copies per call.  Either may be NULL.
extern "C" int mpb_stoch_gpmp_iter_kron_gen_ex(const void* L_kron_gen, const float* Sigma_inv, int sigma_inv_structured,
                                               const mpb_noise_desc* noise, float* mu, float* x, float* cost, float* weights,
                                               float* is_vec, uint8_t* free_flag, float* mu_prev, float* mu_out, int P, int S, int H,
                                               const mpb_robot_desc* robot, const mpb_field_desc* fields, int n_fields,
                                               const mpb_gp_desc* gp, float temp, float step, void* stream) {
    MPB_REQUIRE(robot && L_kron_gen && noise, "mpb_stoch_gpmp_iter_kron_gen: null robot / factor / noise descriptor");
    const int D = 2 * robot->q_dim, M = H * D;
    int rc;
    if (sigma_inv_structured) {      // Sigma^-1 mu rides along in K1 (one extra warp per CTA): no mat-vec launch
        rc = mpb_sample_gp_kron_gen_mv(L_kron_gen, mu, noise, x, P, S, H, robot->q_dim, Sigma_inv, is_vec, mu_prev, stream);
        if (rc) return rc;
    } else {
        if (mu_prev) {
            const cudaError_t e = cudaMemcpyAsync(mu_prev, mu, (size_t)P * M * sizeof(float), cudaMemcpyDeviceToDevice,
                                                  static_cast<cudaStream_t>(stream));
            MPB_REQUIRE(e == cudaSuccess, "mpb_stoch_gpmp_iter_kron_gen: %s", cudaGetErrorString(e));
        }
        rc = mpb_sample_gp_kron_gen(L_kron_gen, mu, noise, x, P, S, H, robot->q_dim, stream);
        if (rc) return rc;
        rc = mpb_prior_matvec(Sigma_inv, mu, is_vec, P, M, 2 * D - 1, stream);
        if (rc) return rc;
    }
    rc = mpb_cost_eval(x, P * S, H, robot, fields, n_fields, gp, is_vec, S, temp, cost, nullptr, free_flag, stream);
    if (rc) return rc;
    return mpb_softmax_update_ex(cost, x, mu, weights, nullptr, temp, step, nullptr, mu_out, P, S, H, D, stream);
}

// The same iteration with the sample rows kept DOF-MAJOR between the three kernels (x_dm: [P, S, dof, 2H]; see
// sample_gp_kron_gen_dm.cu for why that layout halves the sampler's time): K1 writes them, K2 and K3 read them in place.
// Samples, costs, weights and means are bit-identical to mpb_stoch_gpmp_iter_kron_gen_ex; only the memory order of x
// differs (mpb_traj_from_dof_major gives the reference order back for `state_samples`).  Needs the structured precision
// (the mat-vec warp) and mpb_cost_eval_dm_supported(robot, fields, n_fields, H).
extern "C" int mpb_stoch_gpmp_iter_kron_gen_dm(const void* L_kron_gen, const float* Sigma_inv, const mpb_noise_desc* noise, float* mu,
                                               float* x_dm, float* cost, float* weights, float* is_vec, uint8_t* free_flag,
                                               float* mu_prev, float* mu_out, int P, int S, int H, const mpb_robot_desc* robot,
                                               const mpb_field_desc* fields, int n_fields, const mpb_gp_desc* gp, float temp,
                                               float step, void* stream) {
    MPB_REQUIRE(robot && L_kron_gen && noise && Sigma_inv, "mpb_stoch_gpmp_iter_kron_gen_dm: null robot / factor / precision / noise descriptor");
    const int D = 2 * robot->q_dim;
    int rc = mpb_sample_gp_kron_gen_dm(L_kron_gen, mu, noise, x_dm, P, S, H, robot->q_dim, Sigma_inv, is_vec, mu_prev, stream);
    if (rc) return rc;
    rc = mpb_cost_eval_dm(x_dm, P * S, H, robot, fields, n_fields, gp, is_vec, S, temp, cost, nullptr, free_flag, stream);
    if (rc) return rc;
    return mpb_softmax_update_dm(cost, x_dm, mu, weights, nullptr, temp, step, mu_out, P, S, H, D, stream);
}

// STOMP: `n_iters` whole iterations (stomp.py:137-160: sample -> cost -> importance-weighted update) enqueued from ONE
// call -- sample_stomp (noise drawn in the kernel, draw counter noise->offset + it), cost_eval, softmax_update with Sigma_R.
// The small STOMP configurations are bound by launch / host latency (BASELINE.json configs[0]: 64 samples), so the
// per-iteration host work is what matters: three asynchronous launches and nothing else.  cost / x / weights hold the
// LAST iteration on return, exactly what the planner attributes expose after optimize().
extern "C" int mpb_stomp_run(const float* L_R, const float* SigmaR, const mpb_noise_desc* noise, float* mu, float* x, float* cost,
                             float* weights, int P, int S, int H, const mpb_robot_desc* robot, const mpb_field_desc* fields,
                             int n_fields, const mpb_gp_desc* gp, float temp, float lr, int n_iters, void* stream) {
    MPB_REQUIRE(robot && noise && L_R && SigmaR && mu && x && cost && weights, "mpb_stomp_run: null pointer");
    MPB_REQUIRE(n_iters >= 0, "mpb_stomp_run: negative iteration count");
    const int D = 2 * robot->q_dim;
    mpb_noise_desc nd = *noise;
    for (int it = 0; it < n_iters; ++it) {
        int rc = mpb_sample_stomp_rng(L_R, mu, &nd, x, P, S, H, D, stream);
        if (rc) return rc;
        rc = mpb_cost_eval(x, P * S, H, robot, fields, n_fields, gp, nullptr, 1, 0.f, cost, nullptr, nullptr, stream);
        if (rc) return rc;
        rc = mpb_softmax_update(cost, x, mu, weights, nullptr, temp, lr, SigmaR, P, S, H, D, stream);
        if (rc) return rc;
        ++nd.offset;
    }
    return MPB_OK;
}
