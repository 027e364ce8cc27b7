"""mpb_stomp_run: all iterations of STOMP.optimize(opt_iters=n) enqueued from one C call (stomp.py:137-160).  It must be
the same computation as n single-iteration calls: same kernels, same noise (draw counter offset + it) -> bit-identical
means, costs, weights and samples; the draw counter advances by n."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    return dict(device=torch.device('cuda:0'), dtype=torch.float32)


def make(cfg_name, dev, seed=5, S=64):
    import bench_configs
    from motion_planning_baselines_b200 import configs
    from motion_planning_baselines_b200.planners import STOMP
    cfg = configs.config(cfg_name)
    prm, H = cfg['params'] if cfg_name == 'C1' else configs.config('C1')['params'], 64
    d = cfg['robot'].q_dim
    return STOMP(n_dof=d, n_support_points=H, num_particles_per_goal=2, num_samples=S, opt_iters=1, dt=cfg['dt'],
                 start_state=torch.tensor(cfg['start']).to(**dev), cost=bench_configs._collision_cost(cfg, H, prm['sigma_coll'], dev),
                 multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0), temperature=prm['temperature'],
                 step_size=prm['step_size'], sigma_spectral=prm['sigma_spectral'],
                 initial_particle_means=bench_configs._straight(cfg, 2, H, d, dev, jitter=0.2), pos_only=False, tensor_args=dev, seed=seed)


@pytest.mark.parametrize('cfg_name', ['C1', 'C5'])
def test_fused_iterations_equal_single_iteration_calls(cfg_name, dev):
    a, b = make(cfg_name, dev), make(cfg_name, dev)
    assert a._can_run_fused(None, {}), 'the plain collision composite must take the fused path'
    n = 7
    ta = a.optimize(opt_iters=n)
    for _ in range(n):
        tb = b.optimize(opt_iters=1)
    assert a._noise.offset == b._noise.offset
    assert torch.equal(ta, tb) and torch.equal(a.costs, b.costs)
    assert torch.equal(a._weights, b._weights) and torch.equal(a.state_particles, b.state_particles)
    # injected noise and observation kwargs keep the staged path
    assert not a._can_run_fused(torch.zeros(1), {}) and not a._can_run_fused(None, dict(goal=1))
    assert torch.equal(a.optimize(opt_iters=0), ta)


def test_wide_stomp_sampler_is_bit_identical_to_the_per_sample_kernel(dev):
    """mpb_sample_stomp switches to one thread per (particle, sample, column) for many samples (csrc/sample_gp.cu,
    sample_stomp_wide_kernel); same sums in the same order -> same bits, injected or in-kernel noise (stomp.py:97-108)."""
    import ctypes as C
    from motion_planning_baselines_b200 import _lib
    lib = _lib.lib()
    P, S, H, D = 2, 1500, 64, 14
    gen = torch.Generator().manual_seed(3)
    A = torch.randn(H, H, generator=gen)
    L = torch.linalg.cholesky(A @ A.t() + H * torch.eye(H)).to(**dev).contiguous()        # lower triangular, exact zeros above
    mu = torch.randn(P, H, D, generator=gen).to(**dev)
    eps = torch.randn(S, D, P, H, generator=gen).to(**dev)
    x_wide = torch.empty(P, S, H, D, **dev)
    _lib.check(lib.mpb_sample_stomp(_lib.ptr(L), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x_wide), P, S, H, D, _lib.stream_ptr()))
    # the same samples 40 at a time: below the switch-over, i.e. through the CTA-per-sample kernel
    for s0 in (0, 40, 1440):
        e = eps[s0:s0 + 40].contiguous()
        x_small = torch.empty(P, 40, H, D, **dev)
        _lib.check(lib.mpb_sample_stomp(_lib.ptr(L), _lib.ptr(mu), _lib.ptr(e), _lib.ptr(x_small), P, 40, H, D, _lib.stream_ptr()))
        assert torch.equal(x_wide[:, s0:s0 + 40], x_small)
    ref = mu.unsqueeze(1).double() + torch.einsum('hk,sjpk->pshj', L.double(), eps.double())
    ref[:, :, 0], ref[:, :, -1] = mu[:, None, 0].double(), mu[:, None, -1].double()
    assert float((x_wide.double() - ref).abs().max()) < 1e-4
    # in-kernel noise: wide kernel == dump + injected
    nd = _lib.NoiseDesc(seed=5, offset=2, s_offset=7, p_offset=1, P_global=P + 3)
    x_rng = torch.empty(P, S, H, D, **dev)
    _lib.check(lib.mpb_sample_stomp_rng(_lib.ptr(L), _lib.ptr(mu), C.byref(nd), _lib.ptr(x_rng), P, S, H, D, _lib.stream_ptr()))
    e = _lib.philox_normal(nd, _lib.NOISE_STOMP, (S, D, P, H), dev['device'])
    x_inj = torch.empty(P, S, H, D, **dev)
    _lib.check(lib.mpb_sample_stomp(_lib.ptr(L), _lib.ptr(mu), _lib.ptr(e), _lib.ptr(x_inj), P, S, H, D, _lib.stream_ptr()))
    assert torch.equal(x_rng, x_inj)
