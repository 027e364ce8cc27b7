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
