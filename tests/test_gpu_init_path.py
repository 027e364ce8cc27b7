"""Init-path parity against fixtures from the UNMODIFIED reference (oracle/make_golden_init.py):

* OptimizationPlanner.get_random_trajs  (mp_baselines/planners/base.py:155-202): the reference draws the initial
  particles in fp64 (quirk B8).  We build the factor in fp64 on the host and sample in fp32 on the device: the
  particles agree to ~1e-6 of the noise amplitude (asserted: 1e-5 relative + 1e-5 x amplitude).
* StochGPMP.const_vel_trajectories      (mp_baselines/planners/stoch_gpmp.py:197-210), note the H*dt (not (H-1)*dt)
  in the reference's mean velocity.
* StochGPMP's own init sampling (stoch_gpmp.py:140-160 via MultiMPPrior.sample, fp32) on the goldens' `eps_init`.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_gpu_stoch_gpmp import STOCH, T, assert_close, make_planner

pytestmark = pytest.mark.gpu

from motion_planning_baselines_b200 import configs  # noqa: E402


@pytest.fixture(scope='module')
def dev():
    return dict(device=torch.device('cuda:0'), dtype=torch.float32)


@pytest.mark.parametrize('name', ['init_random_pm2d', 'init_random_panda'])
def test_get_random_trajs_vs_reference(name, dev):
    from motion_planning_baselines_b200.costs import CostCollision, CostComposite
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import STOMP
    from motion_planning_baselines_b200.robots import Robot
    g = load_golden(name)
    m = g['meta']
    cfg = configs.config(m['cfg'])
    robot = Robot(cfg['robot'], dt=m['dt'], tensor_args=dev)
    field = CollisionField(cfg['obstacles'], tensor_args=dev)
    cost = CostComposite(robot, m['H'], [CostCollision(robot, m['H'], field=field, sigma_coll=1e-1, tensor_args=dev)], tensor_args=dev)
    planner = STOMP(n_dof=m['d'], n_support_points=m['H'], num_particles_per_goal=m['P'], num_samples=4, opt_iters=1,
                    dt=m['dt'], start_state=T(g['start']).to(**dev), cost=cost, multi_goal_states=T(g['goal']).to(**dev).unsqueeze(0),
                    temperature=1.0, step_size=0.1, sigma_spectral=0.1, pos_only=False, tensor_args=dev,
                    sigma_start_init=m['sigma_start_init'], sigma_goal_init=m['sigma_goal_init'], sigma_gp_init=m['sigma_gp_init'])
    eps = T(g['eps_init'])
    assert eps.dtype == torch.float64 and eps.shape == (m['P'], 1, m['H'] * 2 * m['d'])
    got = planner.get_random_trajs(eps=eps)
    ref = T(g['means0'])
    line = ref.mean(0, keepdim=True)
    amp = float((ref - line).abs().max())
    assert amp > 1e-3, 'the fixture must carry real noise'
    err = float((got.cpu().double() - ref.double()).abs().max())
    print(f'{name}: |ours(fp64 factor, fp32 sampler) - reference(fp64)| = {err:.2e}, noise amplitude {amp:.2e}')
    assert_close(got, ref, rtol=1e-5, atol=1e-5 * amp, what='initial particles (quirk B8: reference samples in fp64)')
    # the default (no injected noise) path draws on the device and has the same mean / scale
    drawn = planner.get_random_trajs()
    assert drawn.shape == ref.shape and bool(torch.isfinite(drawn).all())
    assert float((drawn[:, 0].cpu() - ref[:, 0]).abs().max()) < 20 * m['sigma_start_init']        # pinned start state


@pytest.mark.parametrize('name', ['init_const_vel_pm3d', 'init_const_vel_panda'])
def test_const_vel_trajectories_vs_reference(name, dev):
    g = load_golden(name)
    m = g['meta']
    planner = make_planner(g, dev, S=2)                             # any means: only the constructor is needed here
    got = planner.const_vel_trajectories(planner.start_state, planner.multi_goal_states)
    assert_close(got.flatten(0, 1), g['means0'], rtol=1e-6, atol=1e-7, what='const-velocity trajectories')
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import StochGPMP
    cfg = configs.config(m['cfg'])
    p2 = StochGPMP(robot=planner.robot, n_dof=m['d'], n_support_points=m['H'], num_particles_per_goal=m['P'], opt_iters=1,
                   dt=m['dt'], start_state=T(g['start']).to(**dev), multi_goal_states=T(g['goal']).to(**dev).unsqueeze(0),
                   collision_fields=[CollisionField(cfg['obstacles'], tensor_args=dev)], tensor_args=dev, num_samples=2,
                   initial_particle_means='const_vel',
                   **{k: m[k] for k in m if k.startswith('sigma_') or k in ('temperature', 'step_size')})
    assert_close(p2._particle_means, g['means0'], rtol=1e-6, atol=1e-7, what="initial_particle_means='const_vel'")


@pytest.mark.parametrize('name', STOCH)
def test_stoch_gpmp_init_sampling_vs_reference(name, dev):
    """reset() without initial means samples the init prior (fp32 in the reference): goldens' eps_init -> means0."""
    g = load_golden(name)
    m = g['meta']
    planner = make_planner(g, dev)
    planner.reset(eps_init=T(g['eps_init']).to(**dev).contiguous())
    ref = T(g['means0'])
    amp = float((ref - ref.mean(0, keepdim=True)).abs().max())
    assert_close(planner._particle_means, ref, rtol=1e-5, atol=1e-5 * amp + 1e-7, what='initial particle means')
