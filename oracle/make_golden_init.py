"""Golden fixtures for the INIT path: OptimizationPlanner.get_random_trajs
(mp_baselines/planners/base.py:155-202, fp64 by quirk B8) and StochGPMP.const_vel_trajectories
(mp_baselines/planners/stoch_gpmp.py:197-210), from the UNMODIFIED reference planners with the noise they draw recorded.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only (needs /root/reference):

    python oracle/make_golden_init.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import make_golden as mg  # noqa: E402
from oracle.make_golden import TA, NoiseRecorder, configs, oracle_field, oracle_robot, save  # noqa: E402


def gen_init_stomp(tag, cfg_name, P, H, seed, sig):
    """STOMP / CHOMP / GPMP2 share base.get_random_trajs; STOMP's constructor draws it first."""
    from mp_baselines.planners.costs.cost_functions import CostCollision, CostComposite
    from mp_baselines.planners.stomp import STOMP
    cfg = configs.config(cfg_name)
    model, obst = cfg['robot'], cfg['obstacles']
    robot, field = oracle_robot(model, cfg['dt']), oracle_field(obst, model)
    torch.manual_seed(seed)
    start, goal = torch.tensor(cfg['start'], **TA), torch.tensor(cfg['goal'], **TA)
    cost = CostComposite(robot, H, [CostCollision(robot, H, field=field, sigma_coll=1e-1, tensor_args=TA)], tensor_args=TA)
    with NoiseRecorder() as rec:
        planner = STOMP(n_dof=model.q_dim, n_support_points=H, num_particles_per_goal=P, num_samples=4, opt_iters=1,
                        dt=cfg['dt'], start_state=start, cost=cost, multi_goal_states=goal.unsqueeze(0), temperature=1.0,
                        step_size=0.1, sigma_spectral=0.1, pos_only=False, tensor_args=TA, **sig)
    eps_init = rec.draws[0]
    assert eps_init.dtype == torch.float64, 'quirk B8: the reference draws the initial particles in fp64'
    meta = dict(cfg=cfg_name, P=P, H=H, d=model.q_dim, dt=cfg['dt'], seed=seed, **sig)
    save(tag, meta=np.array(repr(meta)), start=start, goal=goal, eps_init=eps_init, means0=planner._particle_means)


def gen_const_vel(tag, cfg_name, P, H, sig):
    from mp_baselines.planners.stoch_gpmp import StochGPMP
    cfg = configs.config(cfg_name)
    model, obst = cfg['robot'], cfg['obstacles']
    robot, field = oracle_robot(model, cfg['dt']), oracle_field(obst, model)
    start, goal = torch.tensor(cfg['start'], **TA), torch.tensor(cfg['goal'], **TA)
    planner = StochGPMP(robot=robot, n_dof=model.q_dim, n_support_points=H, num_particles_per_goal=P, opt_iters=1,
                        dt=cfg['dt'], start_state=start, multi_goal_states=goal.unsqueeze(0), collision_fields=[field],
                        tensor_args=TA, num_samples=2, initial_particle_means='const_vel', **sig)
    meta = dict(cfg=cfg_name, P=P, H=H, d=model.q_dim, dt=cfg['dt'], **sig)
    save(tag, meta=np.array(repr(meta)), start=start, goal=goal, means0=planner._particle_means)


if __name__ == '__main__':
    torch.set_num_threads(4)
    init = dict(sigma_start_init=1e-3, sigma_goal_init=1e-3, sigma_gp_init=5.0)
    gen_init_stomp('init_random_pm2d', 'C1', P=3, H=16, seed=70, sig=init)
    gen_init_stomp('init_random_panda', 'C4', P=2, H=12, seed=71, sig=dict(init, sigma_gp_init=0.5))
    gen_const_vel('init_const_vel_pm3d', 'C3', P=3, H=16, sig=mg.MODERATE)
    gen_const_vel('init_const_vel_panda', 'C4', P=2, H=64, sig=mg.MODERATE)
