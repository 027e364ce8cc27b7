"""Golden fixtures for the SURVEY.md 8(f) "next" rows, again from the UNMODIFIED reference classes:

    gpmp2_interp_*   GPMP2 with n_interpolated_points (cost_functions.py:107-144,191-231; field_factor.py:41-57):
                     dense A / b / K of the first step and the particle means after every step
    extra_costs_*    CostJointLimits, CostSmoothnessCHOMP, CostGPTrajectory (cost_functions.py:317-429), a composite
                     that holds them, and CostComposite.eval(trajs_interpolated=...) (cost_functions.py:70-87)

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only (needs /root/reference):

    python oracle/make_golden_next.py

Kept apart from make_golden.py so that the round-1 fixtures stay byte-identical.
"""
import numpy as np
import torch

from make_golden import TA, configs, oracle_field, oracle_robot, save  # noqa: F401  (also sets sys.path for mp_baselines)


def gen_gpmp2_interp(tag, cfg_name, P, H, n_interp, iters, seed, prm, dt):
    from mp_baselines.planners.gpmp2 import GPMP2
    cfg = configs.config(cfg_name)
    model, obst = cfg['robot'], cfg['obstacles']
    robot, field = oracle_robot(model, dt), oracle_field(obst, model)
    torch.manual_seed(seed)
    start = torch.tensor(cfg['start'], **TA)
    goal = torch.tensor(cfg['goal'], **TA)
    planner = GPMP2(robot=robot, n_dof=model.q_dim, n_support_points=H, n_interpolated_points=n_interp,
                    num_particles_per_goal=P, opt_iters=1,
                    dt=dt, start_state=start, multi_goal_states=goal.unsqueeze(0), collision_fields=[field],
                    step_size=prm['step_size'], sigma_start_init=prm['sigma_start_init'],
                    sigma_goal_init=prm['sigma_goal_init'], sigma_gp_init=prm['sigma_gp_init'],
                    sigma_start_sample=prm['sigma_start_sample'], sigma_goal_sample=prm['sigma_goal_sample'],
                    solver_params=dict(delta=prm['delta'], trust_region=prm['trust_region'], method=prm['method']),
                    sigma_start=prm['sigma_start'], sigma_gp=prm['sigma_gp'], sigma_coll=prm['sigma_coll'],
                    sigma_goal_prior=prm['sigma_goal_prior'], tensor_args=TA)
    # spread the particles so that the interpolated points touch obstacles the support points miss
    gen = torch.Generator().manual_seed(seed + 1)
    d = model.q_dim
    means = planner._particle_means.clone()
    means[..., :d] += 0.15 * torch.randn(P, H, d, generator=gen).cumsum(1) / np.sqrt(H)
    planner._particle_means = means.clone()
    out = dict(means0=means.clone())
    for it in range(iters):
        A, b, K = planner.cost.get_linear_system(planner._particle_means.clone(), n_interpolated_points=n_interp)
        if it == 0:
            A0, b0, K0 = planner.cost.get_linear_system(planner._particle_means.clone(), n_interpolated_points=None)
            out['A0'], out['b0'], out['Kdiag0'] = A, b, torch.diagonal(K, dim1=-2, dim2=-1)
            out['A0_plain'] = A0
            assert float((A - A0).abs().max()) > 0, 'interpolation must change the collision Jacobian in this fixture'
        planner.optimize(opt_iters=1)
        out[f'means{it + 1}'] = planner._particle_means.clone()
        out[f'costs{it}'] = planner.costs
    meta = dict(cfg=cfg_name, P=P, H=H, d=d, dt=dt, iters=iters, n_interp=n_interp, **prm)
    save(tag, meta=np.array(repr(meta)), start=start, goal=goal, **out)


def gen_extra_costs(tag, cfg_name, B, H, seed, dt, spread):
    from mp_baselines.planners.costs.cost_functions import (CostCollision, CostComposite, CostGP, CostGPTrajectory,
                                                            CostJointLimits, CostSmoothnessCHOMP)
    from torch_robotics.torch_planning_objectives.fields.distance_fields import interpolate_points_v1
    cfg = configs.config(cfg_name)
    model, obst = cfg['robot'], cfg['obstacles']
    robot, field = oracle_robot(model, dt), oracle_field(obst, model)
    d = model.q_dim
    gen = torch.Generator().manual_seed(seed)
    lo, hi = torch.tensor(model.q_min), torch.tensor(model.q_max)
    w = torch.linspace(0, 1, H).view(1, H, 1)
    q = (lo + (hi - lo) * torch.rand(B, 1, d, generator=gen)) * (1 - w) + (lo + (hi - lo) * torch.rand(B, 1, d, generator=gen)) * w
    q = q + spread * torch.randn(B, H, d, generator=gen)          # wanders past the joint limits here and there
    x = torch.cat((q, 0.5 * torch.randn(B, H, d, generator=gen)), dim=-1).to(**TA)
    sigma_gp, sigma_coll, eps_lim = 0.7, 0.3, float(np.deg2rad(3))
    jl = CostJointLimits(robot, H, eps=eps_lim, tensor_args=TA)
    sm = CostSmoothnessCHOMP(robot, H, tensor_args=TA)
    gpt = CostGPTrajectory(robot, H, dt, sigma_gp=sigma_gp, tensor_args=TA)
    start = torch.cat((x[0, 0, :d], torch.zeros(d)))
    gp = CostGP(robot, H, start, dt, dict(sigma_start=0.5, sigma_gp=sigma_gp), tensor_args=TA)
    coll = CostCollision(robot, H, field=field, sigma_coll=sigma_coll, tensor_args=TA)
    weights = [1.0, 2.0, 0.5, 0.25]
    comp = CostComposite(robot, H, [gp, coll, jl, gpt], weights_cost_l=weights, tensor_args=TA)
    out = dict(x=x, joint_limits=jl.eval(x), smoothness=sm.eval(x), gp_traj=gpt.eval(x), composite=comp.eval(x))
    assert out['joint_limits'].ndim == 0 and float(out['joint_limits']) > 0, 'CostJointLimits returns a batch-summed scalar'
    x_interp = interpolate_points_v1(x, 2)
    out['composite_interp'] = comp.eval(x, trajs_interpolated=x_interp)
    # reference quirk: the collision term is still evaluated on the support points (q_pos / H_positions of `trajs`
    # are forwarded, cost_functions.py:71,85), so trajs_interpolated changes nothing
    assert torch.equal(out['composite_interp'], out['composite'])
    terms, _ = comp.eval(x, return_invidual_costs_and_weights=True)
    out['term_gp'], out['term_coll'] = terms[0], terms[1]
    meta = dict(cfg=cfg_name, B=B, H=H, d=d, dt=dt, sigma_gp=sigma_gp, sigma_coll=sigma_coll, sigma_start=0.5, eps=eps_lim,
                weights=weights)
    save(tag, meta=np.array(repr(meta)), **out)


if __name__ == '__main__':
    torch.set_num_threads(4)
    # moderate sigmas: the particles are spread far from the GP mean here, and with the reference's default 1e-5
    # sigmas its fp32 dense Cholesky (cond ~1e12) returns steps that are wrong in the first digit
    prm = dict(configs.config('C2')['params']['gpmp2'], sigma_start=1e-2, sigma_gp=0.5, sigma_coll=1e-2, sigma_goal_prior=1e-2)
    gen_gpmp2_interp('gpmp2_interp_pm2d', 'C2', P=3, H=16, n_interp=3, iters=2, seed=60, prm=prm, dt=5 / 64)
    gen_gpmp2_interp('gpmp2_interp_panda', 'C4', P=2, H=8, n_interp=2, iters=1, seed=61, prm=prm, dt=5 / 64)
    gen_extra_costs('extra_costs_pm2d', 'C2', B=6, H=16, seed=70, dt=0.04, spread=0.05)
    gen_extra_costs('extra_costs_panda', 'C4', B=4, H=12, seed=71, dt=5 / 64, spread=0.08)
