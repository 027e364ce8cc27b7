"""GPU parity tests (through the C ABI) for the plug points an UNMODIFIED reference planner calls on its duck-typed
objects (SURVEY.md 8b plug points 1-2) and for SURVEY 8f row 4:

  robot.fk_map_collision            (cost_functions.py:50-52)        mpb_fk_spheres / mpb_fk_spheres_vjp
  field.compute_cost + autograd     (field_factor.py:39,52-57)       mpb_field_cost
  cost(x).sum().backward()          (chomp.py:134-139)               mpb_cost_grad
  task.compute_collision, stats     (rrt_base.py:100-101, examples)  mpb_collision_query
  HybridPlanner                     (hybrid_planner.py:33-89)

The checker is the CPU oracle (oracle/robots.py, oracle/fields.py, oracle/costs.py) run through the reference's own call
sequence (FieldFactor.get_error: slice waypoints, compute_cost, autograd Jacobian).  Tolerance 1e-5 relative
(BASELINE.json north_star) unless a comment says otherwise.
"""
import numpy as np
import pytest
import torch

from test_gpu_fields import MARGIN_SELF, MARGIN_WS, WS_PANDA, composite, panda_setup, random_panda_trajs
from test_gpu_stoch_gpmp import T, assert_close

pytestmark = pytest.mark.gpu

from motion_planning_baselines_b200 import configs  # noqa: E402


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return dict(device=torch.device('cuda:0'), dtype=torch.float32)


def field_factor_get_error(robot, field, x, d, calc_jacobian=True):
    """The reference's FieldFactor.get_error call sequence (field_factor.py:17-57) on duck-typed robot / field objects:
    q_pos = x[..., :d]; link_pos = robot.fk_map_collision(q_pos); error = field.compute_cost(q_pos[:, 1:], link_pos[:, 1:]);
    H = -d error.sum() / d q_pos[:, 1:]."""
    q_pos = x[..., :d].detach().clone().requires_grad_(calc_jacobian)
    link_pos = robot.fk_map_collision(q_pos)
    error = field.compute_cost(q_pos[:, 1:], link_pos[:, 1:], obstacle_spheres=None, trajs_interp=None).reshape(x.shape[0], -1)
    if not calc_jacobian:
        return error.detach(), None
    H = -torch.autograd.grad(error.sum(), q_pos)[0][:, 1:]
    field.zero_grad()
    return error.detach(), H


def test_panda_fk_map_collision_and_vjp(dev):
    cfg, model, robot, fields, orobot, ofields = panda_setup(dev)
    gen = torch.Generator().manual_seed(5)
    x = random_panda_trajs(model, 37, 9, gen)
    q = x[..., :7]
    ref = orobot.fk_map_collision(q)
    got = robot.fk_map_collision(q.to(**dev))
    assert got.shape == ref.shape == (37, 9, 50, 3)
    assert_close(got, ref, rtol=0, atol=2e-6, what='sphere centres')          # sincosf / FMA vs torch CPU: ~1e-7
    w = torch.randn(37, 9, 50, 3, generator=gen)
    qr = q.clone().requires_grad_(True)
    gref = torch.autograd.grad((orobot.fk_map_collision(qr) * w).sum(), qr)[0]
    qg = q.to(**dev).requires_grad_(True)
    ggot = torch.autograd.grad((robot.fk_map_collision(qg) * w.to(**dev)).sum(), qg)[0]
    assert_close(ggot, gref, rtol=1e-5, atol=2e-5, what='FK vjp')


def test_point_robot_fk_is_a_view(dev):
    from motion_planning_baselines_b200.robots import RobotPointMass
    robot = RobotPointMass(3, radius=0.01, tensor_args=dev)
    q = torch.randn(4, 5, 3, **dev)
    lp = robot.fk_map_collision(q)
    assert lp.shape == (4, 5, 1, 3) and lp.data_ptr() == q.data_ptr()


def test_panda_field_factor_sequence_all_field_kinds(dev):
    """compute_cost + its autograd Jacobian through FK, per field kind, in the reference's call order."""
    cfg, model, robot, fields, orobot, ofields = panda_setup(dev)
    for f in fields:
        f.bind_robot(robot)
    gen = torch.Generator().manual_seed(77)
    B, H = 24, 12
    x = random_panda_trajs(model, B, H, gen)
    for k, name in enumerate(('objects', 'self', 'workspace')):
        e_ref, H_ref = field_factor_get_error(orobot, ofields[k], x, 7)
        e_got, H_got = field_factor_get_error(robot, fields[k], x.to(**dev), 7)
        assert e_got.shape == (B, H - 1)
        assert float(e_ref.max()) > 0, f'{name}: test data must collide'
        assert_close(e_got, e_ref, rtol=1e-5, atol=2e-6, what=f'{name} error')
        assert_close(H_got, H_ref, rtol=2e-5, atol=2e-5, what=f'{name} Jacobian')
    # no gradient requested -> no gradient buffer, same values
    e_ng, _ = field_factor_get_error(robot, fields[0], x.to(**dev), 7, calc_jacobian=False)
    e_g, _ = field_factor_get_error(robot, fields[0], x.to(**dev), 7)
    assert torch.equal(e_ng, e_g)


@pytest.mark.parametrize('cfg_name', ['C2', 'C3'])
def test_point_robot_compute_cost_bit_exact(cfg_name, dev):
    """Point robots: hinge values are produced by separately rounded operations in the oracle's order -> bit-identical."""
    from motion_planning_baselines_b200.fields import CollisionField, WorkspaceBoundaryField
    from motion_planning_baselines_b200.robots import Robot
    from oracle.build import TA, oracle_field, oracle_robot, oracle_workspace_field
    cfg = configs.config(cfg_name)
    model = cfg['robot']
    d = model.q_dim
    robot = Robot(model, dt=cfg['dt'], tensor_args=dev)
    orobot = oracle_robot(model, cfg['dt'], TA)
    lo, hi = [-0.9] * d, [0.85] * d
    fields = [CollisionField(cfg['obstacles'], tensor_args=dev, robot=robot),
              WorkspaceBoundaryField(lo, hi, cutoff_margin=0.04, tensor_args=dev, robot=robot)]
    ofields = [oracle_field(cfg['obstacles'], model), oracle_workspace_field(lo, hi, model, 0.04)]
    gen = torch.Generator().manual_seed(11)
    x = torch.cat((2.2 * torch.rand(64, 20, d, generator=gen) - 1.1, torch.randn(64, 20, d, generator=gen)), dim=-1)
    for k in range(2):
        e_ref, H_ref = field_factor_get_error(orobot, ofields[k], x, d)
        e_got, H_got = field_factor_get_error(robot, fields[k], x.to(**dev), d)
        assert float(e_ref.max()) > 0
        assert torch.equal(e_got.cpu(), e_ref), f'field {k}: hinge values must be bit-identical'
        assert_close(H_got, H_ref, rtol=1e-5, atol=1e-6, what=f'field {k} Jacobian')


def test_unbound_field_raises(dev):
    from motion_planning_baselines_b200 import _lib
    from motion_planning_baselines_b200.fields import CollisionField
    cfg = configs.config('C2')
    f = CollisionField(cfg['obstacles'], tensor_args=dev)
    with pytest.raises(_lib.MpbError):
        f.compute_cost(None, torch.zeros(3, 1, 2, **dev))


@pytest.mark.parametrize('cfg_name', ['C2', 'C3', 'C4'])
def test_cost_object_backward_matches_oracle_autograd(cfg_name, dev):
    """costs = cost(x); costs.sum().backward()  -- the reference CHOMP loop (chomp.py:134-139) -- on the fused
    GPMP2 composite (start + GP + goal + collision), against autograd through the oracle."""
    from motion_planning_baselines_b200.costs import build_gpmp2_cost_composite
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.robots import Robot
    from oracle.build import TA, oracle_field, oracle_robot
    from oracle.costs import CostSpec
    cfg = configs.config(cfg_name)
    model = cfg['robot']
    d = model.q_dim
    H, B, dt = 16, 40, 0.1
    sig = dict(sigma_start=0.05, sigma_gp=0.7, sigma_coll=0.3, sigma_goal_prior=0.08)
    robot = Robot(model, dt=dt, tensor_args=dev)
    start, goal = torch.tensor(cfg['start']), torch.tensor(cfg['goal'])
    cost = build_gpmp2_cost_composite(robot=robot, n_support_points=H, dt=dt, start_state=start.to(**dev),
                                      multi_goal_states=goal.to(**dev).unsqueeze(0), num_particles_per_goal=B,
                                      collision_fields=[CollisionField(cfg['obstacles'], tensor_args=dev)],
                                      tensor_args=dev, **sig)
    spec = CostSpec(oracle_robot(model, dt, TA), H, dt, start, goal, [oracle_field(cfg['obstacles'], model)], tensor_args=TA, **sig)
    gen = torch.Generator().manual_seed(3)
    if model.kind == 'chain':
        x = random_panda_trajs(model, B, H, gen, spread=0.9)
    else:
        x = torch.cat((1.1 * (2 * torch.rand(B, H, d, generator=gen) - 1), torch.randn(B, H, d, generator=gen)), dim=-1)
    wout = torch.rand(B, generator=gen) + 0.5
    xr = x.clone().requires_grad_(True)
    cr = spec.eval(xr)
    (cr * wout).sum().backward()
    xg = x.to(**dev).requires_grad_(True)
    cg = cost(xg)
    assert cg.requires_grad
    (cg * wout.to(**dev)).sum().backward()
    assert_close(cg, cr, rtol=1e-5, atol=1e-5, what='costs')
    # gradient entries are sums of terms of very different size (GP ~1e2, collision ~1e1): compare against the
    # largest entry of the same waypoint row, the scale at which fp32 rounding acts on both sides
    scale = xr.grad.abs().amax(dim=-1, keepdim=True).clamp_min(1e-3)
    err = ((xg.grad.cpu() - xr.grad).abs() / scale).max()
    assert float(err) < 2e-5, f'gradient: max scaled error {float(err):.2e}'
    assert float(xr.grad[..., :d].abs().max()) > 0
    # no_grad / plain tensors keep the single-launch forward path
    with torch.no_grad():
        assert not cost(xg).requires_grad


def test_reference_chomp_loop_on_fused_cost_matches_fused_chomp(dev):
    """The reference CHOMP iteration written out in torch (chomp.py:127-169) on OUR cost object (autograd through
    mpb_cost_grad) and the fused K4 kernel must walk the same trajectory."""
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import CHOMP
    from motion_planning_baselines_b200.robots import Robot
    from oracle import planners as op
    cfg = configs.config('C2')
    model = cfg['robot']
    P, H, d, dt = 32, 24, 2, 0.05
    robot = Robot(model, dt=dt, tensor_args=dev)
    cost = composite(robot, H, [CollisionField(cfg['obstacles'], tensor_args=dev)], 0.5, dev, weights=[3.0])
    gen = torch.Generator().manual_seed(8)
    x0 = torch.cat((0.9 * (2 * torch.rand(P, H, d, generator=gen) - 1), 0.1 * torch.randn(P, H, d, generator=gen)), dim=-1)
    w_prior, lr, clip, iters = 1e-6, 0.02, 0.5, 5
    planner = CHOMP(n_dof=d, n_support_points=H, num_particles_per_goal=P, opt_iters=iters, dt=dt,
                    start_state=x0[0, 0, :d].to(**dev), cost=cost, weight_prior_cost=w_prior, step_size=lr, grad_clip=clip,
                    initial_particle_means=x0.to(**dev), tensor_args=dev)
    fused = planner.optimize(opt_iters=iters)
    R = op.chomp_R(H, dt, dev)
    x = x0.to(**dev)
    for _ in range(iters):
        out = op.chomp_iteration(cost, x, R, w_prior, lr, clip)      # the reference loop body, on the CUDA cost object
        x = out['x']
    assert float((x - x0.to(**dev)).abs().max()) > 1e-3
    assert_close(fused, x, rtol=1e-5, atol=1e-6, what='trajectories after 5 CHOMP iterations')


def test_planning_task_queries_and_statistics(dev):
    from motion_planning_baselines_b200.fields import CollisionField, SelfCollisionField, WorkspaceBoundaryField
    from motion_planning_baselines_b200.task import PlanningTask
    cfg, model, robot, fields, orobot, ofields = panda_setup(dev)
    task = PlanningTask(robot, fields)
    gen = torch.Generator().manual_seed(21)
    q = random_panda_trajs(model, 600, 1, gen)[:, 0, :7]
    link = orobot.fk_map_collision(q)
    err_ref = sum(f.compute_cost(q, link) for f in ofields)
    err = task.compute_collision_cost(q.to(**dev))
    assert_close(err, err_ref, rtol=1e-5, atol=5e-6, what='state collision cost')
    clear = (err_ref == 0) | (err_ref > 1e-4)
    in_coll = task.compute_collision(q.to(**dev)).cpu()
    assert 50 < int(in_coll.sum()) < 590
    assert torch.equal(in_coll[clear], (err_ref > 0)[clear])
    # [B,H,D] trajectories: velocities ride along in the rows (row stride D)
    x = random_panda_trajs(model, 40, 6, gen, spread=0.5)
    x[:10, :, :7] = torch.tensor(cfg['start'])
    coll = task.compute_collision(x.to(**dev)).cpu()
    link = orobot.fk_map_collision(x[..., :7])
    ref = sum(f.compute_cost(x[..., :7], link) for f in ofields) > 0
    assert torch.equal(coll, ref)
    assert abs(task.compute_fraction_free_trajs(x.to(**dev)) - float((~ref.any(1)).float().mean())) < 1e-7
    assert abs(task.compute_collision_intensity_trajs(x.to(**dev)) - float(ref.float().mean())) < 1e-7
    assert task.compute_success_free_trajs(x.to(**dev)) == int(bool((~ref.any(1)).any()))
    tc, tf = task.get_trajs_collision_and_free(x.to(**dev))
    assert (0 if tc is None else tc.shape[0]) + (0 if tf is None else tf.shape[0]) == 40
    # rejection sampler: everything it returns is collision-free
    g = torch.Generator(device=dev['device']).manual_seed(1)
    qs = task.random_coll_free_q(100, max_samples=500, generator=g)
    assert qs.shape == (100, 7) and not bool(task.compute_collision(qs).any())


def test_hybrid_planner_seeds_stoch_gpmp(dev):
    """HybridPlanner with a stand-in sample-based planner (the RRT tree search is out of scope): seeds of different
    lengths (and one failure -> straight line) are re-sampled to H waypoints, handed to reset(), and optimised."""
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import HybridPlanner, StochGPMP
    from motion_planning_baselines_b200.robots import Robot
    cfg = configs.config('C1')
    model = cfg['robot']
    H, P, dt = 32, 3, 0.04
    start, goal = torch.tensor(cfg['start'], **dev), torch.tensor(cfg['goal'], **dev)

    class StubSampler:
        start_state_pos, goal_state_pos = start, goal

        def optimize(self, refill_samples_buffer=False, debug=False, **kw):
            mid1 = torch.tensor([[-0.8, 0.6]], **dev)
            mid2 = torch.tensor([[0.0, -0.9], [0.7, -0.2]], **dev)
            return [torch.cat((start[None], mid1, goal[None])), None, torch.cat((start[None], mid2, goal[None]))]

    robot = Robot(model, dt=dt, tensor_args=dev)
    opt = StochGPMP(robot=robot, n_dof=2, n_support_points=H, num_particles_per_goal=P, opt_iters=3, dt=dt, start_state=start,
                    multi_goal_states=goal.unsqueeze(0), collision_fields=[CollisionField(cfg['obstacles'], tensor_args=dev)],
                    tensor_args=dev, num_samples=16, sigma_start=1e-2, sigma_gp=1.0, sigma_goal_prior=1e-2, sigma_coll=1e-1,
                    sigma_start_init=1e-2, sigma_goal_init=1e-2, sigma_gp_init=1.0, sigma_start_sample=1e-2,
                    sigma_goal_sample=1e-2, sigma_gp_sample=1.0, temperature=1.0, step_size=0.5)
    hybrid = HybridPlanner(StubSampler(), opt, tensor_args=dev)
    iters = hybrid.optimize(return_iterations=True)
    assert iters.shape == (4, P, H, 4)
    seed = iters[0].cpu()
    # seed 0: two straight segments through mid1, uniform in the waypoint index; seed 1: the straight line
    u = np.linspace(0, 2, H)
    knots = np.stack([T(start).numpy(), np.array([-0.8, 0.6], np.float32), T(goal).numpy()])
    want0 = np.stack([np.interp(u, [0, 1, 2], knots[:, k]) for k in range(2)], axis=1)
    assert np.abs(seed[0, :, :2].numpy() - want0).max() < 1e-6
    line = T(start).numpy() + np.linspace(0, 1, H)[:, None] * (T(goal) - T(start)).numpy()
    assert np.abs(seed[1, :, :2].numpy() - line).max() < 1e-6
    v = (T(goal) - T(start)) / (H * dt)
    assert torch.allclose(seed[:, 1:-1, 2:], v.expand(P, H - 2, 2), atol=1e-6) and float(seed[:, [0, -1], 2:].abs().max()) == 0
    assert float((iters[-1] - iters[0]).abs().max()) > 0
    assert torch.equal(hybrid.optimize(), opt.get_traj())
