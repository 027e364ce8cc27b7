"""GPU parity tests (through the C ABI) for the self-collision and workspace-boundary fields:
cost terms and collision-free flags (K2), analytic gradients (K4 CHOMP, K5 GPMP2 linearisation) and a full
Stoch-GPMP iteration with three collision fields, against the CPU oracle (oracle/fields.py) on seeded inputs.
Tolerance: 1e-5 relative (BASELINE.json north_star); point-robot hinge sums and flags bit-identical.
"""
import numpy as np
import pytest
import torch

from test_gpu_stoch_gpmp import T, assert_close

pytestmark = pytest.mark.gpu

from motion_planning_baselines_b200 import configs  # noqa: E402

WS_PANDA = ([-0.55, -0.6, -0.1], [0.7, 0.6, 1.05])
MARGIN_SELF, MARGIN_WS = 0.03, 0.02


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return dict(device=torch.device('cuda:0'), dtype=torch.float32)


def panda_setup(dev, dt=0.1):
    from motion_planning_baselines_b200.fields import CollisionField, SelfCollisionField, WorkspaceBoundaryField
    from motion_planning_baselines_b200.robots import Robot
    from oracle.build import TA, oracle_field, oracle_robot, oracle_self_field, oracle_workspace_field
    cfg = configs.config('C4')
    model = cfg['robot']
    robot = Robot(model, dt=dt, tensor_args=dev)
    selff = SelfCollisionField(model, cutoff_margin=MARGIN_SELF, tensor_args=dev)
    fields = [CollisionField(cfg['obstacles'], tensor_args=dev), selff,
              WorkspaceBoundaryField(*WS_PANDA, cutoff_margin=MARGIN_WS, tensor_args=dev)]
    ofields = [oracle_field(cfg['obstacles'], model), oracle_self_field(selff.pairs, model, MARGIN_SELF),
               oracle_workspace_field(*WS_PANDA, model, MARGIN_WS)]
    return cfg, model, robot, fields, oracle_robot(model, dt, TA), ofields


def random_panda_trajs(model, B, H, gen, spread=1.0):
    lo, hi = torch.tensor(model.q_min), torch.tensor(model.q_max)
    mid, half = (lo + hi) / 2, (hi - lo) / 2
    q = mid + spread * half * (2 * torch.rand(B, H, 7, generator=gen) - 1)
    return torch.cat((q, 0.3 * torch.randn(B, H, 7, generator=gen)), dim=-1)


def composite(robot, H, fields, sigma_coll, dev, weights=None):
    from motion_planning_baselines_b200.costs import CostCollision, CostComposite
    return CostComposite(robot, H, [CostCollision(robot, H, field=f, sigma_coll=sigma_coll, tensor_args=dev) for f in fields],
                         weights_cost_l=weights, tensor_args=dev)


def test_panda_three_fields_cost_terms_vs_oracle(dev):
    from oracle.build import TA
    from oracle.costs import CostSpec
    cfg, model, robot, fields, orobot, ofields = panda_setup(dev)
    B, H = 96, 40                                  # H > 32: two waypoint batches per warp, the second one partial
    gen = torch.Generator().manual_seed(123)
    x = random_panda_trajs(model, B, H, gen)
    cost = composite(robot, H, fields, 1.0, dev, weights=[1.0, 2.0, 0.5])
    terms, w = cost.eval(x.to(**dev), return_invidual_costs_and_weights=True)
    spec = CostSpec(orobot, H, 0.1, torch.zeros(7), None, ofields, sigma_coll=1.0, tensor_args=TA)
    ref = [spec.collision_cost(x, f) for f in ofields]
    for k, name in enumerate(('objects', 'self', 'workspace')):
        assert float(ref[k].max()) > 0, f'test data must violate the {name} field'
        assert_close(terms[k], ref[k], rtol=1e-5, atol=5e-6, what=f'{name} term')
    total = cost.eval(x.to(**dev))
    assert_close(total, 1.0 * ref[0] + 2.0 * ref[1] + 0.5 * ref[2], rtol=1e-5, atol=1e-5, what='weighted total')
    # collision-free flags of the two new fields: identical wherever no hinge sits within rounding distance of its threshold
    q = torch.tensor(cfg['start']) + 0.5 * torch.randn(256, 8, 7, generator=gen)
    q = torch.minimum(torch.maximum(q, torch.tensor(model.q_min)), torch.tensor(model.q_max))
    x2 = torch.cat((q, torch.zeros_like(q)), dim=-1)
    flags = composite(robot, 8, fields[1:], 1.0, dev).collision_free(x2.to(**dev)).cpu()
    spec8 = CostSpec(orobot, 8, 0.1, torch.zeros(7), None, ofields[1:], sigma_coll=1.0, tensor_args=TA)
    link_pos = orobot.fk_map_collision(x2[..., :7])[:, 1:]
    free = spec8.collision_free(x2)
    slack_ws = (ofields[2].sdf(link_pos) - (ofields[2].link_radii + ofields[2].cutoff_margin)).abs().flatten(1).min(1).values
    i, j = ofields[1].pairs[:, 0], ofields[1].pairs[:, 1]
    dist = (link_pos[..., i, :] - link_pos[..., j, :]).norm(dim=-1)
    slack_self = (dist - (ofields[1].link_radii[i] + ofields[1].link_radii[j] + MARGIN_SELF)).abs().flatten(1).min(1).values
    clear = torch.minimum(slack_ws, slack_self) > 1e-5
    assert int(clear.sum()) > 200 and 20 < int(free.sum()) < 236, 'need both free and colliding trajectories'
    assert torch.equal(flags[clear], free[clear]), 'collision-free flags'


@pytest.mark.parametrize('cfg_name', ['C2', 'C3'])
def test_point_robot_workspace_field_bit_exact(cfg_name, dev):
    """Point robots: the workspace hinge is a chain of single subtractions / minima, so per-trajectory sums of the
    field (accumulated in fp64 on both sides) and the flags must be bit-identical."""
    from motion_planning_baselines_b200.fields import CollisionField, WorkspaceBoundaryField
    from motion_planning_baselines_b200.robots import Robot
    from oracle.build import TA, oracle_field, oracle_robot, oracle_workspace_field
    from oracle.costs import CostSpec
    cfg = configs.config(cfg_name)
    model = cfg['robot']
    d = model.q_dim
    B, H = 333, 32            # one waypoint per lane: the fp64 warp sum of the fp32 hinges is exact, rounded once
    lo, hi = [-0.9] * d, [0.85] * d
    gen = torch.Generator().manual_seed(9)
    x = torch.cat((2.2 * torch.rand(B, H, d, generator=gen) - 1.1, torch.randn(B, H, d, generator=gen)), dim=-1)
    x[:50, :, :d] = T(cfg['start']) + 0.03 * (2 * torch.rand(50, H, d, generator=gen) - 1)   # some free trajectories
    robot = Robot(model, dt=cfg['dt'], tensor_args=dev)
    fields = [CollisionField(cfg['obstacles'], tensor_args=dev), WorkspaceBoundaryField(lo, hi, cutoff_margin=0.04, tensor_args=dev)]
    ofields = [oracle_field(cfg['obstacles'], model), oracle_workspace_field(lo, hi, model, 0.04)]
    cost = composite(robot, H, fields, 1.0, dev)
    terms, _ = cost.eval(x.to(**dev), return_invidual_costs_and_weights=True)
    spec = CostSpec(oracle_robot(model, cfg['dt'], TA), H, cfg['dt'], torch.zeros(d), None, ofields, sigma_coll=1.0, tensor_args=TA)
    q = x[..., :d]
    for k in range(2):
        ref = ofields[k].compute_cost(q[:, 1:], q[:, 1:].unsqueeze(-2)).double().sum(1).float()     # fp64 sum over waypoints
        assert float(ref.max()) > 0
        assert torch.equal(terms[k].cpu(), ref), f'field {k} sums must be bit-identical'
    flags = cost.collision_free(x.to(**dev)).cpu()
    free = spec.collision_free(x)
    assert 0 < int(free.sum()) < B
    assert torch.equal(flags, free)


def test_chomp_gradient_three_fields_vs_oracle_autograd(dev):
    from motion_planning_baselines_b200.planners import CHOMP
    from oracle import planners as op
    from oracle.build import TA
    from oracle.costs import CostSpec
    cfg, model, robot, fields, orobot, ofields = panda_setup(dev, dt=0.08)
    P, H, d = 12, 14, 7
    gen = torch.Generator().manual_seed(41)
    x = random_panda_trajs(model, P, H, gen, spread=0.9)
    sigma_coll, w_prior, lr, clip = 0.7, 1e-8, 0.01, 1e9
    weights = [1.5, 2.0, 0.75]
    spec = CostSpec(orobot, H, 0.08, torch.zeros(7), None, ofields, sigma_coll=sigma_coll, tensor_args=TA)
    R = op.chomp_R(H, 0.08, TA)
    cost = composite(robot, H, fields, sigma_coll, dev, weights=weights)
    for k in range(3):          # every field on its own (so that each gradient is checked at its own scale), then all
        ocost = lambda xx, k=k: weights[k] * spec.collision_cost(xx, ofields[k])
        ref = op.chomp_iteration(ocost, x, R, w_prior, lr, clip)
        assert float(ref['grad_raw'][..., :d].abs().max()) > 0, f'field {k} must be active'
        ck = composite(robot, H, [fields[k]], sigma_coll, dev, weights=[weights[k]])
        planner = CHOMP(n_dof=d, n_support_points=H, num_particles_per_goal=P, opt_iters=1, dt=0.08,
                        start_state=x[0, 0, :d].to(**dev), cost=ck, weight_prior_cost=w_prior, step_size=lr, grad_clip=clip,
                        multi_goal_states=x[0, -1, :d].to(**dev).unsqueeze(0), initial_particle_means=x.to(**dev),
                        pos_only=False, tensor_args=dev)
        got = planner.optimize(opt_iters=1)
        grad = (x.to(**dev) - got) / lr
        scale = float(ref['grad'].abs().max())
        assert_close(grad, ref['grad'], rtol=1e-4, atol=1e-4 * scale + 2e-5, what=f'gradient of field {k}')
    ocost = lambda xx: sum(weights[k] * spec.collision_cost(xx, ofields[k]) for k in range(3))
    ref = op.chomp_iteration(ocost, x, R, w_prior, lr, clip)
    planner = CHOMP(n_dof=d, n_support_points=H, num_particles_per_goal=P, opt_iters=1, dt=0.08,
                    start_state=x[0, 0, :d].to(**dev), cost=cost, weight_prior_cost=w_prior, step_size=lr, grad_clip=clip,
                    multi_goal_states=x[0, -1, :d].to(**dev).unsqueeze(0), initial_particle_means=x.to(**dev),
                    pos_only=False, tensor_args=dev)
    got = planner.optimize(opt_iters=1)
    assert_close(got, ref['x'], rtol=1e-5, atol=2e-6, what='updated trajectories')


@pytest.mark.parametrize('robot_kind', ['panda', 'point3d'])
def test_gpmp2_linearisation_with_extra_fields_vs_oracle(robot_kind, dev):
    """err / H_obst rows of every field (mpb_gpmp2_linearize) against autograd through the oracle
    (CostCollision.get_linear_system, cost_functions.py:191-231)."""
    from oracle.build import TA, oracle_field, oracle_robot, oracle_workspace_field
    from oracle.costs import CostSpec
    gen = torch.Generator().manual_seed(17)
    if robot_kind == 'panda':
        cfg, model, robot, fields, orobot, ofields = panda_setup(dev)
        B, H, d = 10, 12, 7
        x = random_panda_trajs(model, B, H, gen, spread=0.9)
        tol = 1e-4
    else:
        from motion_planning_baselines_b200.fields import CollisionField, WorkspaceBoundaryField
        from motion_planning_baselines_b200.robots import Robot
        cfg = configs.config('C3')
        model = cfg['robot']
        B, H, d = 40, 20, 3
        lo, hi = [-0.9, -0.85, -0.8], [0.8, 0.85, 0.9]
        x = torch.cat((2.1 * torch.rand(B, H, d, generator=gen) - 1.05, torch.randn(B, H, d, generator=gen)), dim=-1)
        robot = Robot(model, dt=0.1, tensor_args=dev)
        fields = [CollisionField(cfg['obstacles'], tensor_args=dev), WorkspaceBoundaryField(lo, hi, cutoff_margin=0.05, tensor_args=dev)]
        orobot = oracle_robot(model, 0.1, TA)
        ofields = [oracle_field(cfg['obstacles'], model), oracle_workspace_field(lo, hi, model, 0.05)]
        tol = 1e-5
    cost = composite(robot, H, fields, 1.0, dev)
    err, hobs = cost.linearize_collision(x.to(**dev))
    spec = CostSpec(orobot, H, 0.1, torch.zeros(d), None, ofields, sigma_coll=1.0, tensor_args=TA)
    for k, f in enumerate(ofields):
        xx = x.clone().requires_grad_(True)
        e = spec.collision_errors(xx, f)
        g = torch.autograd.grad(e.sum(), xx)[0]
        assert float(e.detach().max()) > 0
        assert_close(err[k][:, 1:], e.detach(), rtol=1e-5, atol=5e-6, what=f'err of field {k}')
        scale = float(g.abs().max())
        assert_close(hobs[k][:, 1:], -g[:, 1:, :d], rtol=tol, atol=tol * scale, what=f'H_obst of field {k}')
        assert float(hobs[k][:, 0].abs().max()) == 0.0


def test_stoch_gpmp_iteration_with_three_fields_vs_oracle(dev):
    from motion_planning_baselines_b200.planners import StochGPMP
    from oracle import planners as op
    from oracle.build import TA
    from oracle.costs import CostSpec
    cfg, model, robot, fields, orobot, ofields = panda_setup(dev, dt=cfg_dt())
    P, S, H, d = 6, 16, 24, 7
    sig = dict(sigma_start=1e-2, sigma_gp=1.0, sigma_goal_prior=1e-2, sigma_coll=1e-1,
               sigma_start_init=1e-2, sigma_goal_init=1e-2, sigma_gp_init=1.0,
               sigma_start_sample=1e-2, sigma_goal_sample=1e-2, sigma_gp_sample=1.0, temperature=1.0, step_size=0.5)
    start, goal = torch.tensor(cfg['start']), torch.tensor(cfg['goal'])
    torch.manual_seed(3)
    planner = StochGPMP(robot=robot, n_dof=d, n_support_points=H, num_particles_per_goal=P, opt_iters=1, dt=cfg_dt(),
                        start_state=start.to(**dev), multi_goal_states=goal.to(**dev).unsqueeze(0), collision_fields=fields,
                        tensor_args=dev, num_samples=S, **sig)
    spec = CostSpec(orobot, H, cfg_dt(), start, goal, ofields, sigma_start=sig['sigma_start'], sigma_gp=sig['sigma_gp'],
                    sigma_coll=sig['sigma_coll'], sigma_goal_prior=sig['sigma_goal_prior'], tensor_args=TA)
    gen = torch.Generator().manual_seed(8)
    for it in range(2):
        means0 = planner._particle_means.clone().cpu()
        eps = torch.randn(S, P, H * 2 * d, generator=gen)
        traj = planner.optimize(opt_iters=1, eps=[eps.to(**dev)])
        ref = op.stoch_gpmp_iteration(spec, means0, planner._sample_dist.scale_tril.cpu(), planner.Sigma_inv.cpu(), eps,
                                      sig['temperature'], sig['step_size'])
        terms = spec.terms(ref['samples'].reshape(P * S, H, 2 * d))
        assert float(terms[3].max()) > 0 or float(terms[4].max()) > 0, 'the extra fields should be active in this workload'
        # K1 (3xTF32 on tcgen05) is accurate to 7e-6 of the noise amplitude; amplitude here ~1.4
        assert_close(planner.state_samples, ref['samples'], rtol=1e-5, atol=1.5e-5, what='samples')
        assert_close(planner.costs, ref['costs'], rtol=1e-5, atol=1e-3, what='costs')
        assert torch.equal(planner.costs.argmin(1).cpu(), ref['costs'].argmin(1)), 'argmin sample per particle'
        assert_close(traj, ref['means'], rtol=1e-4, atol=1e-5, what='updated means')


def cfg_dt():
    return 5.0 / 64
