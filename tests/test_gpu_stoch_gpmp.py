"""GPU parity tests (through the C ABI) for the Stoch-GPMP path: K1 sampling, K2 cost, K3 update.

Compared against (i) golden vectors from the unmodified reference planners and (ii) the CPU oracle
on seeded inputs.  Tolerances: 1e-5 relative for costs / weights / trajectories (BASELINE.json
north_star); collision-free flags and argmin indices must be identical.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

from motion_planning_baselines_b200 import configs  # noqa: E402


def T(a):
    if torch.is_tensor(a):
        return a.detach().cpu()
    return torch.as_tensor(np.asarray(a))


def rel_err(a, b):
    a, b = a.detach().double().cpu(), T(b).double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())


def assert_close(a, b, rtol=1e-5, atol=0.0, what=''):
    a, b = a.detach().double().cpu(), T(b).double().cpu()
    err = (a - b).abs()
    ok = err <= atol + rtol * b.abs()
    if not bool(ok.all()):
        idx = (~ok).nonzero()[:5].tolist()
        detail = [(i, float(a[tuple(i)]), float(b[tuple(i)])) for i in idx]
        raise AssertionError(f'{what}: {int((~ok).sum())}/{ok.numel()} off, max rel '
                             f'{float((err / b.abs().clamp_min(1e-30)).max()):.3e} max abs {float(err.max()):.3e}; first (index, got, want): {detail}')


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return dict(device=torch.device('cuda:0'), dtype=torch.float32)


def make_planner(g, dev, cfg_name=None, P=None, S=None, **over):
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import StochGPMP
    from motion_planning_baselines_b200.robots import Robot
    m = g['meta']
    cfg = configs.config(cfg_name or m['cfg'])
    robot = Robot(cfg['robot'], dt=m['dt'], tensor_args=dev)
    field = CollisionField(cfg['obstacles'], tensor_args=dev)
    sig = {k: m[k] for k in m if k.startswith('sigma_') or k in ('temperature', 'step_size')}
    sig.update(over)
    means0 = T(g['means0']).to(**dev)
    planner = StochGPMP(robot=robot, n_dof=m['d'], n_support_points=m['H'], num_particles_per_goal=P or m['P'],
                        opt_iters=1, dt=m['dt'], start_state=T(g['start']).to(**dev),
                        multi_goal_states=T(g['goal']).to(**dev).unsqueeze(0), collision_fields=[field],
                        tensor_args=dev, num_samples=S or m['S'], initial_particle_means=means0.unsqueeze(0), **sig)
    return planner


@pytest.mark.parametrize('name', ['prior_d2_H16', 'prior_d7_H8'])
def test_prior_setup_and_sampling_vs_reference(name, dev):
    from motion_planning_baselines_b200.factors import GPFactor, MultiMPPrior, UnaryFactor
    g = load_golden(name)
    m = g['meta']
    d, H = m['d'], m['H']
    K_s = UnaryFactor(2 * d, m['sig_s'], None, dev).K
    K_g = UnaryFactor(2 * d, m['sig_g'], None, dev).K
    Q = GPFactor(d, m['sig_gp'], m['dt'], H - 1, dev).Q_inv[0]
    prior = MultiMPPrior(H - 1, m['dt'], 2 * d, d, K_s, Q, T(g['start']).to(**dev), means=T(g['means']).to(**dev),
                         K_g_inv=K_g, goal_states=T(g['goal']).to(**dev), tensor_args=dev)
    assert torch.equal(prior.Sigma_inv.cpu(), T(g['Sigma_inv'])), 'precision must be bit-identical'
    assert_close(prior.scale_tril, g['scale_tril'], rtol=1e-6, atol=1e-9, what='scale_tril')
    x = prior.sample(m['S'], eps=T(g['eps']).to(**dev))
    assert_close(x, g['samples'], rtol=1e-5, atol=1e-6, what='samples')
    # const-vel mean when no means are given
    prior0 = MultiMPPrior(H - 1, m['dt'], 2 * d, d, K_s, Q, T(g['start']).to(**dev), K_g_inv=K_g,
                          goal_states=T(g['goal']).to(**dev), tensor_args=dev)
    assert_close(prior0.means, g['const_vel_mean'], rtol=1e-6, atol=1e-7, what='const-vel mean')


STOCH = ['stochgpmp_pm2d_moderate', 'stochgpmp_pm3d_moderate', 'stochgpmp_pm3d_frozen',
         'stochgpmp_panda_moderate', 'stochgpmp_panda_frozen']


@pytest.mark.parametrize('name', STOCH)
def test_cost_terms_and_update_vs_reference_golden(name, dev):
    g = load_golden(name)
    m = g['meta']
    planner = make_planner(g, dev)
    assert torch.equal(planner.Sigma_inv.cpu(), T(g['Sigma_inv']))
    assert_close(planner._sample_dist.scale_tril, g['L'], rtol=1e-6, atol=1e-9, what='L')
    P, S = m['P'], m['S']
    for it in range(m['iters']):
        means = T(g['means0'] if it == 0 else g[f'means{it}']).to(**dev)
        xs = T(g[f'samples{it}']).to(**dev).contiguous()
        # K2 on the reference's own samples: individual terms, then total + IS term
        terms, w = planner.cost.eval(xs, return_invidual_costs_and_weights=True)
        assert_close(torch.stack(terms), g[f'terms{it}'], rtol=1e-5, atol=1e-6, what='cost terms')
        planner._particle_means.copy_(means)
        planner.state_samples.copy_(xs)
        costs = planner._get_costs()
        assert_close(costs, g[f'costs{it}'], rtol=1e-5, what='costs + IS')
        assert int(costs.argmin()) == int(T(g[f'costs{it}']).argmin())
        assert torch.equal(costs.argmin(dim=1).cpu(), T(g[f'costs{it}']).argmin(dim=1)), 'per-particle argmin'
        # K3 on the reference's costs
        planner._update_distribution(T(g[f'costs{it}']).to(**dev), xs)
        assert_close(planner._weights.reshape(P, S), g[f'weights{it}'], rtol=1e-5, atol=1e-30, what='weights')
        assert_close(planner._particle_means, g[f'means{it + 1}'], rtol=1e-5, atol=1e-7, what='updated means')


@pytest.mark.parametrize('name', ['stochgpmp_pm2d_moderate', 'stochgpmp_pm3d_moderate', 'stochgpmp_panda_moderate'])
def test_full_iterations_vs_reference_golden(name, dev):
    """optimize() end to end on the recorded noise (moderate sigma regime: weights not one-hot)."""
    g = load_golden(name)
    m = g['meta']
    planner = make_planner(g, dev)
    for it in range(m['iters']):
        traj = planner.optimize(opt_iters=1, eps=[T(g[f'eps{it}']).to(**dev).contiguous()])
        assert_close(planner.state_samples, g[f'samples{it}'], rtol=1e-5, atol=1e-6, what='samples')
        assert_close(planner.costs, g[f'costs{it}'], rtol=2e-5, what='costs')
        # softmax sensitivity: a relative cost error r moves a weight by exp(r*|c|/T) - 1, so the weight /
        # trajectory tolerance is the 1e-5 cost tolerance propagated through the softmax
        wtol = 4 * 1e-5 * float(np.abs(g[f'costs{it}']).max()) / m['temperature'] + 1e-5
        assert_close(planner._weights.reshape(m['P'], m['S']), g[f'weights{it}'], rtol=wtol, atol=1e-6, what='weights')
        step = float(np.abs(g[f'means{it + 1}'] - (g['means0'] if it == 0 else g[f'means{it}'])).max())
        assert_close(traj, g[f'means{it + 1}'], rtol=1e-5, atol=wtol * step + 1e-6, what='trajectory')
        assert traj.data_ptr() != planner._particle_means.data_ptr(), 'optimize() returns a clone'
        planner._particle_means.copy_(T(g[f'means{it + 1}']).to(**dev))


@pytest.mark.parametrize('cfg_name,P,S,H', [('C1', 5, 16, 64), ('C3', 4, 16, 64), ('C4', 3, 8, 64), ('C3', 3, 7, 33),
                                            ('C4', 2, 5, 19), ('C5', 2, 6, 24)])
def test_cost_eval_vs_oracle_random(cfg_name, P, S, H, dev):
    """Seeded random trajectories near the obstacle sets: costs within 1e-5, flags / argmin identical,
    per-waypoint hinge arithmetic bit-exact for point robots."""
    from motion_planning_baselines_b200.costs import build_gpmp2_cost_composite
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.robots import Robot
    from oracle.build import TA, oracle_field, oracle_robot
    from oracle.costs import CostSpec
    cfg = configs.config(cfg_name)
    model, obst = cfg['robot'], cfg['obstacles']
    d = model.q_dim
    gen = torch.Generator().manual_seed(1234 + H)
    start, goal = T(cfg['start']), T(cfg['goal'])
    w = torch.linspace(0, 1, H).view(1, H, 1)
    line = start * (1 - w) + goal * w
    x = torch.zeros(P * S, H, 2 * d)
    x[..., :d] = line + 0.15 * torch.randn(P * S, H, d, generator=gen).cumsum(1) / np.sqrt(H) + 0.05 * torch.randn(P * S, 1, d, generator=gen)
    x[..., d:] = 0.5 * torch.randn(P * S, H, d, generator=gen)
    sig = dict(sigma_start=1e-2, sigma_gp=1.0, sigma_goal_prior=1e-2, sigma_coll=1e-1)
    spec = CostSpec(oracle_robot(model, cfg['dt']), H, cfg['dt'], start, goal, [oracle_field(obst, model)], tensor_args=TA, **sig)
    ref_terms = torch.stack(spec.terms(x))
    ref_free = spec.collision_free(x)
    robot = Robot(model, dt=cfg['dt'], tensor_args=dev)
    comp = build_gpmp2_cost_composite(robot=robot, n_support_points=H, dt=cfg['dt'], start_state=start.to(**dev),
                                      multi_goal_states=goal.to(**dev).unsqueeze(0), num_particles_per_goal=P,
                                      collision_fields=[CollisionField(obst, tensor_args=dev)], num_samples=S,
                                      tensor_args=dev, **sig)
    xg = x.to(**dev)
    terms, _ = comp.eval(xg, return_invidual_costs_and_weights=True)
    assert_close(torch.stack(terms), ref_terms, rtol=1e-5, atol=1e-6, what='terms')
    total = comp.eval(xg.view(P, S, H, 2 * d))
    assert_close(total, spec.eval(x), rtol=1e-5, what='total')
    free = comp.collision_free(xg)
    assert not ref_free.all(), 'test data should contain colliding trajectories'
    assert torch.equal(free.cpu(), ref_free), 'collision-free flags must be identical'
    assert torch.equal(total.view(P, S).argmin(dim=1).cpu(), spec.eval(x).view(P, S).argmin(dim=1))
    if model.kind == 'point':
        # collision term: every per-waypoint hinge is bit-identical, so the only difference is the
        # order of the final sum over waypoints -> agree to a few ulp
        assert_close(terms[2], ref_terms[2], rtol=5e-7, atol=0.0, what='point-mass collision term')


def test_edge_cases(dev):
    from motion_planning_baselines_b200 import _lib
    from motion_planning_baselines_b200.costs import CostCollision, CostComposite
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.models import ObstacleSet
    from motion_planning_baselines_b200.robots import RobotPointMass
    robot = RobotPointMass(2, tensor_args=dev)
    empty = CollisionField(ObstacleSet(2), tensor_args=dev)            # no primitives at all
    comp = CostComposite(robot, 8, [CostCollision(robot, 8, field=empty, sigma_coll=1.0)], tensor_args=dev)
    x = torch.randn(3, 8, 4, **dev)
    assert torch.equal(comp(x), torch.zeros(3, **dev))
    assert bool(comp.collision_free(x).all())
    assert comp(torch.empty(0, 8, 4, **dev)).shape == (0,)              # empty batch
    with pytest.raises(_lib.MpbError):
        comp(torch.randn(3, 8, 4))                                      # CPU tensor: no fallback
    with pytest.raises(_lib.MpbError):
        comp(torch.randn(3, 9, 4, **dev))                               # wrong horizon
    # a trajectory that sits inside an obstacle at waypoint 0 only is still "free": waypoint 0 is skipped
    one = CollisionField(ObstacleSet(2, sphere_centers=[[0., 0.]], sphere_radii=[0.2]), tensor_args=dev)
    comp1 = CostComposite(robot, 4, [CostCollision(robot, 4, field=one, sigma_coll=1.0)], tensor_args=dev)
    xx = torch.full((1, 4, 4), 5.0, **dev)
    xx[0, 0, :2] = 0.0
    assert float(comp1(xx)) == 0.0 and bool(comp1.collision_free(xx))
    xx[0, 2, :2] = 0.0
    assert abs(float(comp1(xx)) - 0.2) < 1e-7 and not bool(comp1.collision_free(xx))


def test_update_kernel_properties_full_size(dev):
    """Size-independent properties at the BASELINE.json C4 shape (P=512,S=64,H=64,D=14)."""
    from motion_planning_baselines_b200 import _lib
    P, S, H, D = 512, 64, 64, 14
    gen = torch.Generator(device='cuda').manual_seed(0)
    mu = torch.randn(P, H, D, generator=gen, **dev)
    x = mu.unsqueeze(1) + 0.1 * torch.randn(P, S, H, D, generator=gen, **dev)
    cost = 50 * torch.rand(P, S, generator=gen, **dev)
    w = torch.empty(P, S, **dev)
    g = torch.empty(P, H, D, **dev)
    mu2 = mu.clone()
    _lib.check(_lib.lib().mpb_softmax_update(_lib.ptr(cost), _lib.ptr(x), _lib.ptr(mu2), _lib.ptr(w), _lib.ptr(g),
                                             1.0, 0.25, None, P, S, H, D, _lib.stream_ptr()))
    assert_close(w.sum(1), torch.ones(P), rtol=1e-5, what='weights sum to one')
    wr = torch.softmax(-cost.double(), dim=1)
    gr = (wr.view(P, S, 1, 1) * (x.double() - mu.double().unsqueeze(1))).sum(1)
    assert_close(w, wr, rtol=1e-5, atol=1e-30, what='weights vs fp64')
    assert_close(g, gr, rtol=1e-4, atol=1e-6, what='weighted mean vs fp64')
    assert_close(mu2, mu.double() + 0.25 * gr, rtol=1e-5, atol=1e-6, what='updated means vs fp64')
    # a one-hot cost picks exactly that sample (step 1)
    cost1 = torch.full((P, S), 1e6, **dev)
    idx = torch.randint(0, S, (P,), generator=gen, device=dev['device'])
    cost1[torch.arange(P), idx] = 0.0
    mu3 = mu.clone()
    _lib.check(_lib.lib().mpb_softmax_update(_lib.ptr(cost1), _lib.ptr(x), _lib.ptr(mu3), _lib.ptr(w), None,
                                             1.0, 1.0, None, P, S, H, D, _lib.stream_ptr()))
    assert_close(mu3, x[torch.arange(P), idx], rtol=1e-6, atol=1e-6, what='one-hot update')


def test_sampling_properties_full_size(dev):
    """K1 at the C4 shape: zero noise returns the means, linearity in eps, and agreement with a
    float64 matmul on a row subset."""
    from motion_planning_baselines_b200 import _lib
    P, S, M = 512, 64, 896
    gen = torch.Generator(device='cuda').manual_seed(1)
    L = torch.tril(torch.randn(M, M, generator=gen, **dev)) / 30
    mu = torch.randn(P, M, generator=gen, **dev)
    eps = torch.randn(S, P, M, generator=gen, **dev)
    x = torch.empty(P, S, M, **dev)
    lib = _lib.lib()
    _lib.check(lib.mpb_sample_gp(_lib.ptr(L), _lib.ptr(mu), _lib.ptr(torch.zeros_like(eps)), _lib.ptr(x), P, S, M, _lib.stream_ptr()))
    assert torch.equal(x, mu.unsqueeze(1).expand(P, S, M)), 'zero noise must return the means exactly'
    _lib.check(lib.mpb_sample_gp(_lib.ptr(L), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x), P, S, M, _lib.stream_ptr()))
    ref = mu[:4].double().unsqueeze(1) + torch.einsum('ik,spk->psi', L.double(), eps[:, :4].double())
    assert_close(x[:4], ref, rtol=1e-5, atol=1e-5, what='samples vs fp64')
    x2 = torch.empty_like(x)
    _lib.check(lib.mpb_sample_gp(_lib.ptr(L), _lib.ptr(torch.zeros_like(mu)), _lib.ptr(2 * eps), _lib.ptr(x2), P, S, M, _lib.stream_ptr()))
    assert_close(x2, 2 * (x - mu.unsqueeze(1)), rtol=1e-4, atol=1e-5, what='linearity')


@pytest.mark.parametrize('n_obst,seed,noise', [(16, 0, 0.4), (40, 1, 0.4), (4, 2, 0.05)])
def test_link_frame_cull_is_conservative(n_obst, seed, noise, dev, monkeypatch):
    """The packed kernel culls sphere-only obstacle lists in the link frame (cost_eval_packed.cuh::cull_link_local:
    expanded pair test with a slack).  That pass only decides which spheres reach the exact pass, so it must flag a
    superset of the world-frame cull's set: identical collision-free flags, collision term equal up to the order of the
    per-lane sums, and both equal to the oracle.  Dense random sphere fields around the arm (also > 32 primitives)."""
    import os
    from motion_planning_baselines_b200.costs import build_gpmp2_cost_composite
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.models import ObstacleSet
    from motion_planning_baselines_b200.robots import Robot
    from oracle.build import TA, oracle_field, oracle_robot
    from oracle.costs import CostSpec
    cfg = configs.config('C4')
    model = cfg['robot']
    rng = np.random.default_rng(100 + seed)
    centers = np.stack([rng.uniform(-0.8, 0.8, n_obst), rng.uniform(-0.8, 0.8, n_obst), rng.uniform(0.0, 1.1, n_obst)], 1)
    obst = ObstacleSet(3, sphere_centers=centers.tolist(), sphere_radii=rng.uniform(0.03, 0.2, n_obst).tolist(),
                       cutoff_margin=0.05, name='random_spheres')
    H, d, B = 64, 7, 96
    gen = torch.Generator().manual_seed(77 + seed)
    start, goal = T(cfg['start']), T(cfg['goal'])
    w = torch.linspace(0, 1, H).view(1, H, 1)
    x = torch.zeros(B, H, 2 * d)
    x[..., :d] = start * (1 - w) + goal * w + noise * torch.randn(B, H, d, generator=gen).cumsum(1) / np.sqrt(H) + 0.75 * noise * torch.randn(B, 1, d, generator=gen)
    x[..., d:] = 0.5 * torch.randn(B, H, d, generator=gen)
    sig = dict(sigma_start=1e-2, sigma_gp=1.0, sigma_goal_prior=1e-2, sigma_coll=1e-1)
    robot = Robot(model, dt=cfg['dt'], tensor_args=dev)
    comp = build_gpmp2_cost_composite(robot=robot, n_support_points=H, dt=cfg['dt'], start_state=start.to(**dev),
                                      multi_goal_states=goal.to(**dev).unsqueeze(0), num_particles_per_goal=B,
                                      collision_fields=[CollisionField(obst, tensor_args=dev)], num_samples=1,
                                      tensor_args=dev, **sig)
    xg = x.to(**dev)
    res = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('MPB_K2_LOCAL', mode)
        os.putenv('MPB_K2_LOCAL', mode)          # the C side reads it with getenv() at every call
        terms, _ = comp.eval(xg, return_invidual_costs_and_weights=True)
        res[mode] = (torch.stack(terms).cpu(), comp.collision_free(xg).cpu())
    os.unsetenv('MPB_K2_LOCAL')
    assert torch.equal(res['0'][1], res['1'][1]), 'collision-free flags: link-frame cull vs world-frame cull'
    assert_close(res['1'][0], res['0'][0], rtol=2e-6, atol=1e-7, what='terms: link-frame vs world-frame cull')
    spec = CostSpec(oracle_robot(model, cfg['dt']), H, cfg['dt'], start, goal, [oracle_field(obst, model)], tensor_args=TA, **sig)
    assert_close(res['1'][0], torch.stack(spec.terms(x)), rtol=1e-5, atol=1e-6, what='terms vs oracle')
    ref_free = spec.collision_free(x)
    assert not ref_free.all(), 'test data should contain colliding trajectories'
    assert torch.equal(res['1'][1], ref_free), 'collision-free flags vs oracle'
