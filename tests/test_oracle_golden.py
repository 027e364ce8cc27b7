"""CPU: the oracle restatement against the golden vectors produced by the UNMODIFIED reference
planners (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from motion_planning_baselines_b200 import configs
from oracle import gp_prior, planners
from oracle.build import TA, oracle_field, oracle_robot
from oracle.costs import CostSpec

from conftest import load_golden


def T(a):
    return torch.as_tensor(np.asarray(a))


def assert_close(a, b, rtol=1e-5, atol=0.0, what=''):
    a, b = T(a).double(), T(b).double()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    assert bool((err <= tol).all()), f'{what}: max rel err {(err / b.abs().clamp_min(1e-30)).max():.3e}, max abs {err.max():.3e}'


@pytest.mark.parametrize('name', ['prior_d2_H16', 'prior_d7_H8'])
def test_prior_precision_factor_and_samples(name):
    g = load_golden(name)
    m = g['meta']
    d, H = m['d'], m['H']
    K_s = gp_prior.unary_K(2 * d, m['sig_s'], TA)
    K_g = gp_prior.unary_K(2 * d, m['sig_g'], TA)
    Q = gp_prior.gp_Q_inv(d, m['dt'], m['sig_gp'], TA)
    assert torch.equal(K_s, T(g['K_s'])) and torch.equal(K_g, T(g['K_g'])) and torch.equal(Q, T(g['Q_inv']))
    Sinv = gp_prior.prior_precision(H, d, m['dt'], K_s, Q, K_g, TA)
    assert torch.equal(Sinv, T(g['Sigma_inv'])), 'precision must be bit-identical (same fp64 assembly)'
    L = gp_prior.precision_to_scale_tril(Sinv)
    assert_close(L, g['scale_tril'], rtol=1e-6, atol=1e-9, what='scale_tril')
    x = gp_prior.sample_prior(T(g['means']).reshape(m['P'], -1), T(g['scale_tril']), T(g['eps']))
    assert_close(x.reshape(m['P'], m['S'], H, 2 * d), g['samples'], rtol=1e-5, atol=1e-6, what='samples')
    cv = gp_prior.const_vel_mean(T(g['start']), T(g['goal'])[0], m['dt'], H, d, TA)
    assert_close(cv.reshape(1, -1), g['const_vel_mean'], rtol=1e-6, atol=1e-7, what='const-vel mean')


def _spec(g, H):
    m = g['meta']
    cfg = configs.config(m['cfg'])
    robot = oracle_robot(cfg['robot'], m['dt'])
    field = oracle_field(cfg['obstacles'], cfg['robot'])
    return CostSpec(robot, H, m['dt'], T(g['start']), T(g['goal']), [field],
                    sigma_start=m['sigma_start'], sigma_gp=m['sigma_gp'], sigma_coll=m['sigma_coll'],
                    sigma_goal_prior=m['sigma_goal_prior'], tensor_args=TA)


STOCH = ['stochgpmp_pm2d_moderate', 'stochgpmp_pm3d_moderate', 'stochgpmp_pm3d_frozen',
         'stochgpmp_panda_moderate', 'stochgpmp_panda_frozen']


@pytest.mark.parametrize('name', STOCH)
def test_stoch_gpmp_iteration(name):
    g = load_golden(name)
    m = g['meta']
    spec = _spec(g, m['H'])
    means = T(g['means0'])
    L, Sinv = T(g['L']), T(g['Sigma_inv'])
    for it in range(m['iters']):
        out = planners.stoch_gpmp_iteration(spec, means, L, Sinv, T(g[f'eps{it}']), m['temperature'], m['step_size'])
        assert_close(out['samples'], g[f'samples{it}'], rtol=1e-5, atol=1e-6, what='samples')
        # cost terms and totals on the reference's own samples (isolates cost maths from sampling)
        xs = T(g[f'samples{it}'])
        terms = torch.stack([t.reshape(-1) for t in spec.terms(xs.reshape(-1, *xs.shape[2:]))])
        assert_close(terms, g[f'terms{it}'], rtol=2e-6, atol=1e-6, what='cost terms')
        c = planners.stoch_gpmp_costs(spec, xs, means, Sinv, m['temperature'])
        assert_close(c, g[f'costs{it}'], rtol=1e-5, what='costs + IS term')
        w, _, new = planners.softmax_update(T(g[f'costs{it}']), xs, means, m['temperature'], m['step_size'])
        assert_close(w, g[f'weights{it}'], rtol=1e-5, atol=1e-30, what='weights (reference costs)')
        assert_close(new, g[f'means{it + 1}'], rtol=1e-5, atol=1e-7, what='updated means')
        assert int(out['costs'].argmin()) == int(T(g[f'costs{it}']).argmin())
        means = T(g[f'means{it + 1}'])


@pytest.mark.parametrize('name', ['stochgpmp_panda_h64_moderate', 'stochgpmp_panda_h64_frozen'])
def test_stoch_gpmp_iteration_h64(name):
    """The benchmarked horizon (H = 64, Panda) from the unmodified reference (oracle/make_golden_h64.py).  These
    fixtures do not carry the 3.2 MB factor: the oracle builds its own and is checked on the stored parts."""
    g = load_golden(name)
    m = g['meta']
    d, H = m['d'], m['H']
    spec = _spec(g, H)
    K_s = gp_prior.unary_K(2 * d, m['sigma_start_sample'], TA)
    K_g = gp_prior.unary_K(2 * d, m['sigma_goal_sample'], TA)
    Q = gp_prior.gp_Q_inv(d, m['dt'], m['sigma_gp_sample'], TA)
    Sinv = gp_prior.prior_precision(H, d, m['dt'], K_s, Q, K_g, TA)
    band = torch.stack([torch.nn.functional.pad(torch.diagonal(Sinv, -k), (0, k)) for k in range(2 * d + 1)])
    assert torch.equal(band, T(g['Sinv_band'])), 'precision band must be bit-identical'
    # The factor: cond(Sigma^-1) ~ 2e6, so the reference's fp32 factorisation is only accurate to ~5e-3 at H = 64 and
    # moves by ~1e-2 with the LAPACK thread count (measured) -- the oracle's own factor can only be held to that; the
    # replay below uses the reference's factor (stored as its d per-dof blocks; every other entry is exactly zero).
    L = torch.zeros(H * 2 * d, H * 2 * d)
    for j in range(d):
        L[j::d, j::d] = T(g['L_blocks'])[j]
    L_own = gp_prior.precision_to_scale_tril(Sinv)
    L64 = gp_prior.precision_to_scale_tril(Sinv.double())
    n = float(L64.abs().max())
    assert float((L_own - L64).abs().max()) / n < 3e-2 and float((L - L64).abs().max()) / n < 3e-2
    means = T(g['means0'])
    for it in range(m['iters']):
        out = planners.stoch_gpmp_iteration(spec, means, L, Sinv, T(g[f'eps{it}']), m['temperature'], m['step_size'])
        assert_close(out['samples'], g[f'samples{it}'], rtol=1e-5, atol=1e-6, what='samples')
        xs = T(g[f'samples{it}'])
        terms = torch.stack([t.reshape(-1) for t in spec.terms(xs.reshape(-1, *xs.shape[2:]))])
        assert_close(terms, g[f'terms{it}'], rtol=2e-6, atol=1e-6, what='cost terms')
        c = planners.stoch_gpmp_costs(spec, xs, means, Sinv, m['temperature'])
        assert_close(c, g[f'costs{it}'], rtol=1e-5, what='costs + IS term')
        w, _, new = planners.softmax_update(T(g[f'costs{it}']), xs, means, m['temperature'], m['step_size'])
        assert_close(w, g[f'weights{it}'], rtol=1e-5, atol=1e-30, what='weights (reference costs)')
        assert_close(new, g[f'means{it + 1}'], rtol=1e-5, atol=1e-7, what='updated means')
        means = T(g[f'means{it + 1}'])


@pytest.mark.parametrize('name', ['stomp_pm2d', 'stomp_panda'])
def test_stomp_iteration(name):
    g = load_golden(name)
    m = g['meta']
    cfg = configs.config(m['cfg'])
    robot = oracle_robot(cfg['robot'], m['dt'])
    field = oracle_field(cfg['obstacles'], cfg['robot'])
    spec = CostSpec(robot, m['H'], m['dt'], T(g['start']), None, [field], sigma_coll=m['sigma_coll'], tensor_args=TA)
    cost = lambda x: spec.collision_cost(x, field)
    R = planners.stomp_R(m['H'], m['dt'], m['sigma_spectral'], TA)
    assert_close(R, g['R'], rtol=1e-6, what='R')
    assert_close(gp_prior.precision_to_scale_tril(R), g['L_R'], rtol=1e-4, atol=1e-9, what='L_R')
    means = T(g['means0'])
    for it in range(m['iters']):
        out = planners.stomp_iteration(cost, means, T(g['L_R']), T(g['Sigma']), T(g[f'eps{it}']),
                                       m['temperature'], m['step_size'])
        assert_close(out['samples'], g[f'samples{it}'], rtol=1e-5, atol=1e-6, what='samples')
        assert_close(out['costs'], g[f'costs{it}'], rtol=1e-4, atol=1e-3, what='costs')
        w = torch.softmax(-T(g[f'costs{it}']) / m['temperature'], dim=1)
        assert_close(w, g[f'weights{it}'], rtol=1e-5, atol=1e-30, what='weights')
        assert_close(out['means'], g[f'means{it + 1}'], rtol=1e-4, atol=1e-5, what='means')
        means = T(g[f'means{it + 1}'])


def test_mppi_iteration():
    g = load_golden('mppi_pm2d')
    m = g['meta']
    cfg = configs.config('C1')
    robot = oracle_robot(cfg['robot'], m['dt'])
    field = oracle_field(cfg['obstacles'], cfg['robot'])
    spec = CostSpec(robot, m['T'], m['dt'], T(g['start']), None, [field], sigma_coll=m['sigma_coll'], tensor_args=TA)

    class Ext:
        def eval(self, x):
            return spec.collision_cost(x, field)
    Cov = planners.mppi_cov(m['T'], [0.15, 0.15], 'const_ctrl', TA)
    assert_close(Cov, g['Cov'], rtol=1e-6, what='Cov')
    mean = T(g['mean0'])
    best = float('inf')
    for it in range(m['iters']):
        out = planners.mppi_iteration(mean, T(g['L_ctrl']), T(g['Cov_inv']), T(g[f'eps{it}']), T(g['start']),
                                      T(g['goal']), m['dt'], torch.tensor([-100., -100.]), torch.tensor([100., 100.]),
                                      m['c_weights'], 1.0, 1.0, ext_cost=Ext())
        assert_close(out['controls'], g[f'controls{it}'], rtol=1e-5, atol=1e-6, what='controls')
        assert_close(out['states'], g[f'states{it}'], rtol=1e-5, atol=1e-6, what='states')
        assert_close(out['costs'], g[f'costs{it}'], rtol=1e-5, what='costs')
        assert_close(out['weights'].reshape(-1), g[f'weights{it}'].reshape(-1), rtol=1e-3, atol=1e-7, what='weights')
        assert_close(out['mean'], g[f'mean{it + 1}'], rtol=1e-3, atol=1e-5, what='mean')
        best = min(best, float(out['best_cost']))
        assert_close(best, g[f'best_cost{it}'], rtol=1e-5, what='best cost')
        mean = T(g[f'mean{it + 1}'])


def test_chomp_iterations():
    g = load_golden('chomp_pm2d')
    m = g['meta']
    cfg = configs.config(m['cfg'])
    robot = oracle_robot(cfg['robot'], m['dt'])
    field = oracle_field(cfg['obstacles'], cfg['robot'])
    spec = CostSpec(robot, m['H'], m['dt'], T(g['start']), None, [field], sigma_coll=m['sigma_coll'], tensor_args=TA)
    cost = lambda x: m['cost_weight'] * spec.collision_cost(x, field)
    R = planners.chomp_R(m['H'], m['dt'], TA)
    assert_close(R, g['R'], rtol=1e-6, what='R')
    x = T(g['x0'])
    for it in range(m['iters']):
        out = planners.chomp_iteration(cost, x, R, m['weight_prior_cost'], m['step_size'], m['grad_clip'])
        assert_close(out['x'], g[f'x{it + 1}'], rtol=1e-5, atol=1e-6, what=f'x after iter {it}')
        x = T(g[f'x{it + 1}'])
    assert_close(T(g[f"x{m['iters']}"]), g['x_multi'], rtol=1e-6, atol=1e-7, what='multi-iteration run')


@pytest.mark.parametrize('name', ['gpmp2_pm2d', 'gpmp2_panda'])
def test_gpmp2_steps(name):
    g = load_golden(name)
    m = g['meta']
    spec = _spec(g, m['H'])
    means = T(g['means0'])
    for it in range(m['iters']):
        out = planners.gpmp2_step(spec, means, m['delta'], m['trust_region'], m['step_size'])
        if it == 0:
            assert_close(out['A'][:2], g['A0'], rtol=1e-4, atol=1e-6, what='A')
            assert_close(out['b'], g['b0'], rtol=1e-5, atol=1e-7, what='b')
            assert_close(torch.diagonal(out['K'], dim1=-2, dim2=-1), g['Kdiag0'], rtol=1e-6, what='K')
        assert_close(out['costs'], g[f'costs{it}'], rtol=1e-4, what='costs')
        step = (T(g[f'means{it + 1}']) - means).abs().max()
        assert_close(out['means'], g[f'means{it + 1}'], rtol=1e-3, atol=float(2e-3 * step), what='means')
        means = T(g[f'means{it + 1}'])


# ---------------------------------------------------------------- SURVEY 8(f) "next" rows (oracle/make_golden_next.py)
@pytest.mark.parametrize('name', ['gpmp2_interp_pm2d', 'gpmp2_interp_panda'])
def test_gpmp2_interpolated_linear_system(name):
    from motion_planning_baselines_b200 import configs
    from oracle import planners as op
    from oracle.build import TA, oracle_field, oracle_robot
    from oracle.costs import CostSpec
    g = load_golden(name)
    m = g['meta']
    cfg = configs.config(m['cfg'])
    spec = CostSpec(oracle_robot(cfg['robot'], m['dt']), m['H'], m['dt'], T(g['start']), T(g['goal']),
                    [oracle_field(cfg['obstacles'], cfg['robot'])], sigma_start=m['sigma_start'], sigma_gp=m['sigma_gp'],
                    sigma_coll=m['sigma_coll'], sigma_goal_prior=m['sigma_goal_prior'], tensor_args=TA)
    x = T(g['means0'])
    A, b, K = spec.linear_system(x, n_interpolated_points=m['n_interp'])
    scale = float(np.abs(g['A0']).max())
    assert np.abs(A.numpy() - g['A0']).max() <= 1e-5 * scale
    assert np.abs(g['A0'] - g['A0_plain']).max() > 1e-3 * scale, 'fixture must exercise the interpolation'
    np.testing.assert_allclose(b.numpy(), g['b0'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(torch.diagonal(K, dim1=-2, dim2=-1).numpy(), g['Kdiag0'], rtol=1e-6)
    A0, _, _ = spec.linear_system(x, n_interpolated_points=None)
    assert np.abs(A0.numpy() - g['A0_plain']).max() <= 1e-5 * scale


@pytest.mark.parametrize('name', ['extra_costs_pm2d', 'extra_costs_panda'])
def test_extra_cost_terms(name):
    from motion_planning_baselines_b200 import configs
    from oracle import costs as oc
    from oracle import gp_prior
    from oracle import planners as op
    from oracle.build import TA, oracle_field, oracle_robot
    g = load_golden(name)
    m = g['meta']
    cfg = configs.config(m['cfg'])
    robot = oracle_robot(cfg['robot'], m['dt'])
    x = T(g['x'])
    d, H = m['d'], m['H']
    jl = oc.joint_limits_cost(x, robot.q_min, robot.q_max, m['eps'])
    assert jl.ndim == 0
    np.testing.assert_allclose(jl.numpy(), g['joint_limits'], rtol=1e-5)
    sm = oc.smoothness_chomp_cost(x, op.chomp_R(H, m['dt'], TA))
    np.testing.assert_allclose(sm.numpy(), g['smoothness'], rtol=1e-5, atol=1e-5 * float(np.abs(g['smoothness']).max()))
    gpt = oc.gp_trajectory_cost(x, gp_prior.phi_matrix(d, m['dt'], TA), gp_prior.gp_Q_inv(d, m['dt'], m['sigma_gp'], TA))
    np.testing.assert_allclose(gpt.numpy(), g['gp_traj'], rtol=1e-5)
    spec = oc.CostSpec(robot, H, m['dt'], x[0, 0, :d], None, [oracle_field(cfg['obstacles'], cfg['robot'])],
                       sigma_start=m['sigma_start'], sigma_gp=m['sigma_gp'], sigma_coll=m['sigma_coll'], tensor_args=TA)
    w = m['weights']
    total = w[0] * spec.gp_cost(x) + w[1] * spec.collision_cost(x, spec.fields[0]) + w[2] * jl + w[3] * gpt
    np.testing.assert_allclose(spec.gp_cost(x).numpy(), g['term_gp'], rtol=1e-5)
    np.testing.assert_allclose(spec.collision_cost(x, spec.fields[0]).numpy(), g['term_coll'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(total.numpy(), g['composite'], rtol=1e-5)
    assert np.array_equal(g['composite_interp'], g['composite']), 'trajs_interpolated is ineffective in the reference'
