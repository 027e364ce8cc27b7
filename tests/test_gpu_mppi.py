"""GPU parity tests (through the C ABI) for MPPI and for the record-based (multi-CTA / multi-GPU) update.

Golden vectors come from the unmodified reference MPPI (tests/golden/mppi_pm2d.npz, oracle/make_golden.py);
seeded cases are checked against the CPU oracle.  Tolerance 1e-5 relative unless derived otherwise; argmin identical.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_gpu_stoch_gpmp import T, assert_close

pytestmark = pytest.mark.gpu

from motion_planning_baselines_b200 import _lib, configs  # noqa: E402


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return dict(device=torch.device('cuda:0'), dtype=torch.float32)


def _mppi(cfg, N, Tn, dev, sigma_coll, c_weights, control_std, split=None, cov='const_ctrl'):
    from motion_planning_baselines_b200.costs import CostCollision, CostComposite
    from motion_planning_baselines_b200.dynamics import PointParticleDynamics
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import MPPI
    from motion_planning_baselines_b200.robots import Robot
    d = cfg['robot'].q_dim
    robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
    field = CollisionField(cfg['obstacles'], tensor_args=dev)
    cost = CostComposite(robot, Tn, [CostCollision(robot, Tn, field=field, sigma_coll=sigma_coll, tensor_args=dev)], tensor_args=dev)
    system = PointParticleDynamics(rollout_steps=Tn, control_dim=d, state_dim=d, dt=cfg['dt'], discount=1.,
                                   goal_state=T(cfg['goal']), ctrl_min=[-100.] * d, ctrl_max=[100.] * d,
                                   c_weights=c_weights, tensor_args=dev)
    planner = MPPI(system, num_ctrl_samples=N, rollout_steps=Tn, opt_iters=1, control_std=control_std, temp=1., step_size=1.,
                   cov_prior_type=cov, tensor_args=dev, sample_split=split)
    return planner, cost


def test_mppi_vs_reference_golden(dev):
    g = load_golden('mppi_pm2d')
    m = g['meta']
    cfg = configs.config('C1')
    planner, cost = _mppi(cfg, m['N'], m['T'], dev, m['sigma_coll'], m['c_weights'], [0.15, 0.15])
    assert_close(planner.ctrl_dist.Cov, g['Cov'], rtol=1e-6, what='Cov')
    assert_close(planner.Cov_inv, g['Cov_inv'], rtol=1e-4, atol=1e-3, what='Cov_inv')
    assert_close(planner.ctrl_dist.scale_tril, g['L_ctrl'], rtol=1e-5, atol=1e-8, what='L_ctrl')
    # use the reference's own constants so that the comparison isolates the hot path
    planner.Cov_inv.copy_(T(g['Cov_inv']).to(**dev))
    planner.ctrl_dist.scale_tril.copy_(T(g['L_ctrl']).to(**dev))
    obs = dict(state=T(g['start']).to(**dev), goal_state=T(g['goal']).to(**dev), cost=cost)
    best = float('inf')
    for it in range(m['iters']):
        planner._mean.copy_(T(g['mean0'] if it == 0 else g[f'mean{it}']).to(**dev))
        planner.update_ctrl_dist()      # controls are sampled around ctrl_dist.mu (as in the reference)
        U, X, c = planner.optimize(opt_iters=1, eps=[T(g[f'eps{it}']).to(**dev).contiguous()], **obs)
        assert_close(U, g[f'controls{it}'], rtol=1e-5, atol=1e-6, what='controls')
        assert_close(X, g[f'states{it}'], rtol=1e-5, atol=1e-6, what='states')
        assert_close(c, g[f'costs{it}'], rtol=1e-5, what='costs')
        assert int(c.argmin()) == int(T(g[f'costs{it}']).argmin())
        # softmax sensitivity: a relative cost error r moves a weight by exp(r |c| / temp) - 1
        wtol = 4 * 1e-5 * float(np.abs(g[f'costs{it}']).max()) + 1e-5
        assert_close(planner.weights.reshape(-1), g[f'weights{it}'].reshape(-1), rtol=wtol, atol=1e-7, what='weights')
        step = float(np.abs(g[f'mean{it + 1}'] - (g['mean0'] if it == 0 else g[f'mean{it}'])).max())
        assert_close(planner._mean, g[f'mean{it + 1}'], rtol=1e-5, atol=wtol * step + 1e-6, what='mean')
        best = min(best, float(T(g[f'costs{it}']).min()))
        assert_close(planner.best_cost, g[f'best_cost{it}'], rtol=1e-5, what='best cost')
        assert_close(planner.best_traj, g[f'best_traj{it}'], rtol=1e-5, atol=1e-6, what='best trajectory')
        # update kernels on the reference's own costs: weights / mean to 1e-5
        planner._mean.copy_(T(g['mean0'] if it == 0 else g[f'mean{it}']).to(**dev))
        planner.update_ctrl_dist()      # controls are sampled around ctrl_dist.mu (as in the reference)
        planner._xu[..., 2:] = T(g[f'controls{it}']).to(**dev)
        planner.update_controller(T(g[f'costs{it}']).to(**dev).contiguous())
        assert_close(planner.weights.reshape(-1), g[f'weights{it}'].reshape(-1), rtol=1e-5, atol=1e-30, what='weights (reference costs)')
        assert_close(planner._mean, g[f'mean{it + 1}'], rtol=1e-5, atol=1e-6, what='mean (reference costs)')


@pytest.mark.parametrize('cfg_name,N,Tn', [('C1', 300, 24), ('C5', 200, 16), ('C5', 257, 64)])
def test_mppi_iteration_vs_oracle(cfg_name, N, Tn, dev):
    from oracle import planners as op
    from oracle.build import TA, oracle_field, oracle_robot
    from oracle.costs import CostSpec
    cfg = configs.config(cfg_name)
    d = cfg['robot'].q_dim
    cw = dict(pos=1., vel=1., ctrl=1., pos_T=1000., vel_T=0.)
    std = [0.15 + 0.01 * i for i in range(d)]
    sigma_coll = 1e-1
    planner, cost = _mppi(cfg, N, Tn, dev, sigma_coll, cw, std)
    spec = CostSpec(oracle_robot(cfg['robot'], cfg['dt']), Tn, cfg['dt'], T(cfg['start']), None,
                    [oracle_field(cfg['obstacles'], cfg['robot'])], sigma_coll=sigma_coll, tensor_args=TA)

    class Ext:
        def eval(self, x):
            return spec.collision_cost(x, spec.fields[0])
    gen = torch.Generator().manual_seed(9)
    mean = 0.3 * torch.randn(Tn, d, generator=gen)
    start, goal = T(cfg['start']), T(cfg['goal'])
    lo, hi = torch.full((d,), -100.), torch.full((d,), 100.)
    obs = dict(state=start.to(**dev), goal_state=goal.to(**dev), cost=cost)
    for it in range(2):
        eps = torch.randn(d, N, Tn, generator=gen)
        ref = op.mppi_iteration(mean, planner.ctrl_dist.scale_tril.cpu(), planner.Cov_inv.cpu(), eps, start, goal, cfg['dt'], lo, hi,
                                cw, 1.0, 1.0, ext_cost=Ext())
        planner._mean.copy_(mean.to(**dev))
        planner.update_ctrl_dist()      # controls are sampled around ctrl_dist.mu (as in the reference)
        U, X, c = planner.optimize(opt_iters=1, eps=[eps.to(**dev)], **obs)
        assert_close(U, ref['controls'], rtol=1e-5, atol=1e-6, what='controls')
        assert_close(X, ref['states'], rtol=1e-5, atol=1e-6, what='states')
        assert_close(c, ref['costs'], rtol=2e-5, what='costs')
        assert int(c.argmin()) == int(ref['argmin'])
        # update on the oracle's costs
        planner._mean.copy_(mean.to(**dev))
        planner.update_ctrl_dist()      # controls are sampled around ctrl_dist.mu (as in the reference)
        planner.update_controller(ref['costs'].to(**dev).contiguous())
        assert_close(planner.weights.reshape(-1), ref['weights'].reshape(-1), rtol=1e-5, atol=1e-30, what='weights')
        assert_close(planner._mean, ref['mean'], rtol=1e-5, atol=1e-6, what='mean')
        assert int(planner._best[1][0]) == int(ref['argmin'])
        mean = ref['mean']


@pytest.mark.parametrize('P,S,H,D,c0,Dw,sigma', [(3, 1000, 16, 4, 0, 4, False), (1, 20000, 64, 14, 7, 7, False),
                                                 (2, 5000, 32, 4, 0, 4, True), (5, 7, 4, 6, 0, 6, False)])
def test_split_update_vs_single_kernel_and_fp64(P, S, H, D, c0, Dw, sigma, dev):
    """Record-based update == one-CTA-per-particle kernel == float64 torch, including the emulated 3-rank split
    (records of three sample blocks concatenated in rank order, exactly what the all-gather delivers)."""
    from motion_planning_baselines_b200.update import SampleSplit, split_softmax_update
    from oracle import planners as op
    lib, st = _lib.lib(), _lib.stream_ptr()
    gen = torch.Generator(device='cuda').manual_seed(S)
    mu = torch.randn(P, H, Dw, generator=gen, **dev)
    x = torch.randn(P, S, H, D, generator=gen, **dev)
    cost = 30 * torch.rand(P, S, generator=gen, **dev)
    cost[0, S // 2] = cost[0, S - 1] = -3.0                       # tie of the minimum: first index must win
    SigmaR = torch.rand(H, H, generator=gen, **dev) if sigma else None
    temp, step = 0.8, 0.4
    xw = x[..., c0:c0 + Dw].double()
    w64 = torch.softmax(-cost.double() / temp, dim=1)
    g64 = (w64.view(P, S, 1, 1) * (xw - mu.double().unsqueeze(1))).sum(1)
    mu64 = mu.double() + step * (SigmaR.double() @ g64 if sigma else g64)
    mu1 = mu.clone()
    r = split_softmax_update(cost, x, mu1, temp, step, H, D, c0=c0, Dw=Dw, SigmaR=SigmaR, want_grad=True)
    assert_close(r['weights'], w64, rtol=1e-5, atol=1e-30, what='weights')
    assert_close(r['grad'], g64, rtol=1e-4, atol=1e-6, what='weighted mean')
    assert_close(mu1, mu64, rtol=1e-5, atol=2e-5 if sigma else 1e-6, what='updated means')   # SigmaR @ g sums 32 O(1) terms
    assert torch.equal(r['best_idx'].cpu().long(), cost.argmin(dim=1).cpu()), 'first-occurrence argmin'
    assert torch.equal(r['best_cost'], cost.min(dim=1).values)
    if c0 == 0 and Dw == D and S <= 40000:
        mu2, w2 = mu.clone(), torch.empty(P, S, **dev)
        _lib.check(lib.mpb_softmax_update(_lib.ptr(cost), _lib.ptr(x), _lib.ptr(mu2), _lib.ptr(w2), None, temp, step,
                                          _lib.ptr(SigmaR), P, S, H, D, st))
        assert_close(mu1, mu2, rtol=1e-5, atol=2e-5 if sigma else 1e-6, what='split vs single-CTA kernel')
    # emulated 3-rank sample split on one GPU
    REC = lib.mpb_softmax_record_len(H, Dw)
    recs, n_chunks = [], 2
    for rnk in range(3):
        off, cnt = SampleSplit(rank=rnk, world=3).local_slice(S)
        rec = torch.empty(n_chunks, P, REC, **dev)
        cl, xl = cost[:, off:off + cnt].contiguous(), x[:, off:off + cnt].contiguous()   # named: must outlive the launch
        _lib.check(lib.mpb_softmax_partial(_lib.ptr(cl), _lib.ptr(xl),
                                           _lib.ptr(mu), _lib.ptr(rec), temp, P, cnt, H, D, c0, Dw, n_chunks, off, st))
        recs.append(rec)
    rec_all = torch.cat(recs)
    mu3 = mu.clone()
    lse, bc, bi = torch.empty(P, 2, **dev), torch.empty(P, **dev), torch.empty(P, device=dev['device'], dtype=torch.int32)
    _lib.check(lib.mpb_softmax_combine(_lib.ptr(rec_all), rec_all.shape[0], _lib.ptr(mu3), None, _lib.ptr(lse), _lib.ptr(bc),
                                       _lib.ptr(bi), step, _lib.ptr(SigmaR), P, H, Dw, st))
    assert_close(mu3, mu64, rtol=1e-5, atol=2e-5 if sigma else 1e-6, what='3-rank emulation')
    assert torch.equal(bi.cpu().long(), cost.argmin(dim=1).cpu())
    # the CUDA records agree with the oracle's record definition (which the gloo CPU test exercises)
    off, cnt = SampleSplit(rank=1, world=3).local_slice(S)
    ref = op.partial_record(cost[:, off:off + cnt].cpu(), x[:, off:off + cnt, :, c0:c0 + Dw].reshape(P, cnt, -1).cpu(),
                            mu.reshape(P, -1).cpu(), temp, sample_offset=off)
    m_c, Z_c = recs[1][:, :, 0].double().cpu(), recs[1][:, :, 1].double().cpu()
    assert_close(torch.logsumexp(m_c + Z_c.log(), dim=0), ref[:, 0] + ref[:, 1].log(), rtol=1e-6, atol=1e-6, what='record log-sum-exp')
    assert torch.equal(recs[1][:, :, 2].min(dim=0).values.cpu().double(), ref[:, 2]), 'record minimum'
    sc = torch.exp(m_c - ref[:, 0].unsqueeze(0)).unsqueeze(-1)
    assert_close((sc * recs[1][:, :, 4:].double().cpu()).sum(0), ref[:, 4:], rtol=1e-4, atol=1e-4, what='record weighted sum')


def test_mppi_large_batch_properties(dev):
    """C5 (Panda, table + shelf) with 20 000 control samples: size-independent properties."""
    cfg = configs.config('C5')
    N, Tn, d = 20000, 64, 7
    cw = cfg['params']['c_weights']
    planner, cost = _mppi(cfg, N, Tn, dev, 1e-1, cw, cfg['params']['control_std'])
    obs = dict(state=T(cfg['start']).to(**dev), goal_state=T(cfg['goal']).to(**dev), cost=cost)
    mean0 = planner._mean.clone()
    U, X, c = planner.optimize(opt_iters=1, **obs)
    assert U.shape == (N, Tn, d) and X.shape == (N, Tn, d) and c.shape == (N, 1)
    assert torch.equal(X[:, 0], T(cfg['start']).to(**dev).expand(N, d)), 'rollouts start at the observed state'
    # rollout recurrence, bit for bit: x_{t+1} = x_t + clamp(u_t) * dt
    Xr = X[:, :-1] + U[:, :-1].clamp(-100., 100.) * cfg['dt']
    assert torch.equal(Xr, X[:, 1:])
    w = planner.weights.double()
    assert abs(float(w.sum()) - 1.0) < 1e-5
    assert int(planner._best[1][0]) == int(c.argmin())
    assert float(planner.best_cost) == float(c.min())
    assert torch.equal(planner.best_traj, X[int(c.argmin())])
    w64 = torch.softmax(-c.double().reshape(-1), dim=0)
    mean64 = mean0.double() + (w64.view(N, 1, 1) * (U.double() - mean0.double().unsqueeze(0))).sum(0)
    assert_close(planner._mean, mean64, rtol=1e-4, atol=1e-5, what='mean update vs fp64')
    # the obstacle cost enters as a constant shift (quirk B2): weights must not depend on it
    planner2, _ = _mppi(cfg, N, Tn, dev, 1e-1, cw, cfg['params']['control_std'])
    planner2._mean.copy_(mean0)
    planner2.update_ctrl_dist()      # controls are sampled around ctrl_dist.mu (as in the reference)
    eps = torch.randn(d, N, Tn, **dev)
    planner._mean.copy_(mean0)
    planner.update_ctrl_dist()      # controls are sampled around ctrl_dist.mu (as in the reference)
    planner.optimize(opt_iters=1, eps=[eps], **obs)
    planner2.optimize(opt_iters=1, eps=[eps], state=obs['state'], goal_state=obs['goal_state'])
    assert_close(planner2.costs + float(planner._energy), planner.costs, rtol=1e-5, what='constant shift')


def test_rollout_shared_factor_flag_is_bit_identical(dev):
    """mpb_mppi_rollout_opt(MPB_MPPI_SHARED_FACTOR): staging one factor for all control dimensions (identical factors, the
    reference's const_ctrl prior) gives the same rows, quadratic costs and IS dots bit for bit as staging all of them."""
    import ctypes as C
    lib = _lib.lib()
    N, Tn, Cn = 777, 64, 7
    gen = torch.Generator().manual_seed(3)
    A = torch.randn(Tn, Tn, generator=gen) * 0.1
    L1 = torch.linalg.cholesky(A @ A.T + 0.05 * torch.eye(Tn))
    L = L1.unsqueeze(0).repeat(Cn, 1, 1).contiguous().to(**dev)
    Cinv = torch.cholesky_inverse(L1).unsqueeze(0).repeat(Cn, 1, 1).contiguous().to(**dev)
    mean = (0.1 * torch.randn(Tn, Cn, generator=gen)).to(**dev)
    state0, goal = torch.zeros(Cn, **dev), torch.ones(Cn, **dev)
    lo, hi = torch.full((Cn,), -100.0, **dev), torch.full((Cn,), 100.0, **dev)
    nd = _lib.NoiseDesc(seed=5, offset=2, s_offset=0, p_offset=0, P_global=N)
    out = []
    for flags in (0, 1):
        xu, quad, isv = torch.empty(N, Tn, 2 * Cn, **dev), torch.empty(N, **dev), torch.empty(N, Cn, **dev)
        _lib.check(lib.mpb_mppi_rollout_opt(_lib.ptr(L), _lib.ptr(Cinv), _lib.ptr(mean), _lib.ptr(mean), None, C.byref(nd),
                                            _lib.ptr(state0), _lib.ptr(goal), _lib.ptr(lo), _lib.ptr(hi), _lib.ptr(xu), _lib.ptr(quad),
                                            _lib.ptr(isv), N, Tn, Cn, Cn, 0.04, 1.0, 1.0, 1.0, 1000.0, flags, _lib.stream_ptr()))
        out.append((xu, quad, isv))
    torch.cuda.synchronize()
    for a, b in zip(*out):
        assert torch.equal(a, b)
