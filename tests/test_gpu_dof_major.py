"""Dof-major sample rows of the fused Stoch-GPMP iteration (csrc/sample_gp_kron_gen_dm.cu, mpb_cost_eval_dm,
mpb_softmax_update_dm, mpb_stoch_gpmp_iter_kron_gen_dm).  Replaces MultiMPPrior.sample (mp_priors_multi.py:253-256), the
cost evaluation (stoch_gpmp.py:235-245) and the update (stoch_gpmp.py:267-279) exactly as the natural-layout entry points
do -- the layout is internal, so the bar here is BIT-IDENTITY with those entry points (which carry the parity with the
oracle): same samples after the permutation, same costs / flags, same weights / means."""
import ctypes as C
import os

import pytest
import torch

from motion_planning_baselines_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    return dict(device=torch.device('cuda:0'), dtype=torch.float32)


def nd(seed, offset, p_off=0, P_glob=1, s_off=0):
    return _lib.NoiseDesc(seed=seed, offset=offset, s_offset=s_off, p_offset=p_off, P_global=P_glob)


def to_natural(xdm, H, dof):
    B = xdm.shape[0]
    return xdm.view(B, dof, H, 2).permute(0, 2, 3, 1).reshape(B, H, 2 * dof).contiguous()


@pytest.mark.parametrize('variant', ['default', '2'])
@pytest.mark.parametrize('d', [7, 3])
@pytest.mark.parametrize('P,S', [(8, 64), (2, 128), (5, 24), (3, 7), (1, 1), (300, 64)])
def test_dm_sampler_is_bit_identical_to_the_natural_one(P, S, d, variant, dev, monkeypatch):
    """Both dof-major kernels (default: 24 producer warps, roles folded; MPB_DM_VARIANT=2: dedicated loader / MMA / mat-vec
    warps) against the reference-layout tcgen05 sampler."""
    from test_gpu_sample_gen import make_prior
    if variant != 'default':
        monkeypatch.setenv('MPB_DM_VARIANT', variant)
    prior, means = make_prior(P, dev, d=d)
    assert prior.scale_tril_kron_gen is not None
    H, M = 64, 64 * 2 * d
    lib, st = _lib.lib(), _lib.stream_ptr()
    desc = nd(11, 4, p_off=3, P_glob=P + 5)
    x = prior.sample(S, noise_desc=desc).clone().view(P * S, H, 2 * d)
    xdm = torch.full((P * S, M), float('nan'), **dev)
    _lib.check(lib.mpb_sample_gp_kron_gen_dm(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(prior.means), C.byref(desc), _lib.ptr(xdm),
                                             P, S, H, d, None, None, None, st))
    assert torch.equal(to_natural(xdm, H, d), x)
    # the permutation kernels themselves
    back = torch.empty_like(x)
    _lib.check(lib.mpb_traj_from_dof_major(_lib.ptr(xdm), _lib.ptr(back), P * S, H, d, st))
    assert torch.equal(back, x)
    again = torch.empty_like(xdm)
    _lib.check(lib.mpb_traj_to_dof_major(_lib.ptr(x), _lib.ptr(again), P * S, H, d, st))
    assert torch.equal(again, xdm)


@pytest.mark.parametrize('variant', ['default', '2'])
def test_dm_sampler_mat_vec_warp(variant, dev, monkeypatch):
    from test_gpu_sample_gen import make_prior
    if variant != 'default':
        monkeypatch.setenv('MPB_DM_VARIANT', variant)
    P, S, d, H = 37, 64, 7, 64
    prior, means = make_prior(P, dev, d=d)
    lib, st = _lib.lib(), _lib.stream_ptr()
    desc = nd(5, 1, P_glob=P)
    M = H * 2 * d
    y0, y1 = torch.empty(P, M, **dev), torch.empty(P, M, **dev)
    c0, c1 = torch.empty(P, M, **dev), torch.empty(P, M, **dev)
    x = torch.empty(P * S, M, **dev)
    xdm = torch.empty(P * S, M, **dev)
    _lib.check(lib.mpb_sample_gp_kron_gen_mv(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(prior.means), C.byref(desc), _lib.ptr(x),
                                             P, S, H, d, _lib.ptr(prior.Sigma_inv), _lib.ptr(y0), _lib.ptr(c0), st))
    _lib.check(lib.mpb_sample_gp_kron_gen_dm(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(prior.means), C.byref(desc), _lib.ptr(xdm),
                                             P, S, H, d, _lib.ptr(prior.Sigma_inv), _lib.ptr(y1), _lib.ptr(c1), st))
    assert torch.equal(y0, y1) and torch.equal(c0, c1) and torch.equal(c1, prior.means.view(P, M))
    assert torch.equal(to_natural(xdm, H, d), x.view(P * S, H, 2 * d))


def build_arm(obstacles_of, P, dev):
    """The C4 Stoch-GPMP problem (Panda, 64 samples x 64 waypoints) against the obstacles of configuration `obstacles_of`
    ('C4+C5': two fields, the C4 spheres and the C5 boxes -- the mixed sphere / box instance of the cost kernel)."""
    from motion_planning_baselines_b200 import configs
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import StochGPMP
    from motion_planning_baselines_b200.robots import Robot
    cfg = configs.config('C4')
    robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
    fields = [CollisionField(configs.config(name)['obstacles'], tensor_args=dev) for name in obstacles_of.split('+')]
    torch.manual_seed(2024)
    return StochGPMP(robot=robot, n_dof=7, n_support_points=64, num_particles_per_goal=P, opt_iters=1, dt=cfg['dt'],
                     start_state=torch.tensor(cfg['start']).to(**dev), multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0),
                     collision_fields=fields, tensor_args=dev, num_samples=64, **cfg['params'])


@pytest.mark.parametrize('obstacles_of,P', [('C4', 512), ('C5', 96), ('C4', 3), ('C4+C5', 40)])
def test_dm_iteration_is_bit_identical_to_the_natural_one(obstacles_of, P, dev, monkeypatch):
    """Three optimize() calls (C4: the bench shape, 16 spheres; C5: the table + shelf boxes), dof-major against MPB_X_DM=0."""
    out = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('MPB_X_DM', mode)
        pl = build_arm(obstacles_of, P, dev)
        pl._noise.offset = 0
        if mode == '1':
            gp, fields, nf, _ = pl.cost._build()
            assert pl._use_dof_major(fields, nf), 'the dof-major iteration must be the default for the 7-dof arm at H = 64'
        trajs = [pl.optimize(opt_iters=1).clone() for _ in range(3)]
        out[mode] = dict(trajs=trajs, costs=pl.costs.clone(), w=pl._w_buf.clone(), free=pl.free_flags.clone(),
                         x=pl.state_samples.clone(), mu=pl._particle_means.clone(), rec=[t.clone() for t in pl.get_recent_samples()])
    a, b = out['0'], out['1']
    assert torch.equal(a['x'], b['x']), 'samples'
    assert torch.equal(a['costs'], b['costs']), 'costs'
    assert torch.equal(a['free'], b['free']) and torch.equal(a['w'], b['w']) and torch.equal(a['mu'], b['mu'])
    for t0, t1 in zip(a['trajs'], b['trajs']):
        assert torch.equal(t0, t1)
    for r0, r1 in zip(a['rec'], b['rec']):
        assert torch.equal(r0, r1)


def test_dm_kernels_one_by_one(dev):
    """mpb_cost_eval_dm and mpb_softmax_update_dm against their natural-layout twins on the same (permuted) rows."""
    from test_gpu_bench_shape import build
    cfg, sig, pl = build('C4', dev)
    pl.optimize(opt_iters=1)
    lib, st = _lib.lib(), _lib.stream_ptr()
    P, S, H, D = pl.num_particles, pl.num_samples, pl.n_support_points, pl.d_state_opt
    gp, fields, nf, _ = pl.cost._build()
    x = pl.state_samples.clone().view(P * S, H, D)
    xdm = torch.empty(P * S, H * D, **dev)
    _lib.check(lib.mpb_traj_to_dof_major(_lib.ptr(x), _lib.ptr(xdm), P * S, H, D // 2, st))
    c0, c1 = torch.empty(P * S, **dev), torch.empty(P * S, **dev)
    t0, t1 = torch.empty(3, P * S, **dev), torch.empty(3, P * S, **dev)
    f0 = torch.empty(P * S, device=dev['device'], dtype=torch.uint8)
    f1 = torch.empty_like(f0)
    _lib.check(lib.mpb_cost_eval(_lib.ptr(x), P * S, H, C.byref(pl.robot.desc), fields, nf, C.byref(gp), _lib.ptr(pl._is_vec), S,
                                 pl.temperature, _lib.ptr(c0), _lib.ptr(t0), _lib.ptr(f0), st))
    _lib.check(lib.mpb_cost_eval_dm(_lib.ptr(xdm), P * S, H, C.byref(pl.robot.desc), fields, nf, C.byref(gp), _lib.ptr(pl._is_vec), S,
                                    pl.temperature, _lib.ptr(c1), _lib.ptr(t1), _lib.ptr(f1), st))
    assert torch.equal(c0, c1) and torch.equal(f0, f1) and torch.equal(t0, t1)
    mu0 = pl._particle_means.clone()
    mu1 = mu0.clone()
    w0, w1 = torch.empty(P, S, **dev), torch.empty(P, S, **dev)
    g0, g1 = torch.empty(P, H, D, **dev), torch.empty(P, H, D, **dev)
    o0, o1 = torch.empty(P, H, D, **dev), torch.empty(P, H, D, **dev)
    _lib.check(lib.mpb_softmax_update_ex(_lib.ptr(c0), _lib.ptr(x), _lib.ptr(mu0), _lib.ptr(w0), _lib.ptr(g0), 3.0e4, 0.7, None,
                                         _lib.ptr(o0), P, S, H, D, st))         # a warm temperature: many non-zero weights
    _lib.check(lib.mpb_softmax_update_dm(_lib.ptr(c1), _lib.ptr(xdm), _lib.ptr(mu1), _lib.ptr(w1), _lib.ptr(g1), 3.0e4, 0.7,
                                         _lib.ptr(o1), P, S, H, D, st))
    assert int((w0 != 0).sum()) > 4 * P
    assert torch.equal(w0, w1) and torch.equal(g0, g1) and torch.equal(mu0, mu1) and torch.equal(o0, o1)


@pytest.mark.parametrize('mode', ['1', '0'])
def test_step_fused_is_one_optimize_iteration(mode, dev, monkeypatch):
    """StochGPMP.step_fused() (one C call, what bench.py times) against optimize(opt_iters=1) on the same draw of the noise stream."""
    monkeypatch.setenv('MPB_X_DM', mode)
    a, b = build_arm('C4', 24, dev), build_arm('C4', 24, dev)
    a._noise.offset = b._noise.offset = 0
    for _ in range(2):
        traj = a.optimize(opt_iters=1)
        b.step_fused()
    assert torch.equal(traj, b._particle_means) and torch.equal(a._particle_means, b._particle_means)
    assert torch.equal(a.costs, b.costs) and torch.equal(a._w_buf, b._w_buf) and torch.equal(a.free_flags, b.free_flags)
    assert torch.equal(a.state_samples, b.state_samples)
